"""Small procedural scenes used by the parity tests, smoke() and quick benches.

`cornell` is a closed box with coloured walls, an emissive ceiling quad, a metallic and a rough
dielectric block, optional sphere light / spotlight / directional light and an optional
alpha-textured card -- every NEE class and hit type of the hot path in a few hundred triangles.
"""
import numpy as np

from . import abi
from .scene import Scene, make_material, look_at_transform, perspective_camera, camera_pairs


def quad(p0, p1, p2, p3, subdiv=1):
    """Two-triangle (or subdivided) quad p0-p1-p2-p3 (counter-clockwise seen from its front)."""
    p0, p1, p2, p3 = [np.asarray(p, dtype=np.float64) for p in (p0, p1, p2, p3)]
    n = np.cross(p1 - p0, p3 - p0)
    n /= np.linalg.norm(n)
    t = (p1 - p0) / np.linalg.norm(p1 - p0)
    s = subdiv
    verts = np.zeros((s + 1) * (s + 1), dtype=abi.VERTEX)
    k = 0
    for j in range(s + 1):
        for i in range(s + 1):
            u, v = i / s, j / s
            pos = (1 - u) * (1 - v) * p0 + u * (1 - v) * p1 + u * v * p2 + (1 - u) * v * p3
            verts[k]["pos"] = pos
            verts[k]["normal"] = n
            verts[k]["uv"] = (u, v)
            verts[k]["tangent"] = tuple(t) + (1.0,)
            k += 1
    idx = []
    for j in range(s):
        for i in range(s):
            a = j * (s + 1) + i
            idx += [a, a + 1, a + s + 2, a, a + s + 2, a + s + 1]
    return verts, np.asarray(idx, dtype=np.uint32)


def box(lo, hi, subdiv=1):
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    faces = [
        ((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)),  # +z
        ((x1, y0, z0), (x0, y0, z0), (x0, y1, z0), (x1, y1, z0)),  # -z
        ((x1, y0, z1), (x1, y0, z0), (x1, y1, z0), (x1, y1, z1)),  # +x
        ((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)),  # -x
        ((x0, y1, z1), (x1, y1, z1), (x1, y1, z0), (x0, y1, z0)),  # +y
        ((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)),  # -y
    ]
    vs, is_ = [], []
    base = 0
    for f in faces:
        v, i = quad(*f, subdiv=subdiv)
        vs.append(v)
        is_.append(i + base)
        base += v.shape[0]
    return np.concatenate(vs), np.concatenate(is_)


def merge(parts):
    vs, is_ = [], []
    base = 0
    for v, i in parts:
        vs.append(v)
        is_.append(i + base)
        base += v.shape[0]
    return np.concatenate(vs), np.concatenate(is_)


def checker_texture(n=64, cells=8, alpha_holes=True, seed=3):
    rng = np.random.RandomState(seed)
    tex = np.zeros((n, n, 4), dtype=np.uint8)
    yy, xx = np.mgrid[0:n, 0:n]
    c = ((xx * cells // n) + (yy * cells // n)) % 2
    tex[..., 0] = np.where(c, 230, 40)
    tex[..., 1] = np.where(c, 200, 90)
    tex[..., 2] = rng.randint(20, 250, size=(n, n))
    if alpha_holes:
        r = np.hypot(xx - n / 2, yy - n / 2)
        tex[..., 3] = np.clip((r - n * 0.15) * 16, 0, 255).astype(np.uint8)
    else:
        tex[..., 3] = 255
    return tex


def cornell(subdiv=4, sphere_light=True, spotlight=True, directional=True, alpha_card=True,
            transmissive=True, rotate_block=True):
    s = Scene()
    w = 2.0
    white = make_material(albedo=(0.8, 0.8, 0.8, 1))
    red = make_material(albedo=(0.8, 0.08, 0.02, 1))
    blue = make_material(albedo=(0.02, 0.05, 0.8, 1), roughness=0.5)
    mirror = make_material(albedo=(0.9, 0.9, 0.9, 1), roughness=0.0)
    # floor, ceiling, back wall (one instance), then the side walls and a smooth front-left wall
    fl = quad((-w, -w, w), (w, -w, w), (w, -w, -w), (-w, -w, -w), subdiv)
    ce = quad((-w, w, -w), (w, w, -w), (w, w, w), (-w, w, w), subdiv)
    ba = quad((-w, -w, -w), (w, -w, -w), (w, w, -w), (-w, w, -w), subdiv)
    s.add_instance(*merge([fl, ce, ba]), white)
    s.add_instance(*quad((-w, -w, w), (-w, -w, -w), (-w, w, -w), (-w, w, w), subdiv), red)
    s.add_instance(*quad((w, -w, -w), (w, -w, w), (w, w, w), (w, w, -w), subdiv), blue)
    # behind the camera: a mirror wall closing the box
    s.add_instance(*quad((w, -w, 3 * w), (-w, -w, 3 * w), (-w, w, 3 * w), (w, w, 3 * w), 1), mirror)
    for zs in (w, 2 * w):
        pass
    side = merge([
        quad((-w, -w, 3 * w), (-w, -w, w), (-w, w, w), (-w, w, 3 * w), 1),
        quad((w, -w, w), (w, -w, 3 * w), (w, w, 3 * w), (w, w, w), 1),
        quad((-w, -w, 3 * w), (w, -w, 3 * w), (w, -w, w), (-w, -w, w), 1),
        quad((-w, w, w), (w, w, w), (w, w, 3 * w), (-w, w, 3 * w), 1),
    ])
    s.add_instance(*side, white)
    # emissive ceiling panel
    em = make_material(albedo=(0, 0, 0, 1), emission=(12.0, 11.0, 8.0))
    s.add_instance(*quad((-0.5, w - 0.01, -0.5), (0.5, w - 0.01, -0.5), (0.5, w - 0.01, 0.5), (-0.5, w - 0.01, 0.5), 2), em)
    # blocks
    metal = make_material(albedo=(0.72, 0.22, 0.05, 1), metallic=1.0, roughness=0.2)
    bv, bi = box((-0.6, -0.6, -0.6), (0.6, 0.6, 0.6), 2)
    t = np.eye(4)
    if rotate_block:
        a = 0.5
        t[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1.3, 0], [-np.sin(a), 0, np.cos(a)]]
    t[:3, 3] = (-0.8, -w + 0.78, -0.7)
    s.add_instance(bv, bi, metal, transform=t)
    glass = make_material(albedo=(0.48, 0.6, 0.28, 1), roughness=0.2, transmittance=1.0 if transmissive else 0.0)
    t2 = np.eye(4)
    t2[:3, 3] = (0.9, -w + 0.5, 0.4)
    t2[0, 0] = t2[1, 1] = t2[2, 2] = 0.8
    s.add_instance(bv, bi, glass, transform=t2)
    if alpha_card:
        tid = s.add_texture(checker_texture())
        card = make_material(albedo=(1, 1, 1, 1), albedo_tex=tid, roughness=0.6)
        s.add_instance(*quad((-1.2, -1.0, 1.0), (0.2, -1.0, 1.0), (0.2, 0.4, 1.0), (-1.2, 0.4, 1.0), 1), card)
    if directional:
        s.add_directional_light((1.0, 0.87, 0.6), (-0.3, -1.0, -0.4), angle_deg=2.0)
    if sphere_light:
        s.add_point_light((3.0, 3.0, 4.0), (1.2, 1.2, 0.2), radius=0.15)
    if spotlight:
        s.add_spotlight((20.0, 40.0, 16.0), (-1.5, 1.5, 1.5), (0.6, -1.0, -0.8), 25.0, 2.0, radius=0.05)
    return s


def cornell_camera(aspect=1.0, fov_deg=45.0):
    return perspective_camera(look_at_transform((0, 0, 5.5), (0, -0.2, 0)), fov_deg, aspect, 0.1, 100.0)


def cornell_cameras(aspect=1.0):
    return camera_pairs([cornell_camera(aspect)])


def sky_texels(width=64, height=32, sun=True, seed=11):
    """Procedural lat-long HDR sky (H, W, 4) float16: gradient + noise + a small very bright sun
    patch, i.e. an importance distribution with light and heavy texels for the alias table."""
    rng = np.random.RandomState(seed)
    v = (np.arange(height) + 0.5) / height
    u = (np.arange(width) + 0.5) / width
    uu, vv = np.meshgrid(u, v)
    up = np.clip(1.0 - 2.0 * vv, 0.0, 1.0)           # v = 0 is straight up (dir.y = +1)
    rgb = np.stack([0.25 + 0.3 * up, 0.35 + 0.45 * up, 0.5 + 0.9 * up], axis=-1)
    rgb[vv > 0.5] = (0.12, 0.1, 0.08)                # ground
    rgb *= 0.75 + 0.5 * rng.rand(height, width, 1)
    if sun:
        d = np.hypot((uu - 0.62) * 2.0, vv - 0.22)
        rgb += np.exp(-(d / 0.035) ** 2)[..., None] * np.array([260.0, 230.0, 180.0])
    out = np.ones((height, width, 4), dtype=np.float16)
    out[..., :3] = rgb.astype(np.float16)
    return out


def courtyard(subdiv=3, envmap=True, sun_in_sky=True, directional=False, emissive=False, env_size=(64, 32)):
    """Open scene lit by an environment map: ground plane, a metal and a diffuse block, a glass
    block and an alpha-textured card. Camera rays also see the sky directly."""
    s = Scene()
    ground = make_material(albedo=(0.7, 0.68, 0.6, 1), roughness=0.9)
    s.add_instance(*quad((-6, -1, 6), (6, -1, 6), (6, -1, -6), (-6, -1, -6), subdiv * 2), ground)
    bv, bi = box((-0.6, -0.6, -0.6), (0.6, 0.6, 0.6), subdiv)
    metal = make_material(albedo=(0.9, 0.6, 0.2, 1), metallic=1.0, roughness=0.15)
    t = np.eye(4)
    a = 0.6
    t[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1.5, 0], [-np.sin(a), 0, np.cos(a)]]
    t[:3, 3] = (-1.1, -0.1, -0.5)
    s.add_instance(bv, bi, metal, transform=t)
    diffuse = make_material(albedo=(0.2, 0.5, 0.8, 1), roughness=0.6)
    t2 = np.eye(4)
    t2[:3, 3] = (0.9, -0.4, 0.3)
    s.add_instance(bv, bi, diffuse, transform=t2)
    glass = make_material(albedo=(0.9, 0.95, 0.9, 1), roughness=0.05, transmittance=1.0)
    t3 = np.eye(4)
    t3[0, 0] = t3[1, 1] = t3[2, 2] = 0.6
    t3[:3, 3] = (0.0, -0.64, 1.6)
    s.add_instance(bv, bi, glass, transform=t3)
    tid = s.add_texture(checker_texture())
    card = make_material(albedo=(1, 1, 1, 1), albedo_tex=tid, roughness=0.6)
    s.add_instance(*quad((-2.4, -1.0, 0.8), (-1.2, -1.0, 1.4), (-1.2, 0.3, 1.4), (-2.4, 0.3, 0.8), 1), card)
    if emissive:
        em = make_material(albedo=(0, 0, 0, 1), emission=(6.0, 3.0, 1.0))
        s.add_instance(*quad((1.6, -0.99, 1.8), (2.4, -0.99, 1.8), (2.4, -0.99, 1.0), (1.6, -0.99, 1.0), 1), em)
    if directional:
        s.add_directional_light((2.0, 1.8, 1.5), (-0.4, -1.0, -0.3), angle_deg=1.0)
    if envmap:
        s.set_environment_map(sky_texels(env_size[0], env_size[1], sun=sun_in_sky), factor=(1.0, 0.9, 0.8))
    return s


def courtyard_cameras(aspect=1.0, fov_deg=55.0):
    cam = perspective_camera(look_at_transform((0.3, 0.6, 5.0), (0, -0.1, 0)), fov_deg, aspect, 0.1, 100.0)
    return camera_pairs([cam])


def deep_chain(shells=44, emissive=True):
    """A scene whose LBVH degenerates into a long chain: shells of small quads shrinking geometrically towards one
    corner of a room (each shell is half the size of the previous one, so every Morton split peels one shell off).
    Used to exercise deep traversal stacks (sibling groups + postponed triangle groups)."""
    s = Scene()
    w = 2.0
    white = make_material(albedo=(0.8, 0.8, 0.8, 1))
    room = merge([
        quad((-w, -w, w), (w, -w, w), (w, -w, -w), (-w, -w, -w), 2),
        quad((-w, w, -w), (w, w, -w), (w, w, w), (-w, w, w), 2),
        quad((-w, -w, -w), (w, -w, -w), (w, w, -w), (-w, w, -w), 2),
        quad((-w, -w, w), (-w, -w, -w), (-w, w, -w), (-w, w, w), 2),
        quad((w, -w, -w), (w, -w, w), (w, w, w), (w, w, -w), 2),
    ])
    s.add_instance(*room, white)
    corner = np.array([-w + 0.05, -w + 0.05, -w + 0.05])
    parts = []
    for k in range(shells):
        r = 3.0 * 0.5 ** k
        c = corner + r * np.array([1.0, 0.8, 0.6])
        h = 0.22 * r
        parts.append(quad(c + (-h, -h, 0), c + (h, -h, 0), c + (h, h, 0), c + (-h, h, 0), 1))
        parts.append(quad(c + (0, -h, -h), c + (0, -h, h), c + (0, h, h), c + (0, h, -h), 1))
        parts.append(quad(c + (-h, 0, h), c + (h, 0, h), c + (h, 0, -h), c + (-h, 0, -h), 1))
    s.add_instance(*merge(parts), make_material(albedo=(0.9, 0.5, 0.2, 1), roughness=0.4))
    if emissive:
        em = make_material(albedo=(0, 0, 0, 1), emission=(10.0, 10.0, 9.0))
        s.add_instance(*quad((-0.6, w - 0.01, -0.6), (0.6, w - 0.01, -0.6), (0.6, w - 0.01, 0.6), (-0.6, w - 0.01, 0.6), 1), em)
    s.add_directional_light((1.0, 0.9, 0.7), (-0.4, -1.0, -0.3), angle_deg=1.0)
    return s


def deep_chain_cameras(aspect=1.0):
    return camera_pairs([perspective_camera(look_at_transform((1.6, 1.2, 1.8), (-1.9, -1.9, -1.9)), 50.0, aspect, 0.01, 100.0)])
