"""Deterministic synthetic "Sponza-class" atrium: the benchmark workload of BASELINE.json.

SURVEY.md section 8(d): a 30 x 12 x 14 m hall with an open skylight, two arcades of 2 x 12 tessellated
columns joined by half-torus arches, an upper gallery, and 400 noise-displaced icospheres, tessellated to
EXACTLY `triangles` (default 1,500,000) triangles; u32 indices, 48-byte vertices, world space (identity
instance transforms, like Tauray's merged static BLAS). 8 opaque GGX materials (no textures, alpha 1,
transmittance 0 -> any-hit never fires), 64 emissive ceiling-lamp triangles (emission 20) and one 0.5 degree
directional sun of colour 5; no environment map, no point lights. Seed 0x5EED5EED. No files are read.
"""
import math

import numpy as np

from . import abi
from .scene import Scene, make_material, look_at_transform, perspective_camera, camera_pairs, camera_grid

SEED = 0x5EED5EED


# ---- primitive builders (all return (pos[N,3], normal[N,3], uv[N,2], idx[M])) --------------------
def _grid(p0, du, dv, nu, nv):
    p0, du, dv = (np.asarray(a, dtype=np.float64) for a in (p0, du, dv))
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    uu, vv = np.meshgrid(u, v)
    pos = p0[None, None, :] + uu[..., None] * du[None, None, :] + vv[..., None] * dv[None, None, :]
    n = np.cross(du, dv)
    n /= np.linalg.norm(n)
    nor = np.broadcast_to(n, pos.shape)
    uv = np.stack([uu, vv], axis=-1)
    i = np.arange(nu)[None, :] + (nu + 1) * np.arange(nv)[:, None]
    a, b, c, d = i, i + 1, i + nu + 2, i + nu + 1
    idx = np.stack([a, b, c, a, c, d], axis=-1).reshape(-1)
    return pos.reshape(-1, 3), nor.reshape(-1, 3).copy(), uv.reshape(-1, 2), idx


def _surface(fn, nu, nv, wrap_u=False):
    """Parametric surface fn(u, v) -> (pos, normal) on a (nu x nv) grid."""
    u = np.linspace(0.0, 1.0, nu + 1)
    v = np.linspace(0.0, 1.0, nv + 1)
    uu, vv = np.meshgrid(u, v)
    pos, nor = fn(uu, vv)
    uv = np.stack([uu, vv], axis=-1)
    i = np.arange(nu)[None, :] + (nu + 1) * np.arange(nv)[:, None]
    a, b, c, d = i, i + 1, i + nu + 2, i + nu + 1
    idx = np.stack([a, b, c, a, c, d], axis=-1).reshape(-1)
    return pos.reshape(-1, 3), nor.reshape(-1, 3), uv.reshape(-1, 2), idx


def _cylinder(cx, cz, y0, y1, radius, nseg, nrings, flare=0.0):
    def fn(u, v):
        ang = 2 * math.pi * u
        r = radius * (1.0 + flare * (np.abs(2 * v - 1) ** 6))
        pos = np.stack([cx + r * np.cos(ang), y0 + (y1 - y0) * v, cz + r * np.sin(ang)], axis=-1)
        nor = np.stack([np.cos(ang), np.zeros_like(ang), np.sin(ang)], axis=-1)
        return pos, nor
    p, n, uv, idx = _surface(fn, nseg, nrings)
    # outward orientation: (du x dv) points inward for this parametrisation -> flip winding
    return p, n, uv, idx.reshape(-1, 3)[:, ::-1].reshape(-1)


def _arch(x0, x1, y, z, minor, nmaj, nmin):
    """Half torus spanning x0..x1 at height y (springing line), in the plane z = const."""
    cx, R = 0.5 * (x0 + x1), 0.5 * (x1 - x0)

    def fn(u, v):
        a = math.pi * u            # along the arch
        b = 2 * math.pi * v        # around the tube
        cxa, sya = np.cos(a), np.sin(a)
        nx = cxa * np.cos(b)
        ny = sya * np.cos(b)
        nz = np.sin(b)
        pos = np.stack([cx - (R + minor * np.cos(b)) * cxa, y + (R + minor * np.cos(b)) * sya, z + minor * nz], axis=-1)
        nor = np.stack([-nx, ny, nz], axis=-1)
        return pos, nor
    return _surface(fn, nmaj, nmin)


_ICO_CACHE = {}


def _icosphere(level):
    if level in _ICO_CACHE:
        return _ICO_CACHE[level]
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                  [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(level):
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        key = np.sort(edges, axis=1)
        uniq, inv = np.unique(key, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = v.shape[0]
        v = np.concatenate([v, mid])
        n = f.shape[0]
        m01, m12, m20 = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([
            np.stack([f[:, 0], m01, m20], axis=1), np.stack([f[:, 1], m12, m01], axis=1),
            np.stack([f[:, 2], m20, m12], axis=1), np.stack([m01, m12, m20], axis=1)])
    _ICO_CACHE[level] = (v, f)
    return v, f


def _noise3(p, freq, phase):
    """Smooth deterministic pseudo-noise in [-1, 1] (sum of sines; no tables, no files)."""
    q = p * freq
    return (np.sin(q[:, 0] * 1.7 + phase[0]) * np.sin(q[:, 1] * 2.3 + phase[1]) * np.sin(q[:, 2] * 1.3 + phase[2])
            + 0.5 * np.sin(q[:, 0] * 3.1 + q[:, 1] * 2.7 + phase[1]) * np.sin(q[:, 2] * 3.7 + phase[0])) / 1.5


class _Part:
    def __init__(self):
        self.pos, self.nor, self.uv, self.idx = [], [], [], []
        self.nverts = 0

    def add(self, p, n, uv, idx):
        self.pos.append(np.asarray(p, dtype=np.float64))
        self.nor.append(np.asarray(n, dtype=np.float64))
        self.uv.append(np.asarray(uv, dtype=np.float64))
        self.idx.append(np.asarray(idx, dtype=np.int64) + self.nverts)
        self.nverts += p.shape[0]

    @property
    def triangles(self):
        return sum(i.size for i in self.idx) // 3

    def finish(self):
        v = np.zeros(self.nverts, dtype=abi.VERTEX)
        if self.nverts:
            v["pos"] = np.concatenate(self.pos).astype(np.float32)
            n = np.concatenate(self.nor)
            n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-20)
            v["normal"] = n.astype(np.float32)
            v["uv"] = np.concatenate(self.uv).astype(np.float32)
            v["tangent"] = (1.0, 0.0, 0.0, 1.0)
        idx = np.concatenate(self.idx).astype(np.uint32) if self.idx else np.zeros(0, dtype=np.uint32)
        return v, idx


def atrium(triangles=1_500_000, spheres=400, sphere_level=3, seed=SEED):
    """Build the benchmark scene. Returns a Scene with exactly `triangles` triangles."""
    rng = np.random.RandomState(seed & 0x7FFFFFFF)
    LX, LY, LZ = 15.0, 12.0, 7.0     # half length, height, half width: 30 x 12 x 14 m
    parts = {k: _Part() for k in ("floor", "walls", "ceiling", "columns", "arches", "gallery", "metal", "paint", "lamps")}

    # ---- scale the fixed tessellation with the triangle budget --------------------------------------
    scale = max(triangles / 1_500_000.0, 1e-3)
    s1 = math.sqrt(scale)
    ico_v, ico_f = _icosphere(sphere_level)
    nspheres = max(int(round(spheres * min(scale, 1.0))), 0) if triangles < 1_500_000 else spheres
    col_seg, col_rings = max(int(32 * s1), 6), max(int(64 * s1), 4)
    arch_maj, arch_min = max(int(64 * s1), 4), max(int(24 * s1), 4)
    wall_n = max(int(128 * s1), 2)

    # ---- lamps: 32 quads = 64 emissive triangles under the gallery slabs -------------------------
    for side in (-1.0, 1.0):
        for k in range(16):
            x = -LX + 1.5 + k * (2 * LX - 3.0) / 15.0
            z = side * (LZ - 1.25)
            p = _grid((x - 0.35, 5.78, z - 0.35), (0.7, 0, 0), (0, 0, 0.7), 1, 1)
            # facing down: flip winding
            parts["lamps"].add(p[0], -p[1], p[2], p[3].reshape(-1, 3)[:, ::-1].reshape(-1))

    # ---- shell: walls, ceiling with a skylight (open to the sun) -----------------------------------
    W = parts["walls"]
    W.add(*_grid((-LX, 0, -LZ), (2 * LX, 0, 0), (0, LY, 0), 2 * wall_n, wall_n))            # -z wall, faces +z
    W.add(*_grid((LX, 0, LZ), (-2 * LX, 0, 0), (0, LY, 0), 2 * wall_n, wall_n))             # +z wall, faces -z
    W.add(*_grid((-LX, 0, LZ), (0, 0, -2 * LZ), (0, LY, 0), wall_n, wall_n))                # -x wall, faces +x
    W.add(*_grid((LX, 0, -LZ), (0, 0, 2 * LZ), (0, LY, 0), wall_n, wall_n))                 # +x wall, faces -x
    C = parts["ceiling"]
    sky_x, sky_z = 10.0, 3.0       # half extents of the skylight opening
    cn = max(int(48 * s1), 1)
    C.add(*_grid((-LX, LY, -LZ), (2 * LX, 0, 0), (0, 0, LZ - sky_z), 4 * cn, cn))           # strips around the opening
    C.add(*_grid((-LX, LY, sky_z), (2 * LX, 0, 0), (0, 0, LZ - sky_z), 4 * cn, cn))
    C.add(*_grid((-LX, LY, -sky_z), (LX - sky_x, 0, 0), (0, 0, 2 * sky_z), cn, cn))
    C.add(*_grid((sky_x, LY, -sky_z), (LX - sky_x, 0, 0), (0, 0, 2 * sky_z), cn, cn))
    for part in (C,):  # ceiling faces down
        for k in range(len(part.idx)):
            part.idx[k] = part.idx[k].reshape(-1, 3)[:, ::-1].reshape(-1)
            part.nor[k] = -part.nor[k]

    # ---- arcades: 2 sides x 2 storeys x 12 columns, arches between neighbours ------------------------
    col_x = np.linspace(-LX + 2.0, LX - 2.0, 12)
    for side in (-1.0, 1.0):
        z = side * (LZ - 2.5)
        for storey, (y0, y1, r) in enumerate(((0.0, 4.2, 0.32), (6.0, 9.6, 0.24))):
            for x in col_x:
                parts["columns"].add(*_cylinder(x, z, y0, y1, r, col_seg, col_rings, flare=0.35))
            for x0, x1 in zip(col_x[:-1], col_x[1:]):
                parts["arches"].add(*_arch(x0, x1, y1, z, 0.14 if storey == 0 else 0.11, arch_maj, arch_min))

    # ---- upper gallery: slabs + balustrade along both long walls ------------------------------------
    G = parts["gallery"]
    gn = max(int(40 * s1), 1)
    for side in (-1.0, 1.0):
        z_in, z_out = side * (LZ - 2.9), side * LZ
        za, zb = min(z_in, z_out), max(z_in, z_out)
        G.add(*_grid((-LX, 6.0, za), (0, 0, zb - za), (2 * LX, 0, 0), gn, 8 * gn))            # top, faces +y
        p = _grid((-LX, 5.8, za), (2 * LX, 0, 0), (0, 0, zb - za), 8 * gn, gn)                 # underside, faces -y
        G.add(p[0], p[1], p[2], p[3])
        # front face of the slab and the balustrade
        front = _grid((-LX, 5.8, z_in), (2 * LX, 0, 0), (0, 0.2, 0), 8 * gn, 1)
        rail = _grid((-LX, 6.0, z_in), (2 * LX, 0, 0), (0, 0.9, 0), 8 * gn, max(gn // 4, 1))
        for q in (front, rail):
            if side < 0:   # must face +z (towards the nave)
                G.add(*q)
            else:
                G.add(q[0], -q[1], q[2], q[3].reshape(-1, 3)[:, ::-1].reshape(-1))

    # ---- 400 noise-displaced icospheres (pots on the floor, drapes/lanterns hanging) ----------------
    for k in range(nspheres):
        hanging = k % 3 == 0
        x = rng.uniform(-LX + 1.0, LX - 1.0)
        z = rng.uniform(-LZ + 3.2, LZ - 3.2) if not hanging else rng.uniform(-LZ + 0.8, LZ - 0.8)
        radius = rng.uniform(0.18, 0.55)
        y = radius * 0.92 if not hanging else rng.uniform(7.0, 10.5)
        phase = rng.uniform(0, 6.28, size=3)
        amp = rng.uniform(0.05, 0.22)
        stretch = np.array([1.0, rng.uniform(0.8, 1.9) if hanging else rng.uniform(0.7, 1.3), 1.0])
        d = 1.0 + amp * _noise3(ico_v, 3.0, phase)
        p = ico_v * d[:, None] * radius * stretch[None, :] + np.array([x, y, z])[None, :]
        uv = np.stack([np.arctan2(ico_v[:, 2], ico_v[:, 0]) / (2 * math.pi) + 0.5, np.arccos(np.clip(ico_v[:, 1], -1, 1)) / math.pi], axis=-1)
        parts["metal" if k % 2 == 0 else "paint"].add(p, ico_v / stretch[None, :], uv, ico_f.reshape(-1))

    # ---- floor: takes whatever is left of the budget -----------------------------------------------
    used = sum(p.triangles for p in parts.values())
    remaining = triangles - used
    if remaining < 2:
        raise ValueError("triangle budget %d too small for the fixed geometry (%d)" % (triangles, used))
    F = parts["floor"]
    # nu x nv grid with 2*nu*nv <= remaining, aspect ~ 30:14, then a filler fan strip for the rest
    nv = max(int(math.sqrt(remaining / 2 * (2 * LZ) / (2 * LX))), 1)
    nu = max((remaining // 2) // nv, 1)
    while 2 * nu * nv > remaining:
        nu -= 1
    F.add(*_grid((-LX, 0, LZ), (2 * LX, 0, 0), (0, 0, -2 * LZ), nu, nv))
    rest = remaining - 2 * nu * nv
    if rest > 0:
        # a strip of `rest` thin triangles (a raised threshold at the far end wall): k+2 vertices in a zig-zag
        t = np.arange(rest + 2)
        zz = -LZ + 0.5 + (2 * LZ - 1.0) * (t // 2) / max((rest + 1) // 2, 1)
        pos = np.stack([np.where(t % 2 == 0, LX - 0.6, LX - 0.3), np.full(t.shape, 0.02), zz], axis=-1)
        nor = np.broadcast_to(np.array([0.0, 1.0, 0.0]), pos.shape)
        uv = np.stack([(t % 2).astype(np.float64), (t // 2) / max(rest / 2, 1)], axis=-1)
        i = np.arange(rest)
        tri = np.where((i % 2 == 0)[:, None], np.stack([i, i + 2, i + 1], axis=-1), np.stack([i, i + 1, i + 2], axis=-1))
        F.add(pos, nor, uv, tri.reshape(-1))

    # ---- materials: 8 opaque GGX + the lamp emitter ----------------------------------------------
    mats = {
        "floor": make_material(albedo=(0.55, 0.52, 0.48, 1), roughness=0.5),
        "walls": make_material(albedo=(0.78, 0.70, 0.58, 1), roughness=1.0),
        "ceiling": make_material(albedo=(0.82, 0.80, 0.76, 1), roughness=1.0),
        "columns": make_material(albedo=(0.85, 0.83, 0.78, 1), roughness=0.5),
        "arches": make_material(albedo=(0.62, 0.35, 0.25, 1), roughness=1.0),
        "gallery": make_material(albedo=(0.42, 0.27, 0.16, 1), roughness=0.5),
        "metal": make_material(albedo=(0.95, 0.78, 0.36, 1), roughness=0.2, metallic=1.0),
        "paint": make_material(albedo=(0.12, 0.32, 0.62, 1), roughness=0.2),
        "lamps": make_material(albedo=(0, 0, 0, 1), emission=(20.0, 20.0, 20.0)),
    }
    scene = Scene()
    for name in ("floor", "walls", "ceiling", "columns", "arches", "gallery", "metal", "paint", "lamps"):
        v, idx = parts[name].finish()
        if idx.size:
            scene.add_instance(v, idx, mats[name], opaque=True)
    # sun through the skylight: angle 0.5 degrees, colour 5
    scene.add_directional_light((5.0, 5.0, 5.0), (0.25, -1.0, 0.12), angle_deg=0.5)
    assert scene.triangle_count == triangles, (scene.triangle_count, triangles)
    return scene


def atrium_camera_transform():
    return look_at_transform((-13.5, 2.2, 0.0), (10.0, 3.2, 0.0))


def atrium_cameras(aspect=16.0 / 9.0):
    """Perspective, vfov 60, near 0.1, far 100, at the nave's end looking down its axis."""
    return camera_pairs([perspective_camera(atrium_camera_transform(), 60.0, aspect, 0.1, 100.0)])


def atrium_camera_grid(aspect=16.0 / 9.0, grid=(5, 9), spacing=(0.02, 0.02)):
    """BASELINE config 5: --camera-grid=5,9,0.02,0.02 (45 views)."""
    return camera_pairs(camera_grid(atrium_camera_transform(), 60.0, aspect, grid[0], grid[1], spacing[0], spacing[1], 0.1, 100.0))
