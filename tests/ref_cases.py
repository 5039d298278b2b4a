"""Input generators and option sets for the function-level parity between the oracle's restatement and the reference's
own shader text (oracle/_ref, built from /root/reference by oracle/ref_glsl/build_ref.py).

VARIANTS maps a shader permutation of build_ref.py to the tr_pt_options that select the same code in the oracle.
FUNCTIONS: id -> (name, kind of every output word, generator of packed input records)."""
import numpy as np

from tauray_b200 import abi

VARIANTS = {
    "v0": dict(max_ray_depth=4, local_sampler=abi.SAMPLER_SOBOL_OWEN, tri_light_mode=abi.TRI_LIGHT_SOLID_ANGLE, bounce_mode=abi.BOUNCE_MATERIAL),
    "v1": dict(max_ray_depth=8, local_sampler=abi.SAMPLER_UNIFORM_RANDOM, tri_light_mode=abi.TRI_LIGHT_SOLID_ANGLE, bounce_mode=abi.BOUNCE_MATERIAL),
    "v2": dict(max_ray_depth=3, local_sampler=abi.SAMPLER_SOBOL_Z_ORDER_3D, tri_light_mode=abi.TRI_LIGHT_AREA, bounce_mode=abi.BOUNCE_HEMISPHERE,
               mis_mode=abi.MIS_BALANCE_HEURISTIC, film=abi.FILM_BOX, projection=abi.PROJECTION_ORTHOGRAPHIC),
    "v3": dict(max_ray_depth=9, local_sampler=abi.SAMPLER_SOBOL_Z_ORDER_2D, tri_light_mode=abi.TRI_LIGHT_HYBRID, bounce_mode=abi.BOUNCE_COSINE_HEMISPHERE,
               mis_mode=abi.MIS_DISABLED, film=abi.FILM_BLACKMAN_HARRIS, depth_of_field=1, russian_roulette_delta=2.0, regularization_gamma=0.3),
    "v4": dict(max_ray_depth=2, local_sampler=abi.SAMPLER_SOBOL_OWEN, sampling_weight_point_lights=0.0, sampling_weight_directional_lights=0.0,
               sampling_weight_envmap=0.0, sampling_weight_emissive_triangles=0.0, hide_lights=1, use_white_albedo_on_first_bounce=1,
               transparent_background=1),
}


def f2u(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def unit(rng, n, k=1):
    return rng.uniform(0, 1, size=(n, k)).astype(np.float32)


def direction(rng, n):
    d = rng.normal(size=(n, 3))
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def upper(rng, n):
    d = direction(rng, n)
    d[:, 2] = np.abs(d[:, 2]) + 1e-3
    return (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)


def u32(rng, n, k=1):
    return rng.randint(0, 2 ** 32, size=(n, k), dtype=np.uint64).astype(np.uint32)


def material(rng, n):
    """albedo4, metallic, roughness (squared like sample_material does), transmittance, ior_in, ior_out, f0"""
    m = np.zeros((n, 10), dtype=np.float32)
    m[:, :3] = rng.uniform(0.02, 1.0, size=(n, 3))
    m[:, 3] = 1.0
    m[:, 4] = rng.choice([0.0, 1.0, 0.5], size=n)
    # roughness >= 0.3 before squaring: below that the GGX distribution term a^2 / (pi ((n.h)^2 (a^2 - 1) + 1)^2) cancels
    # catastrophically in binary32 and any two correct implementations differ by percents (0 = the delta-lobe branch)
    rough = rng.choice([0.0, 0.3, 0.5, 0.7, 1.0], size=n) * rng.uniform(0.9, 1.0, size=n)
    m[:, 5] = rough * rough
    m[:, 6] = rng.choice([0.0, 0.0, 1.0, 0.4], size=n)
    ior = rng.choice([1.45, 1.33, 2.0], size=n)
    back = rng.uniform(size=n) < 0.3
    m[:, 7] = np.where(back & (m[:, 6] > 0), ior, 1.0)
    m[:, 8] = np.where(back & (m[:, 6] > 0), 1.0, ior)
    f0 = (m[:, 8] - m[:, 7]) / (m[:, 8] + m[:, 7])
    m[:, 9] = (f0.astype(np.float32) * f0.astype(np.float32))
    return m


def triangle(rng, n, scale=3.0):
    """A, B, C relative to the shading point, not degenerate, not through the origin"""
    c = rng.uniform(-scale, scale, size=(n, 3)) + np.sign(rng.uniform(-1, 1, size=(n, 3))) * 0.5
    a = c + rng.uniform(-1, 1, size=(n, 3))
    b = c + rng.uniform(-1, 1, size=(n, 3))
    e = c + rng.uniform(-1, 1, size=(n, 3))
    return a.astype(np.float32), b.astype(np.float32), e.astype(np.float32)


def point_lights(rng, n):
    pl = np.zeros(n, dtype=abi.POINT_LIGHT)
    pl["color"] = rng.uniform(0.5, 20, size=(n, 3))
    pl["dir"] = direction(rng, n)
    pl["pos"] = rng.uniform(-3, 3, size=(n, 3))
    pl["radius"] = rng.choice([0.0, 0.05, 0.3], size=n)
    spot = rng.uniform(size=n) < 0.5
    pl["dir_cutoff"] = np.where(spot, np.cos(np.radians(rng.uniform(10, 60, size=n))), 0.0)
    pl["dir_falloff"] = np.where(spot, rng.uniform(0.5, 4, size=n), 0.0)
    pl["spot_radius"] = -1.0
    pl["shadow_map_index"] = -1
    return pl


def directional_lights(rng, n):
    dl = np.zeros(n, dtype=abi.DIRECTIONAL_LIGHT)
    dl["color"] = rng.uniform(0.5, 5, size=(n, 3))
    dl["shadow_map_index"] = -1
    dl["dir"] = direction(rng, n)
    dl["dir_cutoff"] = np.where(rng.uniform(size=n) < 0.3, 1.0, np.cos(np.radians(rng.uniform(0.1, 10, size=n)))).astype(np.float32)
    return dl


def cameras(rng, n, ortho):
    from tauray_b200.scene import look_at_transform, orthographic_camera, perspective_camera
    out = np.zeros(n, dtype=abi.CAMERA_DATA)
    for i in range(n):
        eye = rng.uniform(-4, 4, size=3)
        t = look_at_transform(eye, eye + direction(rng, 1)[0] + (0, 0, 1e-3))
        if ortho:
            out[i] = orthographic_camera(t, -2, 2, -1.5, 1.5, 0.1, 50.0)
        else:
            sides = rng.choice([0.0, 0.0, 5.0, 6.0])
            out[i] = perspective_camera(t, rng.uniform(30, 90), rng.uniform(1.0, 2.0), 0.1, 100.0,
                                        focus=(rng.uniform(1, 8), rng.uniform(0.0, 0.1), rng.uniform(0, 1), sides))
    return out


def cat(*parts):
    cols = []
    for p in parts:
        p = np.asarray(p)
        if p.dtype != np.uint32:
            p = f2u(p.astype(np.float32))
        cols.append(p.reshape(p.shape[0], -1))
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def structs(a):
    return np.ascontiguousarray(a).view(np.uint32).reshape(a.shape[0], -1)


def _tri_inputs(rng, n, lead):
    a, b, c = triangle(rng, n)
    return cat(lead, a, b, c)


def _point_on_triangle(rng, n):
    a, b, c = triangle(rng, n)
    w = rng.dirichlet((1, 1, 1), size=n).astype(np.float32)
    p = w[:, :1] * a + w[:, 1:2] * b + w[:, 2:] * c
    return cat(p, a, b, c)


I, F = "i", "f"
# id: (name, output word kinds, input generator(rng, n, variant))
FUNCTIONS = {
    0: ("pcg", I * 2, lambda r, n, v: u32(r, n, 1)),
    1: ("pcg4d", I * 8, lambda r, n, v: u32(r, n, 4)),
    2: ("generate_sobol_sample", I * 4, lambda r, n, v: cat(u32(r, n, 1), r.randint(0, 12, size=(n, 1)).astype(np.uint32))),
    3: ("owen_scramble_2d", I * 4, lambda r, n, v: u32(r, n, 8)),
    4: ("init_local_sampler + generate_ray_sample_uint + generate_uniform_random_uint", I * 8,
        lambda r, n, v: cat(r.randint(0, 4096, size=(n, 2)).astype(np.uint32), r.randint(0, 45, size=(n, 1)).astype(np.uint32),
                            r.randint(0, 5000, size=(n, 1)).astype(np.uint32), r.randint(0, 100000, size=(n, 1)).astype(np.uint32),
                            u32(r, n, 1), r.randint(0, 20, size=(n, 1)).astype(np.uint32))),
    5: ("owen_scramble_4d", I, lambda r, n, v: u32(r, n, 2)),
    6: ("owen_scramble_8d", I, lambda r, n, v: u32(r, n, 2)),
    7: ("morton_2d", I, lambda r, n, v: u32(r, n, 2)),
    8: ("morton_3d", I, lambda r, n, v: u32(r, n, 3)),
    9: ("create_tangent_space", F * 9, lambda r, n, v: cat(direction(r, n))),
    10: ("sample_concentric_disk", F * 2, lambda r, n, v: cat(unit(r, n, 2))),
    11: ("sample_cosine_hemisphere", F * 3, lambda r, n, v: cat(unit(r, n, 2))),
    12: ("sample_cone", F * 3, lambda r, n, v: cat(unit(r, n, 2), direction(r, n), np.cos(np.radians(r.uniform(0.1, 60, size=(n, 1)))))),
    13: ("sample_regular_polygon", F * 2, lambda r, n, v: cat(unit(r, n, 2), unit(r, n, 1), r.randint(3, 9, size=(n, 1)).astype(np.uint32))),
    14: ("sample_blackman_harris_concentric_disk", F * 2, lambda r, n, v: cat(unit(r, n, 2))),
    15: ("sample_spherical_triangle", F * 4, lambda r, n, v: _tri_inputs(r, n, unit(r, n, 2))),
    16: ("spherical_triangle_solid_angle", F, lambda r, n, v: cat(direction(r, n), direction(r, n), direction(r, n))),
    17: ("sample_triangle_area", F * 3, lambda r, n, v: _tri_inputs(r, n, unit(r, n, 2))),
    18: ("triangle_area_pdf", F, lambda r, n, v: _point_on_triangle(r, n)),
    19: ("view_to_tangent_space", F * 3, None),   # needs an orthonormal basis: built in the test from function 9
    20: ("material_bsdf_sample", F * 8, lambda r, n, v: cat(unit(r, n, 4), upper(r, n), material(r, n))),
    21: ("material_bsdf_pdf", F * 5, lambda r, n, v: cat(direction(r, n), upper(r, n), material(r, n))),
    22: ("sample_point_light", F * 8, lambda r, n, v: cat(structs(point_lights(r, n)), unit(r, n, 2), r.uniform(-3, 3, size=(n, 3)) + 7.0)),
    23: ("sample_point_light_pdf", F, lambda r, n, v: cat(structs(point_lights(r, n)), r.uniform(-3, 3, size=(n, 3)) + 7.0)),
    24: ("sample_directional_light", F * 7, lambda r, n, v: cat(structs(directional_lights(r, n)), unit(r, n, 2))),
    25: ("sample_triangle_light", F * 4, lambda r, n, v: _tri_inputs(r, n, unit(r, n, 2))),
    26: ("sample_triangle_light_pdf", F, lambda r, n, v: _point_on_triangle(r, n)),
    27: ("r9g9b9e5_to_rgb", F * 3, lambda r, n, v: u32(r, n, 1)),
    28: ("rgb_to_r9g9b9e5", I, lambda r, n, v: cat(r.uniform(0, 1, size=(n, 3)) * np.exp2(r.randint(-8, 8, size=(n, 1))))),
    29: ("modulate_bsdf", F * 3, lambda r, n, v: cat(material(r, n), unit(r, n, 4))),
    30: ("modulate_color", F * 3, lambda r, n, v: cat(material(r, n), unit(r, n, 3) * 4, unit(r, n, 3) * 4)),
    31: ("get_camera_ray", F * 6, lambda r, n, v: cat(structs(cameras(r, n, VARIANTS[v].get("projection", 0) == abi.PROJECTION_ORTHOGRAPHIC)),
                                                    r.uniform(0, 1000, size=(n, 2)), np.full((n, 2), 1000.0), unit(r, n, 2))),
    32: ("octahedral_pack", F * 2, lambda r, n, v: cat(direction(r, n))),
    33: ("inverse_srgb_correction", F * 3, lambda r, n, v: cat(unit(r, n, 3))),
    34: ("rgb_to_luminance", F, lambda r, n, v: cat(unit(r, n, 3) * 10)),
    35: ("latlong_direction_to_pixel_id", I, lambda r, n, v: cat(direction(r, n), np.full((n, 1), 64, dtype=np.uint32), np.full((n, 1), 32, dtype=np.uint32))),
    36: ("uv_to_latlong_direction", F * 3, lambda r, n, v: cat(unit(r, n, 2))),
    37: ("ggx_vndf_sample", F * 3, lambda r, n, v: cat(upper(r, n), r.uniform(0.002, 1.0, size=(n, 1)), unit(r, n, 2))),
    38: ("get_spotlight_intensity", F, lambda r, n, v: cat(structs(point_lights(r, n)), direction(r, n))),
    39: ("ray_plane_intersection_dist", F, lambda r, n, v: _tri_inputs(r, n, direction(r, n))),
    40: ("get_barycentric_coords", F * 3, lambda r, n, v: _point_on_triangle(r, n)),
    41: ("sample_sphere", F * 3, lambda r, n, v: cat(unit(r, n, 2))),
    42: ("sample_hemisphere", F * 3, lambda r, n, v: cat(unit(r, n, 2))),
    43: ("sample_blackman_harris", F, lambda r, n, v: cat(unit(r, n, 1))),
}
