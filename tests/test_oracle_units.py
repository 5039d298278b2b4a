"""Unit checks of the oracle's building blocks: TR math accuracy, PCG / Sobol / Owen known answers computed
independently in numpy, the Sobol table against the reference's, packing helpers, the committed known-answer vectors."""
import os

import numpy as np
import pytest

from tauray_b200 import abi, distribution, testscenes
from oracle.oracle_py import Oracle, math_op

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ulp_err(got, want):
    want32 = want.astype(np.float32)
    ulp = np.spacing(np.abs(want32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - want) / np.maximum(ulp, 1e-45)


def test_tr_math_accuracy(oracle_lib):
    rng = np.random.RandomState(0)
    x = rng.uniform(-7, 7, 20000).astype(np.float32)
    assert ulp_err(math_op(0, x), np.sin(x.astype(np.float64))).max() < 2.5 or np.abs(math_op(0, x) - np.sin(x.astype(np.float64))).max() < 2e-7
    assert np.abs(math_op(1, x) - np.cos(x.astype(np.float64))).max() < 2e-7
    y = rng.uniform(-3, 3, 20000).astype(np.float32)
    assert np.abs(math_op(2, y, x) - np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() < 4e-7
    p = rng.uniform(1e-3, 4, 20000).astype(np.float32)
    e = rng.uniform(0.2, 5, 20000).astype(np.float32)
    want = np.power(p.astype(np.float64), e.astype(np.float64))
    assert (np.abs(math_op(3, p, e) - want) / want).max() < 5e-6
    assert np.abs(math_op(4, p) - np.log2(p.astype(np.float64))).max() < 3e-7 * 12
    z = rng.uniform(-20, 20, 20000).astype(np.float32)
    assert (np.abs(math_op(5, z) - np.exp2(z.astype(np.float64))) / np.exp2(z.astype(np.float64))).max() < 3e-7
    u = rng.uniform(-1, 1, 20000).astype(np.float32)
    assert np.abs(math_op(6, u) - np.arcsin(u.astype(np.float64))).max() < 6e-7
    assert np.abs(math_op(7, u) - np.arccos(u.astype(np.float64))).max() < 6e-7
    # exact special values the shaders rely on
    assert math_op(0, np.float32([0.0]))[0] == 0.0 and math_op(1, np.float32([0.0]))[0] == 1.0
    assert math_op(3, np.float32([0.0]), np.float32([2.4]))[0] == 0.0


# ---- independent numpy restatement of the integer samplers (shader/math.glsl:75-179) ------------
def np_pcg(seed):
    seed = (seed * np.uint32(747796405) + np.uint32(2891336453)).astype(np.uint32)
    seed = (((seed >> ((seed >> np.uint32(28)) + np.uint32(4))) ^ seed) * np.uint32(277803737)).astype(np.uint32)
    return ((seed >> np.uint32(22)) ^ seed).astype(np.uint32)


def np_pcg4d(s):
    s = (s * np.uint32(1664525) + np.uint32(1013904223)).astype(np.uint32)
    def mix(s):
        yzxy = s[:, [1, 2, 0, 1]]
        wxyz = s[:, [3, 0, 1, 2]]
        return (s + yzxy * wxyz).astype(np.uint32)
    s = mix(s)
    s = (s ^ (s >> np.uint32(16))).astype(np.uint32)
    return mix(s)


def np_bitrev(x):
    x = x.astype(np.uint32)
    x = ((x >> 1) & 0x55555555) | ((x & 0x55555555) << 1)
    x = ((x >> 2) & 0x33333333) | ((x & 0x33333333) << 2)
    x = ((x >> 4) & 0x0F0F0F0F) | ((x & 0x0F0F0F0F) << 4)
    x = ((x >> 8) & 0x00FF00FF) | ((x & 0x00FF00FF) << 8)
    return (((x >> 16) | (x << 16)) & 0xFFFFFFFF).astype(np.uint32)


def np_owen(x, seed):
    x = np_bitrev(x)
    x = (x ^ (x * np.uint32(0x3D20ADEA))).astype(np.uint32)
    x = (x + seed).astype(np.uint32)
    x = (x * ((seed >> np.uint32(16)) | np.uint32(1))).astype(np.uint32)
    x = (x ^ (x * np.uint32(0x05526C56))).astype(np.uint32)
    x = (x ^ (x * np.uint32(0x53A22864))).astype(np.uint32)
    return np_bitrev(x)


def np_sobol(index, bounce, table, max_sobol):
    out = np.zeros((index.shape[0], 4), dtype=np.uint32)
    for k, idx in enumerate(index):
        idx = int(idx)
        if bounce >= max_sobol:
            out[k] = np_pcg4d(np.array([[idx, bounce, (bounce * idx) & 0xFFFFFFFF, 0]], dtype=np.uint32))[0]
            continue
        if idx == 0:
            continue
        lsb = (idx & -idx).bit_length() - 1
        msb = idx.bit_length() - 1
        acc = np.zeros(4, dtype=np.uint32)
        for bit in range(lsb, msb):              # stops BEFORE the MSB, as math.glsl:137 does
            if (idx >> bit) & 1:
                acc ^= table[bounce * 32 + bit]
        out[k] = acc
    return out


def test_sampler_matches_independent_numpy(oracle_lib, golden_dir):
    table = np.load(os.path.join(golden_dir, "sobol_table_ref.npy"))
    coords = np.random.RandomState(3).randint(0, 3000, size=(200, 4)).astype(np.uint32)
    depth = 4
    o = Oracle(abi.default_options(max_ray_depth=depth, local_sampler=abi.SAMPLER_SOBOL_OWEN))
    with np.errstate(over="ignore"):
        # random_sampler.glsl:11-19 and the film draw
        rs = coords.copy()
        rs[:, 1] ^= np_pcg(rs[:, 0]); rs[:, 0] = np_pcg(rs[:, 0])
        rs[:, 2] ^= np_pcg(rs[:, 1]); rs[:, 1] = np_pcg(rs[:, 1])
        rs[:, 3] ^= np_pcg(rs[:, 2]); rs[:, 2] = np_pcg(rs[:, 2])
        assert (o.sample_test(coords, -1) == np_pcg4d(rs)).all()
        # sobol_owen_sampler.glsl:11-29
        ss = np_pcg4d(coords.copy())
        for bounce in (0, 1, 3, 4, 7):
            seed = ss.copy()
            index = seed[:, 3].copy()
            seed[:, 3] = bounce
            seed = np_pcg4d(seed)                       # mutated in place in the shader
            shuffled = np_owen(index, seed[:, 0])
            seed[:, 3] = index
            sob = np_sobol(shuffled, bounce, table, min(depth, 8))
            want = np.stack([np_owen(sob[:, 0], seed[:, 1]), np_owen(sob[:, 1], seed[:, 2]),
                             np_owen(sob[:, 2], seed[:, 3]), np_owen(sob[:, 3], seed[:, 0])], axis=1)
            assert (o.sample_test(coords, bounce) == want).all(), bounce


# ---- sobol-z samplers: pure-Python restatement of sobol_z_sampler.glsl + math.glsl:85-93,155-203,260-278 -----
M32 = 0xFFFFFFFF


def py_pcg2d_x(a, b):
    a = (a * 1664525 + 1013904223) & M32
    b = (b * 1664525 + 1013904223) & M32
    a, b = (a + b * 1664525) & M32, (b + a * 1664525) & M32
    a ^= a >> 16
    b ^= b >> 16
    a = (a + b * 1664525) & M32
    return a ^ (a >> 16)


def py_permutation(n, permutation, dimension):
    """math.glsl:155-166 -- cross-checked below against an explicit Lehmer-code decoder."""
    res = 0
    for i in range(n - 1, -1, -1):
        q = permutation % (n - i)
        permutation //= (n - i)
        if i == dimension:
            res = q
        if dimension > i:
            res += 1 if res >= q else 0
    return res


def py_owen_nd(x, seed, bits):
    digit = (1 << bits) - 1
    count = 24 if bits == 2 else 40320
    out = 0
    for i in range(0, 32, bits):
        prefix = x & ((~digit << i) & M32)
        out |= (py_permutation(1 << bits, py_pcg2d_x(seed, prefix) % count, (x >> i) & digit) << i) & M32
    return out


def py_morton(coords):
    out = 0
    for bit in range(16 if len(coords) == 2 else 10):
        for k, c in enumerate(coords):
            out |= ((c >> bit) & 1) << (bit * len(coords) + k)
    return out & M32


def test_get_permutation_n_enumerates_permutations():
    import math
    for n in (4, 8):
        seen = set()
        for p in range(math.factorial(n) if n == 4 else 500):
            perm = tuple(py_permutation(n, p, d) for d in range(n))
            assert sorted(perm) == list(range(n))
            seen.add(perm)
        assert len(seen) == (24 if n == 4 else 500)


@pytest.mark.parametrize("sampler,dims", [(abi.SAMPLER_SOBOL_Z_ORDER_2D, 2), (abi.SAMPLER_SOBOL_Z_ORDER_3D, 3)])
def test_sobol_z_sampler_matches_independent_python(oracle_lib, golden_dir, sampler, dims):
    table = np.load(os.path.join(golden_dir, "sobol_table_ref.npy"))
    rng = np.random.RandomState(9)
    coords = rng.randint(0, 3000, size=(120, 4)).astype(np.uint32)
    coords[:8, 3] = [0, 1, 1023, 1024, 1025, 4096, 70000, 2 ** 31 + 5]      # around the 1024-sample re-scramble
    depth = 4
    seed_opt = 77
    o = Oracle(abi.default_options(max_ray_depth=depth, local_sampler=sampler, rng_seed=seed_opt))
    # the Z samplers ignore coord.z (view + rng_seed): only x, y and the path index enter
    index = []
    for x, y, _, w in coords.tolist():
        if dims == 3:
            index.append(py_owen_nd(py_morton((x & 1023, y & 1023, w & 1023)), w >> 10, 3))
        else:
            index.append(py_owen_nd(py_morton((x & 0xFFFF, y & 0xFFFF)), w, 2))
    index = np.asarray(index, dtype=np.uint32)
    with np.errstate(over="ignore"):
        for bounce in (0, 1, 3, 4, 9):
            want = np_sobol(index, bounce, table, min(depth, 8))
            assert (o.sample_test(coords, bounce) == want).all(), bounce
    other_view = coords.copy()
    other_view[:, 2] += 5
    assert (o.sample_test(coords, 1) == o.sample_test(other_view, 1)).all()


def test_sobol_table_is_the_reference_table(golden_dir):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_sobol_table
    mine = np.array(gen_sobol_table.table(), dtype=np.uint32)
    assert (mine == np.load(os.path.join(golden_dir, "sobol_table_ref.npy"))).all()
    inc = open(os.path.join(ROOT, "include", "tr_sobol_table.inc")).read()
    import re
    vals = np.array([int(x, 16) for x in re.findall(r"0x([0-9A-F]{8})u", inc)], dtype=np.uint32).reshape(256, 4)
    assert (vals == mine).all()
    if os.path.exists("/root/reference/shader/sobol_lookup_table.glsl"):
        src = open("/root/reference/shader/sobol_lookup_table.glsl").read()
        ref = re.findall(r"uvec4\(0x([0-9A-Fa-f]+)u, 0x([0-9A-Fa-f]+)u, 0x([0-9A-Fa-f]+)u, 0x([0-9A-Fa-f]+)u\)", src)
        assert (np.array([[int(x, 16) for x in r] for r in ref], dtype=np.uint32) == mine).all()


def test_committed_known_answers(oracle_lib, golden_dir):
    kat = np.load(os.path.join(golden_dir, "oracle_kat.npz"))
    coords = kat["coords"]
    for sampler, tag in ((abi.SAMPLER_SOBOL_OWEN, "sobol"), (abi.SAMPLER_UNIFORM_RANDOM, "random")):
        o = Oracle(abi.default_options(max_ray_depth=4, local_sampler=sampler))
        for d in (-1, 0, 1, 2, 7, 9):
            assert (o.sample_test(coords, d) == kat["%s_draw_%d" % (tag, d)]).all()
    o = Oracle(abi.default_options(max_ray_depth=4, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN))
    o.set_scene(testscenes.cornell())
    o.set_cameras(testscenes.cornell_cameras(1.0))
    o.set_distribution(distribution.make_distribution((32, 32), abi.DISTRIBUTION_DUPLICATE))
    o.render()
    assert o.read_color().tobytes() == kat["cornell_32x32_2spp"].tobytes()
    assert o.tri_lights().view(np.uint8).tobytes() == kat["cornell_tri_lights"].tobytes()


def test_tri_light_extraction(oracle_lib):
    """extract_tri_lights.comp: world-space corners, RGB9E5 emission, ids."""
    sc = testscenes.cornell()
    o = Oracle(abi.default_options())
    o.set_scene(sc)
    tl = o.tri_lights()
    inst = sc.instance_table()
    emissive = [i for i in range(len(sc.meshes)) if inst["light_base_id"][i] >= 0]
    assert len(emissive) == 1 and tl.shape[0] == sc.meshes[emissive[0]][1].size // 3
    v, idx = sc.meshes[emissive[0]]
    assert np.allclose(tl["pos"][0], v["pos"][idx[:3]])
    assert (tl["instance_id"] == emissive[0]).all() and (tl["primitive_id"] == np.arange(tl.shape[0])).all()
    rgbe = int(tl["emission_factor"][0])
    e = (rgbe >> 27) - 16
    rgb = np.array([rgbe & 511, (rgbe >> 9) & 511, (rgbe >> 18) & 511]) / 512.0 * 2.0 ** e
    assert np.allclose(rgb, (12.0, 11.0, 8.0), rtol=1 / 256)


def test_estimator_invariants(oracle_lib):
    """No reference test pins the estimator; check what must hold: MIS modes and NEE on/off agree in the mean,
    nothing is NaN, accumulation over frames equals the running mean."""
    sc = testscenes.cornell(alpha_card=False, transmissive=False)
    cams = testscenes.cornell_cameras(1.0)
    size = (48, 48)
    means = {}
    for name, kw in {
        "power": dict(mis_mode=abi.MIS_POWER_HEURISTIC), "balance": dict(mis_mode=abi.MIS_BALANCE_HEURISTIC),
        "cos_hemisphere": dict(bounce_mode=abi.BOUNCE_COSINE_HEMISPHERE),
        "area_tri": dict(tri_light_mode=abi.TRI_LIGHT_AREA), "hybrid_tri": dict(tri_light_mode=abi.TRI_LIGHT_HYBRID),
        "uniform_random": dict(local_sampler=abi.SAMPLER_UNIFORM_RANDOM),
        "sobol_z2": dict(local_sampler=abi.SAMPLER_SOBOL_Z_ORDER_2D), "sobol_z3": dict(local_sampler=abi.SAMPLER_SOBOL_Z_ORDER_3D),
    }.items():
        kw = dict(dict(local_sampler=abi.SAMPLER_SOBOL_OWEN), **kw)
        o = Oracle(abi.default_options(max_ray_depth=3, samples_per_pixel=96, **kw))
        o.set_scene(sc); o.set_cameras(cams)
        o.set_distribution(distribution.make_distribution(size, abi.DISTRIBUTION_DUPLICATE))
        o.render()
        img = o.read_color()
        assert not np.isnan(img).any(), name
        means[name] = img[..., :3].mean()
    base = means["power"]
    for name, m in means.items():
        assert abs(m - base) / base < 0.06, (name, m, base)
    # running mean over frames (gbuffer.glsl:18-28): 2 frames x 2 spp == mean of the two 2-spp frames' samples
    o = Oracle(abi.default_options(max_ray_depth=3, samples_per_pixel=2))
    o.set_scene(sc); o.set_cameras(cams)
    o.set_distribution(distribution.make_distribution(size, abi.DISTRIBUTION_DUPLICATE))
    o.render(); a = o.read_color().copy()
    o.render(); ab = o.read_color().copy()
    o2 = Oracle(abi.default_options(max_ray_depth=3, samples_per_pixel=2))
    o2.set_scene(sc); o2.set_cameras(cams)
    o2.set_distribution(distribution.make_distribution(size, abi.DISTRIBUTION_DUPLICATE))
    o2.render(); o2.reset_accumulation(False); o2.render(); b = o2.read_color()
    assert np.allclose(ab, 0.5 * (a + b), rtol=2e-5, atol=1e-6)
