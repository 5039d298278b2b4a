"""GPU parity on the BASELINE configs: test.glb (C2, against the reference goldens AND the oracle), the synthetic
atrium (C3) at sizes the oracle finishes in seconds, and size-independent properties at BASELINE's full size."""
import os

import cv2
import numpy as np
import pytest

from tauray_b200 import abi, distribution, synthetic, testscenes
from tauray_b200.library import PathTracer
from tauray_b200.scene import camera_pairs, Scene, make_material
from oracle.oracle_py import Oracle

from helpers import assert_parity, bits_equal, full_dist, run
from test_oracle_golden import FEATURES, imagemagick_mse, read_exr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def glb_gpu(glb_scene):
    pt = PathTracer(abi.default_options())
    pt.set_scene(glb_scene.scene)
    pt.set_cameras(camera_pairs([glb_scene.camera_data(0, aspect=1.0)]))
    pt.set_distribution(distribution.make_distribution((512, 512), abi.DISTRIBUTION_DUPLICATE))
    yield pt
    pt.close()


@pytest.mark.parametrize("name", sorted(FEATURES))
def test_test_glb_feature_goldens_on_gpu(glb_gpu, glb_scene, golden_dir, name):
    ref = read_exr(os.path.join(golden_dir, "references", "validate_%s.exr" % name))
    glb_gpu.render_feature(FEATURES[name], (0, 0, 0, 0))
    glb_gpu.sync()
    img = glb_gpu.read_color()
    assert imagemagick_mse(img[0, :, :, :3], ref) < 1.0        # the reference's criterion (test/CMakeLists.txt:32-37)
    o = Oracle(abi.default_options())
    o.set_scene(glb_scene.scene)
    o.set_cameras(camera_pairs([glb_scene.camera_data(0, aspect=1.0)]))
    o.set_distribution(distribution.make_distribution((512, 512), abi.DISTRIBUTION_DUPLICATE))
    o.render_feature(FEATURES[name], (0, 0, 0, 0))
    assert assert_parity(img, o.read_color(), name)


def test_config2_test_glb_512_1spp_2bounce(glb_scene):
    """BASELINE configs[1]: test.glb path tracer 512x512, 1 spp, 2 bounces (textures, alpha any-hit, transmission,
    sphere-light intersection, tri-light / spot / directional NEE)."""
    cams = camera_pairs([glb_scene.camera_data(0, aspect=1.0)])
    for sampler in (abi.SAMPLER_SOBOL_OWEN, abi.SAMPLER_UNIFORM_RANDOM):
        opt = abi.default_options(max_ray_depth=2, samples_per_pixel=1, local_sampler=sampler)
        ref = run(Oracle(opt), glb_scene.scene, cams, full_dist((512, 512)))
        for mode in (abi.KERNEL_WAVEFRONT, abi.KERNEL_MEGAKERNEL):
            o = abi.copy_options(opt)
            o.kernel_mode = mode
            with PathTracer(o) as pt:
                img = run(pt, glb_scene.scene, cams, full_dist((512, 512)))
            assert assert_parity(img, ref, "test.glb sampler %d mode %d" % (sampler, mode))


def test_test_glb_depth8_cli_defaults(glb_scene):
    cams = camera_pairs([glb_scene.camera_data(0, aspect=1.0)])
    opt = abi.default_options(samples_per_pixel=2)
    ref = run(Oracle(opt), glb_scene.scene, cams, full_dist((256, 256)))
    with PathTracer(opt) as pt:
        img = run(pt, glb_scene.scene, cams, full_dist((256, 256)))
    assert assert_parity(img, ref, "test.glb depth 8")


@pytest.fixture(scope="module")
def atrium_scene():
    return synthetic.atrium()


def bench_opt(**kw):
    base = dict(max_ray_depth=4, samples_per_pixel=4, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    base.update(kw)
    return abi.default_options(**base)


def test_config3_atrium_reduced_resolution_vs_oracle(atrium_scene):
    """The full 1.5 M-triangle scene and the bench options, at 240x135 so the oracle finishes in seconds."""
    size = (240, 135)
    cams = synthetic.atrium_cameras(16 / 9)
    opt = bench_opt()
    o = Oracle(opt)
    ref = run(o, atrium_scene, cams, full_dist(size))
    so = o.stats()
    with PathTracer(opt) as pt:
        img = run(pt, atrium_scene, cams, full_dist(size))
        sg = pt.stats()
    assert assert_parity(img, ref, "atrium 240x135")
    assert (sg["closest_rays"], sg["shadow_rays"] + sg["culled_shadow_rays"], sg["paths"]) == (so["closest_rays"], so["shadow_rays"], so["paths"])
    assert sg["culled_shadow_rays"] > 0     # zero-weight NEE samples are not traced
    assert sg["triangle_count"] == 1_500_000 and sg["tri_light_count"] == 64


def test_config3_full_size_properties(atrium_scene):
    """1920x1080, 4 spp, 4 bounces: (1) wavefront == megakernel bit for bit, (2) the 8-way shuffled-strips split renders
    exactly the single-device image (strips + un-shuffle on device), (3) no NaN, rays/path within the bound."""
    size = (1920, 1080)
    cams = synthetic.atrium_cameras(16 / 9)
    opt = bench_opt()
    with PathTracer(opt) as pt:
        full = run(pt, atrium_scene, cams, full_dist(size))
        st = pt.stats()
        assert not np.isnan(full).any()
        assert st["paths"] == 1920 * 1080 * 4
        assert st["paths"] < st["closest_rays"] <= 4 * st["paths"] and st["shadow_rays"] + st["culled_shadow_rays"] <= 3 * st["paths"]
        # (2) strips: 8 "devices" rendered one after the other on this GPU, gathered by tr_pt_unshuffle
        import torch
        stitched = torch.zeros((1, 1080, 1920, 4), dtype=torch.float32, device="cuda")
        dists = distribution.device_distributions(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, distribution.even_workloads(8))
        for d in dists:
            pt.set_distribution(d)
            pt.reset_accumulation()
            pt.render()
            pt.sync()
            ptr, nbytes = pt.color_buffer()
            assert nbytes == d.count * 16
            pt.unshuffle(ptr, d.index, d.count, stitched.data_ptr())
            pt.sync()
        assert bits_equal(stitched.cpu().numpy(), full)
    mk = bench_opt(kernel_mode=abi.KERNEL_MEGAKERNEL)
    with PathTracer(mk) as pt:
        assert bits_equal(run(pt, atrium_scene, cams, full_dist(size)), full)


def test_config3_full_size_vs_oracle(atrium_scene):
    """BASELINE configs[2] at its full size -- 1920x1080, 4 spp, 4 bounces, 1.5 M triangles -- against the oracle, bit for
    bit (8.3 M paths: a few seconds on the GPU box's host cores), ray counts included."""
    size = (1920, 1080)
    cams = synthetic.atrium_cameras(16 / 9)
    opt = bench_opt()
    o = Oracle(opt)
    ref = run(o, atrium_scene, cams, full_dist(size))
    so = o.stats()
    with PathTracer(opt) as pt:
        img = run(pt, atrium_scene, cams, full_dist(size))
        sg = pt.stats()
    assert assert_parity(img, ref, "atrium 1920x1080 4 spp")
    assert (sg["closest_rays"], sg["shadow_rays"] + sg["culled_shadow_rays"], sg["paths"]) == (so["closest_rays"], so["shadow_rays"], so["paths"])


def test_config4_long_accumulation_vs_oracle(atrium_scene):
    """BASELINE configs[3]: 4096 spp accumulated as 1024 frames of 4 spp. Against the oracle at 48x27 (5.3 M paths), frame
    by frame running mean, sample counter running."""
    size = (48, 27)
    cams = synthetic.atrium_cameras(16 / 9)
    opt = bench_opt()
    frames = 1024
    ref = run(Oracle(opt), atrium_scene, cams, full_dist(size), frames=frames)
    with PathTracer(opt) as pt:
        img = run(pt, atrium_scene, cams, full_dist(size), frames=frames)
        ck = pt.checkpoint()
    assert ck["accumulated_samples"] == 4096 and ck["frame_counter"] == frames
    assert assert_parity(img, ref, "atrium 48x27 4096 spp")


def test_config4_full_size_accumulation_resumes_bit_exactly(atrium_scene):
    """configs[3] at 1920x1080: 64 accumulated frames (256 spp) straight through == 40 frames, checkpoint, a fresh context,
    restore, 24 more frames; and the running mean converges (frame-to-frame change shrinks)."""
    size = (1920, 1080)
    cams = synthetic.atrium_cameras(16 / 9)
    opt = bench_opt()
    dist = full_dist(size)
    with PathTracer(opt) as pt:
        pt.set_scene(atrium_scene); pt.set_cameras(cams); pt.set_distribution(dist)
        for _ in range(8):
            pt.render()
        pt.sync()
        at8 = pt.read_color().copy()
        for _ in range(32):
            pt.render()
        pt.sync()
        ck = pt.checkpoint()
        at40 = ck["color"]
        for _ in range(24):
            pt.render()
        pt.sync()
        want = pt.read_color().copy()
    with PathTracer(opt) as pt:
        pt.set_scene(atrium_scene); pt.set_cameras(cams); pt.set_distribution(dist)
        pt.restore(ck)
        for _ in range(24):
            pt.render()
        pt.sync()
        got = pt.read_color()
    assert bits_equal(got, want)
    d1 = np.abs(at40[..., :3] - at8[..., :3]).mean()
    d2 = np.abs(want[..., :3] - at40[..., :3]).mean()
    assert d2 < d1 and np.isfinite(want).all()


def test_config5_full_size_vs_oracle(atrium_scene):
    """BASELINE configs[4] at its full size: 5x9 camera grid = 45 views of 960x540, 2 spp, 4 bounces (46.7 M paths: the pass
    batch is bounded by the 16 Mi-slot budget, the view index enters the sampler key). Bit for bit against the oracle."""
    size = (960, 540)
    cams = synthetic.atrium_camera_grid(16 / 9)
    assert len(cams) == 45
    opt = bench_opt(samples_per_pixel=2)
    ref = run(Oracle(opt), atrium_scene, cams, full_dist(size))
    with PathTracer(opt) as pt:
        img = run(pt, atrium_scene, cams, full_dist(size))
    assert img.shape == (45, 540, 960, 4)
    assert assert_parity(img, ref, "light field 45 x 960x540")
    assert not bits_equal(img[0], img[44])


def test_fused_peer_stitch_on_one_gpu(atrium_scene):
    """tr_pt_set_peer_images: two contexts (two 'ranks') store their strips straight into one full image."""
    size = (640, 360)
    cams = synthetic.atrium_cameras(16 / 9)
    opt = bench_opt(samples_per_pixel=2)
    with PathTracer(opt) as single:
        want = run(single, atrium_scene, cams, full_dist(size))
    dists = distribution.device_distributions(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, [0.6, 0.4])
    a, b = PathTracer(opt), PathTracer(opt)
    try:
        for pt, d in ((a, dists[0]), (b, dists[1])):
            pt.set_scene(atrium_scene)
            pt.set_cameras(cams)
            pt.set_distribution(d)
        full_ptr, nbytes = a.alloc_full_image()
        assert nbytes == 640 * 360 * 16
        a.set_peer_images([full_ptr])
        b.set_peer_images([full_ptr])
        a.render(); b.render()
        a.sync(); b.sync()
        got = a.read_full_image(np.empty((1, 360, 640, 4), dtype=np.float32))
        assert bits_equal(got, want)
    finally:
        a.close(); b.close()


def test_config5_camera_grid_small(atrium_scene):
    """BASELINE configs[4] in miniature: a 5x9 camera grid (45 views), 2 spp, 4 bounces."""
    size = (96, 54)
    cams = synthetic.atrium_camera_grid(16 / 9)
    assert len(cams) == 45
    opt = bench_opt(samples_per_pixel=2)
    small = synthetic.atrium(20000)
    ref = run(Oracle(opt), small, cams, full_dist(size))
    with PathTracer(opt) as pt:
        img = run(pt, small, cams, full_dist(size))
    assert img.shape == (45, 54, 96, 4)
    assert assert_parity(img, ref, "camera grid")


def test_edge_cases():
    opt = abi.default_options(max_ray_depth=3, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    cams = testscenes.cornell_cameras(1.0)
    dist = full_dist((32, 32))
    # empty scene: every ray misses; a directional light makes the miss shader do work
    empty = Scene()
    empty.add_directional_light((1, 2, 3), (0, 0, -1), angle_deg=30.0)
    ref = run(Oracle(opt), empty, cams, dist)
    with PathTracer(opt) as pt:
        assert assert_parity(run(pt, empty, cams, dist), ref, "empty scene")
    # one triangle, degenerate (zero-area) triangles, a triangle far from the origin
    for verts in ([(-1, -1, 0), (1, -1, 0), (0, 1, 0)],
                  [(-1, -1, 0), (1, -1, 0), (0, 1, 0), (0.5, 0.5, -1), (0.5, 0.5, -1), (0.5, 0.5, -1), (0, 0, -2), (1, 1, -2), (2, 2, -2)],
                  [(-1000, -1000, -3000), (1000, -1000, -3000), (0, 1000, -3000)]):
        s = Scene()
        v = np.zeros(len(verts), dtype=abi.VERTEX)
        v["pos"] = verts
        v["normal"] = (0, 0, 1)
        s.add_instance(v, np.arange(len(verts), dtype=np.uint32), make_material(albedo=(0.7, 0.6, 0.5, 1), emission=(1, 1, 1)))
        s.add_point_light((5, 5, 5), (0, 0, 2), radius=0.0)
        ref = run(Oracle(opt), s, cams, dist)
        with PathTracer(opt) as pt:
            assert assert_parity(run(pt, s, cams, dist), ref, "tiny scene %d" % len(verts))


def test_cuda_frames_against_the_reference_pipeline_images():
    """The CUDA path against frames rendered by the reference's OWN shader code (oracle/_ref: path_tracer.rgen and its
    hit shaders compiled for the CPU; committed by tools/make_ref_fixtures.py as tests/golden/ref_glsl_images.npz): the
    chain reference text -> oracle -> CUDA closed on the GPU box. Criterion as in tests/test_ref_glsl.py (a path tracer
    is chaotic: >= 98.5 % of the pixels within 1e-4 relative, image means within 1e-3)."""
    import test_ref_glsl as T
    from ref_compare import compare_images
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_glsl_images.npz"))
    for name in sorted(T.IMAGE_CASES):
        variant, scene, cams, size, opt = T.image_case(name)
        with PathTracer(opt) as pt:
            img = run(pt, scene, cams, distribution.make_distribution(size, abi.DISTRIBUTION_DUPLICATE), frames=2)
        close, mean_rel = compare_images(fx[name], img, name)
        min_close, max_mean = T.IMAGE_LIMITS.get(name, T.DEFAULT_LIMITS)
        assert close >= min_close and mean_rel <= max_mean, (name, close, mean_rel)
