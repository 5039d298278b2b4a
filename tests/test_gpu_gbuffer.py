"""GPU parity of the aux G-buffer layers (write_all_outputs, shader/path_tracer.glsl:547-574): every layer byte-identical
to the oracle's, through the C ABI, for both kernel organisations."""
import numpy as np
import pytest

from tauray_b200 import abi, distribution, testscenes
from tauray_b200.library import PathTracer, TaurayB200Error
from oracle.oracle_py import Oracle

from helpers import bits_equal, full_dist

pytestmark = pytest.mark.gpu

SIZE = (64, 40)
ALL = [abi.GBUFFER_DIFFUSE, abi.GBUFFER_REFLECTION, abi.GBUFFER_ALBEDO, abi.GBUFFER_MATERIAL, abi.GBUFFER_NORMAL,
       abi.GBUFFER_POS, abi.GBUFFER_SCREEN_MOTION, abi.GBUFFER_INSTANCE_ID]
NAMES = {abi.GBUFFER_DIFFUSE: "diffuse", abi.GBUFFER_REFLECTION: "reflection", abi.GBUFFER_ALBEDO: "albedo",
         abi.GBUFFER_MATERIAL: "material", abi.GBUFFER_NORMAL: "normal", abi.GBUFFER_POS: "pos",
         abi.GBUFFER_SCREEN_MOTION: "screen_motion", abi.GBUFFER_INSTANCE_ID: "instance_id"}


def _same(a, b):
    if a.dtype.kind == "f" and a.dtype.itemsize == 4:
        return bits_equal(a, b)
    return a.tobytes() == b.tobytes()


def _run(t, scene, cams, dist, layers, frames):
    t.set_scene(scene)
    t.set_cameras(cams)
    t.set_distribution(dist)
    t.set_gbuffer_layers(layers)
    for _ in range(frames):
        t.render()
    if hasattr(t, "sync"):
        t.sync()
    return t.read_color(), {l: t.read_gbuffer_layer(l) for l in layers}


def _check(opt, scene, cams, dist=None, layers=ALL, frames=1, modes=(abi.KERNEL_WAVEFRONT, abi.KERNEL_MEGAKERNEL)):
    dist = dist or full_dist(SIZE)
    ref_col, ref = _run(Oracle(opt), scene, cams, dist, layers, frames)
    for mode in modes:
        o = abi.copy_options(opt)
        o.kernel_mode = mode
        with PathTracer(o) as pt:
            col, got = _run(pt, scene, cams, dist, layers, frames)
        assert bits_equal(col, ref_col), "colour, kernel_mode %d" % mode
        for l in layers:
            assert got[l].shape == ref[l].shape
            if not _same(got[l], ref[l]):
                bad = int((got[l] != ref[l]).any(axis=-1).sum())
                raise AssertionError("layer %s differs in %d texels (kernel_mode %d)" % (NAMES[l], bad, mode))
    return ref


CASES = {
    "one_sample": dict(max_ray_depth=4),
    "sobol_4spp_batched_passes": dict(max_ray_depth=4, samples_per_pixel=4, local_sampler=abi.SAMPLER_SOBOL_OWEN),
    "two_samples_per_pass": dict(max_ray_depth=3, samples_per_pixel=4, samples_per_pass=2, film=abi.FILM_BOX),
    "depth_1": dict(max_ray_depth=1),
    "transparent_background": dict(max_ray_depth=2, transparent_background=1),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_layers_cornell(name):
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    ref = _check(abi.default_options(**CASES[name]), testscenes.cornell(), cams)
    assert (ref[abi.GBUFFER_INSTANCE_ID] >= 0).mean() > 0.9


def test_layers_open_scene_with_sky_pixels_and_moving_instances():
    scene = testscenes.courtyard(envmap=True, emissive=True)
    # previous-frame transforms differ from the current ones -> non-trivial screen motion
    for k in range(len(scene.transforms)):
        prev = scene.transforms[k].copy()
        prev[:3, 3] += (0.05 * k, 0.02, -0.03 * k)
        scene.prev_transforms[k] = prev
    cams = testscenes.courtyard_cameras(SIZE[0] / SIZE[1])
    cams[0]["previous"]["view_proj"] = cams[0]["current"]["view_proj"] * np.float32(1.01)
    ref = _check(abi.default_options(max_ray_depth=3, samples_per_pixel=2), scene, cams, frames=2)
    inst = ref[abi.GBUFFER_INSTANCE_ID][0, :, 0]
    assert (inst < 0).any() and (inst >= 0).any()


def test_layers_on_a_non_primary_strip():
    dist = distribution.make_distribution(SIZE, abi.DISTRIBUTION_SHUFFLED_STRIPS, 500, 1200, False)
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    _check(abi.default_options(max_ray_depth=3), testscenes.cornell(), cams, dist=dist, modes=(abi.KERNEL_WAVEFRONT,))


def test_layer_subset_and_api_errors():
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    _check(abi.default_options(max_ray_depth=2), testscenes.cornell(), cams, layers=[abi.GBUFFER_NORMAL, abi.GBUFFER_DIFFUSE],
           modes=(abi.KERNEL_WAVEFRONT,))
    with PathTracer(abi.default_options()) as pt:
        with pytest.raises(TaurayB200Error):
            pt.set_gbuffer_layers([12])
        pt.set_scene(testscenes.cornell())
        pt.set_cameras(cams)
        pt.set_distribution(full_dist(SIZE))
        with pytest.raises(TaurayB200Error):
            pt.read_gbuffer_layer(abi.GBUFFER_POS)       # not enabled
