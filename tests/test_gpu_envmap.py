"""GPU parity for environment-map NEE: alias table built by tr_pt_set_scene, envmap
importance sampling, miss lookup -- CUDA through the C ABI against the oracle, bit for bit."""
import numpy as np
import pytest

from tauray_b200 import abi, testscenes
from tauray_b200.library import PathTracer, TaurayB200Error
from oracle.oracle_py import Oracle

from helpers import assert_parity, full_dist, run

pytestmark = pytest.mark.gpu

SIZE = (64, 40)

CASES = {
    "uniform_random": dict(max_ray_depth=4),
    "sobol_owen_2spp": dict(max_ray_depth=4, local_sampler=abi.SAMPLER_SOBOL_OWEN, samples_per_pixel=2),
    "envmap_not_sampled": dict(max_ray_depth=3, sampling_weight_envmap=0.0),
    "only_envmap_sampled": dict(max_ray_depth=3, sampling_weight_emissive_triangles=0.0, sampling_weight_directional_lights=0.0),
    "weighted": dict(max_ray_depth=3, sampling_weight_envmap=3.0, sampling_weight_emissive_triangles=0.5),
    "mis_balance": dict(max_ray_depth=3, mis_mode=abi.MIS_BALANCE_HEURISTIC),
    "mis_disabled": dict(max_ray_depth=3, mis_mode=abi.MIS_DISABLED),
    "transparent_background": dict(max_ray_depth=2, transparent_background=1),
    "box_filter_4spp": dict(max_ray_depth=3, film=abi.FILM_BOX, samples_per_pixel=4, local_sampler=abi.SAMPLER_SOBOL_OWEN),
}


def _check(opt, scene, frames=1, modes=(abi.KERNEL_WAVEFRONT, abi.KERNEL_MEGAKERNEL)):
    cams = testscenes.courtyard_cameras(SIZE[0] / SIZE[1])
    dist = full_dist(SIZE)
    ora = Oracle(opt)
    ref = run(ora, scene, cams, dist, frames)
    ost = ora.stats()
    for mode in modes:
        o = abi.copy_options(opt)
        o.kernel_mode = mode
        with PathTracer(o) as pt:
            img = run(pt, scene, cams, dist, frames)
            st = pt.stats()
        assert st["kernel_launches"] > 0
        assert assert_parity(img, ref, "kernel_mode %d" % mode), "within 1e-5 but not bit-exact"
        assert st["closest_rays"] == ost["closest_rays"] * 1
        assert st["shadow_rays"] + st["culled_shadow_rays"] == ost["shadow_rays"]
    return ref


@pytest.mark.parametrize("name", sorted(CASES))
def test_envmap_option_matrix(name):
    scene = testscenes.courtyard(envmap=True, emissive=True, directional=True)
    ref = _check(abi.default_options(**CASES[name]), scene)
    assert np.isfinite(ref).all() and float(ref[..., :3].mean()) > 0.01


def test_envmap_only_light_source_and_accumulation_over_frames():
    scene = testscenes.courtyard(envmap=True)
    _check(abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN), scene, frames=3,
           modes=(abi.KERNEL_WAVEFRONT,))


def test_rgba32f_environment_map():
    scene = testscenes.courtyard(envmap=True)
    scene.environment_map = scene.environment_map.astype(np.float32)
    _check(abi.default_options(max_ray_depth=3), scene, modes=(abi.KERNEL_WAVEFRONT,))


@pytest.mark.parametrize("size", [(64, 32), (37, 19), (256, 128)])
def test_alias_table_is_bit_identical(size):
    scene = testscenes.courtyard(envmap=True, env_size=size)
    o = Oracle(abi.default_options())
    o.set_scene(scene)
    want = o.alias_table()
    with PathTracer(abi.default_options()) as pt:
        pt.set_scene(scene)
        got = pt.alias_table()
    assert got.shape == want.shape == (size[0] * size[1],)
    assert got.tobytes() == want.tobytes()


def test_scene_without_envmap_has_no_alias_table_and_bad_envmaps_fail_loudly():
    with PathTracer(abi.default_options()) as pt:
        pt.set_scene(testscenes.cornell())
        assert pt.alias_table().shape[0] == 0
        scene = testscenes.courtyard(envmap=True)
        d = scene.desc()
        d.environment_map = None
        with pytest.raises(TaurayB200Error):
            pt.set_scene(d)
