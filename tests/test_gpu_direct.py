"""GPU parity of the `direct` estimator (direct_stage, shader/direct.rgen) through the C ABI: colour and light layers
bit-identical to the oracle's over an option matrix."""
import numpy as np
import pytest

from tauray_b200 import abi, distribution, testscenes
from tauray_b200.library import PathTracer
from oracle.oracle_py import Oracle

from helpers import assert_parity, bits_equal, full_dist

pytestmark = pytest.mark.gpu

SIZE = (80, 56)
LAYERS = [abi.GBUFFER_DIFFUSE, abi.GBUFFER_REFLECTION, abi.GBUFFER_ALBEDO, abi.GBUFFER_NORMAL, abi.GBUFFER_POS,
          abi.GBUFFER_INSTANCE_ID]

CASES = {
    "defaults_1spp": dict(),
    "sobol_owen_4_per_pass": dict(samples_per_pixel=8, samples_per_pass=4, local_sampler=abi.SAMPLER_SOBOL_OWEN),
    "sobol_z3_3_per_pass_box": dict(samples_per_pixel=6, samples_per_pass=3, local_sampler=abi.SAMPLER_SOBOL_Z_ORDER_3D,
                                    film=abi.FILM_BOX),
    "blackman_harris_dof_area_lights": dict(samples_per_pixel=4, samples_per_pass=2, film=abi.FILM_BLACKMAN_HARRIS,
                                            film_radius=1.0, depth_of_field=1, tri_light_mode=abi.TRI_LIGHT_AREA),
    "hybrid_tri_lights_cosine_pdf": dict(samples_per_pixel=2, samples_per_pass=2, tri_light_mode=abi.TRI_LIGHT_HYBRID,
                                         bounce_mode=abi.BOUNCE_COSINE_HEMISPHERE),
    "only_point_lights": dict(samples_per_pixel=2, sampling_weight_emissive_triangles=0.0, sampling_weight_directional_lights=0.0),
    "mis_option_is_ignored": dict(samples_per_pixel=2, mis_mode=abi.MIS_BALANCE_HEURISTIC, hide_lights=1, max_ray_depth=3),
    "transparent_background": dict(samples_per_pixel=2, transparent_background=1),
}


def _run(t, scene, cams, dist, frames):
    t.set_scene(scene)
    t.set_cameras(cams)
    t.set_distribution(dist)
    t.set_gbuffer_layers(LAYERS)
    for _ in range(frames):
        t.render()
    if hasattr(t, "sync"):
        t.sync()
    return t.read_color(), {l: t.read_gbuffer_layer(l) for l in LAYERS}, t.stats()


@pytest.mark.parametrize("name", sorted(CASES))
def test_direct_bit_exact(name):
    opt = abi.default_options(estimator=abi.ESTIMATOR_DIRECT, **CASES[name])
    scene = testscenes.cornell()
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    if CASES[name].get("depth_of_field"):
        cams[0]["current"]["dof_params"] = (5.0, 0.05, 0.3, 6.0)
    dist = full_dist(SIZE)
    want, want_layers, so = _run(Oracle(opt), scene, cams, dist, 2)
    with PathTracer(opt) as pt:
        got, got_layers, sg = _run(pt, scene, cams, dist, 2)
    assert_parity(got, want, name)
    assert bits_equal(got, want), name
    for l in LAYERS:
        assert got_layers[l].tobytes() == want_layers[l].tobytes(), (name, l)
    assert sg["closest_rays"] == so["closest_rays"] and sg["shadow_rays"] == so["shadow_rays"]
    assert sg["kernel_launches"] > 0


def test_direct_open_scene_with_environment_map_and_strips():
    """Sky pixels (terminal camera rays), env-map NEE, and a non-primary strip of a 3-way split."""
    opt = abi.default_options(estimator=abi.ESTIMATOR_DIRECT, samples_per_pixel=4, samples_per_pass=2,
                              local_sampler=abi.SAMPLER_SOBOL_OWEN)
    scene = testscenes.courtyard(envmap=True, emissive=True)
    cams = testscenes.courtyard_cameras(SIZE[0] / SIZE[1])
    per = distribution.calculate_shuffled_strips_pixels_per_device(SIZE, 1.0 / 3.0)
    for dist in (full_dist(SIZE), distribution.make_distribution(SIZE, abi.DISTRIBUTION_SHUFFLED_STRIPS, per, per, False)):
        ref = Oracle(opt)
        ref.set_scene(scene); ref.set_cameras(cams); ref.set_distribution(dist); ref.render()
        with PathTracer(opt) as pt:
            pt.set_scene(scene); pt.set_cameras(cams); pt.set_distribution(dist); pt.render(); pt.sync()
            assert bits_equal(pt.read_color(), ref.read_color())
