"""The `direct` estimator (direct_stage, shader/direct.rgen) in the oracle. The reference has no golden for it, so it
is pinned against the golden-pinned path tracer: with one sample per pass, MIS off and lights hidden from the camera,
a depth-2 path tracer draws the SAME sampler key, camera ray and next-event sample; what its extra BSDF bounce finds
gets MIS weight 0 (every light class is NEE-sampled), so the two colours agree up to the rounding of
modulate_color vs modulate_bsdf."""
import numpy as np

from tauray_b200 import abi, testscenes
from oracle.oracle_py import Oracle
from helpers import full_dist, bits_equal

SIZE = (56, 40)


def _render(opt, scene, layers=(), frames=1):
    o = Oracle(opt)
    o.set_scene(scene)
    o.set_cameras(testscenes.cornell_cameras(SIZE[0] / SIZE[1]))
    o.set_distribution(full_dist(SIZE))
    if layers:
        o.set_gbuffer_layers(list(layers))
    for _ in range(frames):
        o.render()
    return o.read_color(), {l: o.read_gbuffer_layer(l) for l in layers}, o.stats()


def test_direct_equals_depth2_path_tracer_without_mis():
    scene = testscenes.cornell()
    for sampler in (abi.SAMPLER_UNIFORM_RANDOM, abi.SAMPLER_SOBOL_OWEN):
        common = dict(samples_per_pixel=4, samples_per_pass=1, local_sampler=sampler, film=abi.FILM_POINT)
        direct, layers, st = _render(abi.default_options(estimator=abi.ESTIMATOR_DIRECT, max_ray_depth=2, **common), scene,
                                     (abi.GBUFFER_INSTANCE_ID,))
        pt, _, st_pt = _render(abi.default_options(max_ray_depth=2, mis_mode=abi.MIS_DISABLED, hide_lights=1, **common), scene)
        assert st["closest_rays"] * 2 == st_pt["closest_rays"]          # one camera ray per pass, no BSDF bounce
        assert st["shadow_rays"] == st_pt["shadow_rays"]                # the same NEE samples
        d, p = direct[..., :3].astype(np.float64), pt[..., :3].astype(np.float64)
        close = (np.abs(d - p) <= 1e-4 * np.abs(p) + 1e-6).all(axis=-1)
        # the emissive ceiling panel (instance 5) seen directly: the path tracer adds its emission twice
        # (primary_lobes = (0,0,0,1) puts it in `reflection` AND it is first_hit_material.emission),
        # direct.rgen once
        panel = layers[abi.GBUFFER_INSTANCE_ID][0, :, 0].reshape(close.shape[1:]) == 5
        assert panel.any() and close[0][~panel].all()
        assert np.allclose(p[0][panel] - d[0][panel], [12.0, 11.0, 8.0], rtol=1e-5)
        assert np.isfinite(direct).all() and direct[..., :3].mean() > 0.01


def test_direct_samples_per_pass_and_layers():
    scene = testscenes.cornell(alpha_card=False)
    layers = (abi.GBUFFER_DIFFUSE, abi.GBUFFER_REFLECTION, abi.GBUFFER_ALBEDO, abi.GBUFFER_INSTANCE_ID)
    one, l1, _ = _render(abi.default_options(estimator=abi.ESTIMATOR_DIRECT, samples_per_pixel=64, samples_per_pass=1), scene, layers)
    many, l4, st = _render(abi.default_options(estimator=abi.ESTIMATOR_DIRECT, samples_per_pixel=64, samples_per_pass=4), scene, layers)
    # 16 passes x (1 camera ray + 4 NEE samples) per pixel
    assert st["closest_rays"] == 16 * SIZE[0] * SIZE[1]
    # both are 64 next-event samples per pixel of the same integrand
    assert abs(one[..., :3].mean() - many[..., :3].mean()) < 0.03 * one[..., :3].mean()
    # first-sample layers do not depend on the estimator's sample count
    assert l1[abi.GBUFFER_INSTANCE_ID].tobytes() == l4[abi.GBUFFER_INSTANCE_ID].tobytes()
    assert l1[abi.GBUFFER_ALBEDO].tobytes() == l4[abi.GBUFFER_ALBEDO].tobytes()
    # direct.rgen divides diffuse (alpha included) by SAMPLES_PER_PASS but not reflection (kept as is):
    hit = l1[abi.GBUFFER_INSTANCE_ID][0, :, 0] >= 0
    d1 = l1[abi.GBUFFER_DIFFUSE][0].astype(np.float32)[hit]
    d4 = l4[abi.GBUFFER_DIFFUSE][0].astype(np.float32)[hit]
    r1 = l1[abi.GBUFFER_REFLECTION][0].astype(np.float32)[hit]
    r4 = l4[abi.GBUFFER_REFLECTION][0].astype(np.float32)[hit]
    assert np.allclose(d4[:, 3] * 4, d1[:, 3], rtol=2e-3)               # alpha = (1 / distance) / 4
    assert np.allclose(r4[:, 3], r1[:, 3], rtol=2e-3)                   # alpha = 1 / distance
    assert abs(r4[:, :3].mean() - 4 * r1[:, :3].mean()) < 0.12 * 4 * r1[:, :3].mean()   # reflection holds the SUM of the 4 samples
    assert abs(d4[:, :3].mean() - d1[:, :3].mean()) < 0.05 * d1[:, :3].mean()


def test_direct_accumulates_over_frames_like_the_path_tracer():
    scene = testscenes.cornell(alpha_card=False, sphere_light=False)
    opt = abi.default_options(estimator=abi.ESTIMATOR_DIRECT, samples_per_pixel=2, samples_per_pass=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    a, _, _ = _render(opt, scene, frames=1)
    b, _, _ = _render(opt, scene, frames=3)
    assert not bits_equal(a, b)
    assert abs(a[..., :3].mean() - b[..., :3].mean()) < 0.1 * a[..., :3].mean()
