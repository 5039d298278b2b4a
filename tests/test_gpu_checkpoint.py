"""Accumulation checkpoint / resume (tr_pt_get_accumulation_state + tr_pt_restore_accumulation): a run interrupted after
two frames and resumed in a fresh context must end bit-identical to the uninterrupted three-frame accumulation."""
import numpy as np
import pytest

from tauray_b200 import abi, distribution, testscenes
from tauray_b200.library import PathTracer, TaurayB200Error

from helpers import bits_equal, full_dist

pytestmark = pytest.mark.gpu

SIZE = (72, 48)


def _setup(pt, scene, cams, dist):
    pt.set_scene(scene)
    pt.set_cameras(cams)
    pt.set_distribution(dist)


@pytest.mark.parametrize("strip", [False, True])
def test_resume_equals_an_uninterrupted_accumulation(strip):
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    scene = testscenes.cornell()
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    if strip:
        per = distribution.calculate_shuffled_strips_pixels_per_device(SIZE, 0.5)
        dist = distribution.make_distribution(SIZE, abi.DISTRIBUTION_SHUFFLED_STRIPS, per, per, False)
    else:
        dist = full_dist(SIZE)
    with PathTracer(opt) as pt:
        _setup(pt, scene, cams, dist)
        for _ in range(3):
            pt.render()
        pt.sync()
        want = pt.read_color()
    with PathTracer(opt) as pt:
        _setup(pt, scene, cams, dist)
        for _ in range(2):
            pt.render()
        pt.sync()
        ck = pt.checkpoint()
    assert ck["frame_counter"] == 2 and ck["accumulated_samples"] == 4
    with PathTracer(opt) as pt:
        _setup(pt, scene, cams, dist)
        pt.restore(ck)
        pt.render()
        pt.sync()
        got = pt.read_color()
        assert pt.checkpoint()["accumulated_samples"] == 6
    assert bits_equal(got, want)
    assert not bits_equal(ck["color"], want)


def test_restore_checks_its_arguments():
    opt = abi.default_options(max_ray_depth=2)
    scene = testscenes.cornell()
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    with PathTracer(opt) as pt:
        ck = {"color": np.zeros((1, SIZE[1], SIZE[0], 4), np.float32), "frame_counter": 1, "accumulated_samples": 1}
        with pytest.raises(TaurayB200Error):
            PathTracer.restore(pt, dict(ck, samples_per_pixel=3))   # sample_counter = frame_counter * spp would be wrong
        with pytest.raises(TaurayB200Error):
            pt.restore(ck)                               # no scene / cameras / distribution yet
        _setup(pt, scene, cams, full_dist(SIZE))
        pt.restore(ck)
        with pytest.raises(TaurayB200Error):
            pt.restore(dict(ck, color=ck["color"][:, :-1]))   # wrong size


def test_resume_with_aux_layers_restores_them_too():
    """First-sample layers are never written again once samples are accumulated and the light layers are running
    means: a resumed context must get all of them from the checkpoint, and rendering without them fails loudly."""
    layers = [abi.GBUFFER_DIFFUSE, abi.GBUFFER_REFLECTION, abi.GBUFFER_ALBEDO, abi.GBUFFER_NORMAL, abi.GBUFFER_POS,
              abi.GBUFFER_INSTANCE_ID]
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    scene = testscenes.cornell()
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    dist = full_dist(SIZE)

    def fresh():
        pt = PathTracer(opt)
        _setup(pt, scene, cams, dist)
        pt.set_gbuffer_layers(layers)
        return pt

    with fresh() as pt:
        for _ in range(3):
            pt.render()
        pt.sync()
        want = pt.read_color()
        want_layers = {l: pt.read_gbuffer_layer(l) for l in layers}
    with fresh() as pt:
        for _ in range(2):
            pt.render()
        pt.sync()
        ck = pt.checkpoint()
    assert sorted(ck["layers"]) == sorted(layers)
    with fresh() as pt:
        no_layers = dict(ck, layers={})
        pt.restore(no_layers)
        with pytest.raises(TaurayB200Error):
            pt.render()                                  # layers still missing
        pt.restore(ck)
        pt.render()
        pt.sync()
        assert bits_equal(pt.read_color(), want)
        for l in layers:
            assert pt.read_gbuffer_layer(l).tobytes() == want_layers[l].tobytes(), l
    with PathTracer(opt) as pt:
        _setup(pt, scene, cams, dist)
        with pytest.raises(TaurayB200Error):
            pt.restore(ck)                               # context without the checkpoint's layers
