"""Scanline distribution across devices (src/distribution_strategy.cc:39-43,100-105; stitch_scanline.comp:21-50) on one
GPU: three contexts play three devices. Both stitch paths -- the scatter kernel behind tr_pt_unshuffle and the fused
peer stores of the last accumulation pass -- must rebuild the single-device image bit for bit."""
import numpy as np
import pytest

from tauray_b200 import abi, distribution, testscenes
from tauray_b200.library import PathTracer, TaurayB200Error

from helpers import bits_equal, full_dist, run

pytestmark = pytest.mark.gpu

SIZE = (88, 50)          # 50 rows over 3 devices: 17 + 17 + 16


def _setup(views=2):
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    scene = testscenes.cornell()
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    cams = np.concatenate([cams] * views)
    with PathTracer(opt) as single:
        want = run(single, scene, cams, full_dist(SIZE))
    return opt, scene, cams, want


def test_scanline_rows_scatter_into_the_full_image():
    opt, scene, cams, want = _setup()
    dists = distribution.device_distributions(SIZE, abi.DISTRIBUTION_SCANLINE, [1 / 3] * 3)
    assert [distribution.get_ray_count(d)[1] for d in dists] == [17, 17, 16]
    tracers = [PathTracer(opt) for _ in dists]
    try:
        for pt, d in zip(tracers, dists):
            pt.set_scene(scene); pt.set_cameras(cams); pt.set_distribution(d)
        full_ptr, nbytes = tracers[0].alloc_full_image()
        assert nbytes == 2 * SIZE[0] * SIZE[1] * 16
        for pt, d in zip(tracers, dists):
            pt.render(); pt.sync()
            rows = pt.read_color()
            assert rows.shape == (2, distribution.get_ray_count(d)[1], SIZE[0], 4)
            assert bits_equal(rows, want[:, d.index::d.count])
        for pt, d in zip(tracers, dists):
            ptr, _ = pt.color_buffer()
            tracers[0].unshuffle(ptr, d.index, d.count, full_ptr)
        tracers[0].sync()
        got = tracers[0].read_full_image(np.empty((2, SIZE[1], SIZE[0], 4), dtype=np.float32))
        assert bits_equal(got, want)
        # blend: mix(old, new, 0.25) of identical images is the image itself only where fp32 rounding allows; check the formula
        tracers[0].unshuffle(tracers[1].color_buffer()[0], 1, 3, full_ptr, None, 0.25)
        tracers[0].sync()
        again = tracers[0].read_full_image(np.empty((2, SIZE[1], SIZE[0], 4), dtype=np.float32))
        assert bits_equal(again[:, 0::3], want[:, 0::3]) and bits_equal(again[:, 2::3], want[:, 2::3])
        assert np.allclose(again[:, 1::3], want[:, 1::3], rtol=1e-6, atol=1e-7)
        with pytest.raises(TaurayB200Error):
            tracers[0].unshuffle(tracers[1].color_buffer()[0], 3, 3, full_ptr)      # device index out of range
    finally:
        for pt in tracers:
            pt.close()


def test_scanline_fused_peer_stores_with_a_primary_device():
    opt, scene, cams, want = _setup(views=1)
    dists = distribution.device_distributions(SIZE, abi.DISTRIBUTION_SCANLINE, [1 / 3] * 3, primary_device=0)
    tracers = [PathTracer(opt) for _ in dists]
    try:
        for pt, d in zip(tracers, dists):
            pt.set_scene(scene); pt.set_cameras(cams); pt.set_distribution(d)
        full_ptr, nbytes = tracers[0].alloc_full_image()
        tracers[0].set_color_buffer(full_ptr, nbytes)             # the display device renders in place
        for pt in tracers[1:]:
            pt.set_peer_images([full_ptr])
        for pt in tracers:
            pt.render()
        for pt in tracers:
            pt.sync()
        got = tracers[0].read_full_image(np.empty((1, SIZE[1], SIZE[0], 4), dtype=np.float32))
        assert bits_equal(got, want)
    finally:
        for pt in tracers:
            pt.close()
