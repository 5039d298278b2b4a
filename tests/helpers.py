"""Shared helpers of the parity tests."""
import numpy as np

from tauray_b200 import abi, distribution


def full_dist(size, strategy=abi.DISTRIBUTION_SHUFFLED_STRIPS):
    if strategy == abi.DISTRIBUTION_SHUFFLED_STRIPS:
        return distribution.make_distribution(size, strategy, 0,
                                              distribution.calculate_shuffled_strips_pixels_per_device(size, 1.0), True)
    return distribution.make_distribution(size, strategy, 0, 1, True)


def run(tracer, scene, cams, dist, frames=1, reset=True):
    tracer.set_scene(scene)
    tracer.set_cameras(cams)
    tracer.set_distribution(dist)
    for _ in range(frames):
        tracer.render()
    if hasattr(tracer, "sync"):
        tracer.sync()
    return tracer.read_color()


def bits_equal(a, b):
    """Bit-exact float comparison that treats identical NaN payload-agnostic NaNs as equal."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if a.shape != b.shape:
        return False
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    return bool(same.all())


# The tolerance BASELINE.json's north_star states: 1e-5 relative per channel, before tonemap.
NORTH_STAR_RTOL = 1e-5


def assert_parity(img, ref, what=""):
    """GPU-vs-oracle parity: within 1e-5 relative per channel (the north-star tolerance); the build is
    designed to be bit-exact, which is reported when it is not."""
    assert img.shape == ref.shape, (what, img.shape, ref.shape)
    nan_a, nan_b = np.isnan(img), np.isnan(ref)
    assert (nan_a == nan_b).all(), "%s: NaN pattern differs" % what
    a = np.where(nan_a, 0, img).astype(np.float64)
    b = np.where(nan_b, 0, ref).astype(np.float64)
    err = np.abs(a - b)
    tol = NORTH_STAR_RTOL * np.abs(b) + 1e-12
    bad = err > tol
    assert not bad.any(), "%s: %d values beyond 1e-5 relative (max rel %.3e)" % (
        what, int(bad.sum()), float((err / np.maximum(np.abs(b), 1e-30)).max()))
    return bits_equal(img, ref)
