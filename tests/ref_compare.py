"""Comparison rules of the oracle-vs-reference-shader parity tests (tests/test_ref_glsl.py, tools/make_ref_fixtures.py)."""
import numpy as np

# Fraction of random inputs allowed beyond 1e-5 relative (of the record's magnitude). Default 0.5 %: inputs on a branch
# (`u <= cutoff`, cone rim) or next to a cancellation. Exceptions, all bounded at 1e-3 by OUTLIERS_1E3:
#   15/25/26/40  spherical-triangle sampling, its pdf and barycentrics: thin triangles / points near an edge cancel
#   20           material_bsdf_sample: pow(x, 5) (libm vs repeated multiplication) and fused vs unfused mix() give a
#                systematic few-1e-5 difference of the lobe weights; nothing beyond 1e-4
#   22           sample_point_light: out_length = -b - sqrt(b*b - dist2 + r*r) is the square root of rounding noise for
#                small radii -- 1e-4..1e-3 relative in any binary32 implementation
OUTLIERS_1E5 = {15: 0.03, 25: 0.03, 26: 0.03, 40: 0.04, 20: 0.09, 22: 0.2, 12: 0.01, 39: 0.01, 36: 0.01}
DEFAULT_1E5 = 0.005
OUTLIERS_1E3 = 0.001


def floats(words):
    return np.ascontiguousarray(words).view(np.float32).astype(np.float64)


def compare_function(fn, kinds, ref_out, orc_out):
    """Returns a list of failure messages (empty = parity holds)."""
    msgs = []
    ii = [k for k, kind in enumerate(kinds) if kind == "i"]
    fi = [k for k, kind in enumerate(kinds) if kind == "f"]
    if ii:
        bad = int((ref_out[:, ii] != orc_out[:, ii]).any(axis=1).sum())
        if bad:
            msgs.append("%d records differ in an integer output" % bad)
    if fi:
        x, y = floats(ref_out)[:, fi], floats(orc_out)[:, fi]
        nan_x, nan_y = np.isnan(x), np.isnan(y)
        if (nan_x != nan_y).any():
            msgs.append("NaN pattern differs in %d records" % int((nan_x != nan_y).any(axis=1).sum()))
        with np.errstate(all="ignore"):
            inf_mis = (np.isinf(x) != np.isinf(y)) | (np.isinf(x) & np.isinf(y) & (np.sign(x) != np.sign(y)))
            err = np.abs(x - y)
            den = np.maximum(np.abs(x), np.abs(y))
            err = np.where(np.isfinite(err), err, 0.0)
            den = np.where(np.isfinite(den), den, 0.0)
        scale = np.maximum(den.max(axis=1, keepdims=True), 1e-3)
        n = x.shape[0]
        for rt, limit in ((1e-5, OUTLIERS_1E5.get(fn, DEFAULT_1E5)), (1e-3, OUTLIERS_1E3)):
            bad = ((err > rt * den + 0.1 * rt * scale) | inf_mis).any(axis=1)
            limit = max(limit, 2.0 / n)          # small fixtures: a couple of records may sit on a branch
            if bad.mean() > limit:
                msgs.append("%.3f %% of the records beyond %g relative (limit %.3f %%)" % (100 * bad.mean(), rt, 100 * limit))
    return msgs


def compare_images(ref, img, what=""):
    """Reference-shader image vs oracle image: a path tracer is chaotic (a last-bit difference in a direction can hit
    another triangle), so the criterion is per pixel with a small allowance: >= 98.5 % of the pixels within 1e-4
    relative, and the image means within 1e-3 relative."""
    assert ref.shape == img.shape
    a, b = ref[..., :3].astype(np.float64), img[..., :3].astype(np.float64)
    assert (np.isnan(a) == np.isnan(b)).mean() > 0.999, what
    a, b = np.nan_to_num(a), np.nan_to_num(b)
    rel = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-3)
    close = (rel.max(axis=-1) <= 1e-4).mean()
    mean_rel = abs(a.mean() - b.mean()) / max(a.mean(), 1e-9)
    return close, mean_rel
