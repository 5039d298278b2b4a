"""Shared by the gloo and NCCL workers: the expected images of a workload change under accumulation, derived from
single-device renders only (src/rt_renderer.cc:136-183 + stitch_shuffled_strips.comp:50-57).

Sequence: F1 accumulated frames with the old strips; set_device_workloads(new); frame A (non-primary devices restart
their accumulation, their strips are blended with 1/(F1+1)); frame B (blend back to 1).

  pixel owned (new strips) by the primary : keeps accumulating in place  -> M[F1+1], then M[F1+2]
  pixel owned by any other device          : mix(M[F1], S1, 1/(F1+1)), then S2 stored plainly
where M[k] = k frames accumulated from scratch, S1 / S2 = 1 / 2 frames accumulated after reset_accumulation(False)
at frame F1 (the sample counter keeps running). Every device computes identical per-pixel samples (the sampler key
excludes the device), so single-device renders give all of these."""
import numpy as np

from tauray_b200 import distribution as dist_mod


def mix_f32(old, new, a):
    """mix(old, new, a) = fma(new, a, old * (1 - a)) in fp32 (new * a is exact in fp64)."""
    a = np.float32(a)
    t = (old.astype(np.float32) * (np.float32(1.0) - a)).astype(np.float32)
    return (new.astype(np.float64) * np.float64(a) + t.astype(np.float64)).astype(np.float32)


def expected_images(single, full_dist, f1):
    """`single`: a tracer with scene + cameras set. Returns (M[f1], expA_primary, expA_other, expB_primary, expB_other)."""
    single.set_distribution(full_dist)
    single.reset_accumulation(True)
    acc = {}
    for k in range(1, f1 + 3):
        single.render()
        if hasattr(single, "sync"):
            single.sync()
        acc[k] = single.read_color().copy()
    single.reset_accumulation(True)
    for _ in range(f1):
        single.render()
    single.reset_accumulation(False)
    single.render()
    if hasattr(single, "sync"):
        single.sync()
    s1 = single.read_color().copy()
    single.render()
    if hasattr(single, "sync"):
        single.sync()
    s2 = single.read_color().copy()
    exp_a_other = mix_f32(acc[f1], s1, 1.0 / (f1 + 1))
    return acc[f1], acc[f1 + 1], exp_a_other, acc[f1 + 2], s2


def primary_mask(size, dists, primary):
    """Boolean (H*W,) mask of the pixels the primary device owns under `dists`."""
    W, H = size
    m = np.zeros(W * H, dtype=bool)
    if primary is None:
        return m
    d = dists[primary]
    j = dist_mod.permute_region_id(np.arange(d.index, d.index + d.count, dtype=np.uint64), size)
    m[j[j < W * H].astype(np.int64)] = True
    return m


def compose(mask, img_primary, img_other):
    v = img_primary.shape[0]
    out = img_other.reshape(v, -1, 4).copy()
    out[:, mask] = img_primary.reshape(v, -1, 4)[:, mask]
    return out.reshape(img_primary.shape)


def mismatches(a, b):
    return int((a.view(np.uint32) != b.view(np.uint32)).sum())
