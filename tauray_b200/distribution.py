"""Pixel partitioning across devices: mirror of src/distribution_strategy.{hh,cc}.

Same names and argument meaning as the reference (``calculate_shuffled_strips_b``,
``get_device_distribution_params``, ``get_ray_count`` ...); pure integer host logic, covered by
CPU tests. The shuffled-strips permutation itself (shader/rt.glsl:172-179) is ``permute_region_id``.
"""
import math

import numpy as np

from . import abi


def calculate_shuffled_strips_b(size):
    """src/distribution_strategy.cc:62-69."""
    n = (size[0] * size[1]) & 0xFFFFFFFF
    b = 31
    while (n >> b) < 128 and b > 0:
        b -= 1
    return b


def get_region_size(image_size, b):
    """src/distribution_strategy.cc:73-77."""
    n_regions = 1 << b
    return (image_size + n_regions - 1) // n_regions


def calculate_shuffled_strips_pixels_per_device(size, max_ratio):
    """src/distribution_strategy.cc:79-83. The reference's product is evaluated in binary32: `float * size_t`
    and `float * int` both convert the integer to float, and each product rounds once."""
    b = calculate_shuffled_strips_b(size)
    r = np.float32(max_ratio)
    prod = np.float32(r * np.float32(get_region_size(size[0] * size[1], b)))
    prod = np.float32(prod * np.float32(1 << b))
    return int(math.ceil(float(prod))) & 0xFFFFFFFF


def make_distribution(size, strategy, index=0, count=1, primary=True):
    d = abi.Distribution()
    d.size[0], d.size[1] = int(size[0]), int(size[1])
    d.strategy = int(strategy)
    d.index = int(index)
    d.count = int(count)
    d.primary = 1 if primary else 0
    return d


def get_device_distribution_params(full_image_size, strategy, workload_offset, workload_size, device_index,
                                   device_count, primary):
    """src/distribution_strategy.cc:85-126."""
    if strategy in (abi.DISTRIBUTION_DUPLICATE, abi.DISTRIBUTION_SCANLINE):
        return make_distribution(full_image_size, strategy, device_index, device_count, primary)
    before = calculate_shuffled_strips_pixels_per_device(full_image_size, workload_offset)
    after = calculate_shuffled_strips_pixels_per_device(full_image_size, workload_offset + workload_size)
    return make_distribution(full_image_size, strategy, before, after - before, primary)


def even_workloads(device_count):
    """load_balancer::normalize_workloads with all-zero input (src/load_balancer.cc:34-51)."""
    return [1.0 / device_count] * device_count


def device_distributions(full_image_size, strategy, workloads, primary_device=None):
    """rt_renderer::init_resources' per-device split (src/rt_renderer.cc:245-279): cumulative
    workload offsets -> one distribution per device. primary_device=None: no device renders into
    the full image (every rank keeps a strip; the gather produces the image)."""
    out = []
    offset = 0.0
    n = len(workloads)
    for i, w in enumerate(workloads):
        out.append(get_device_distribution_params(full_image_size, strategy, offset, w, i, n,
                                                  primary_device is not None and i == primary_device))
        offset += w
    return out


def get_distribution_render_size(d):
    """src/distribution_strategy.cc:33-49."""
    if d.strategy == abi.DISTRIBUTION_DUPLICATE:
        return (d.size[0], d.size[1])
    if d.strategy == abi.DISTRIBUTION_SCANLINE:
        return (d.size[0], (d.size[1] - d.index + d.count - 1) // d.count)
    return (d.count, 1)


def get_ray_count(d):
    """src/distribution_strategy.cc:51-60."""
    return get_distribution_render_size(d)


def target_shape(d, views):
    """Shape of the colour target of one device as float32 (views, ..., 4): the full image for the
    primary device, else the strip image (get_distribution_target_size,
    src/distribution_strategy.cc:5-19; we keep strips unpadded: `count` pixels per view)."""
    if d.primary or d.strategy == abi.DISTRIBUTION_DUPLICATE:
        return (views, d.size[1], d.size[0], 4)
    rx, ry = get_ray_count(d)
    if d.strategy == abi.DISTRIBUTION_SCANLINE:
        return (views, ry, rx, 4)
    return (views, rx, 4)


def permute_region_id(i, size, b=None):
    """shader/rt.glsl:172-179 / stitch_shuffled_strips.comp:21-27, vectorised over numpy uint32."""
    i = np.asarray(i, dtype=np.uint64)
    if b is None:
        b = calculate_shuffled_strips_b(size)
    region_size = ((size[0] * size[1]) + (1 << b) - 1) >> b
    region_id = (i // region_size).astype(np.uint32)
    x = region_id
    x = ((x >> 1) & 0x55555555) | ((x & 0x55555555) << 1)
    x = ((x >> 2) & 0x33333333) | ((x & 0x33333333) << 2)
    x = ((x >> 4) & 0x0F0F0F0F) | ((x & 0x0F0F0F0F) << 4)
    x = ((x >> 8) & 0x00FF00FF) | ((x & 0x00FF00FF) << 8)
    x = ((x >> 16) | (x << 16)) & 0xFFFFFFFF
    k = (x.astype(np.uint64) >> (32 - b))
    return (k * region_size + i % region_size).astype(np.uint64)
