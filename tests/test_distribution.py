"""Host logic of the pixel distribution (src/distribution_strategy.cc, shader/rt.glsl:170-222)."""
import numpy as np
import pytest

from tauray_b200 import abi, distribution as D
from tauray_b200.renderer import LoadBalancer


@pytest.mark.parametrize("size,b,region", [((1920, 1080), 13, 254), ((512, 512), 11, 128), ((960, 540), 11, 254)])
def test_shuffled_strips_parameters_from_the_survey(size, b, region):
    assert D.calculate_shuffled_strips_b(size) == b
    assert D.get_region_size(size[0] * size[1], b) == region


def test_1080p_eight_gpu_split():
    size = (1920, 1080)
    ds = D.device_distributions(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, D.even_workloads(8))
    assert [d.count for d in ds] == [260096] * 8
    assert [d.index for d in ds] == [260096 * i for i in range(8)]
    assert sum(d.count for d in ds) == 254 * 8192 >= 1920 * 1080


@pytest.mark.parametrize("size", [(1920, 1080), (512, 512), (960, 540), (160, 120), (33, 17)])
@pytest.mark.parametrize("n", [1, 2, 3, 8])
def test_partition_is_a_bijection_onto_the_image(size, n):
    ds = D.device_distributions(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, D.even_workloads(n))
    seen = np.zeros(size[0] * size[1], dtype=np.int32)
    end = 0
    for d in ds:
        assert d.index == end
        end = d.index + d.count
        j = D.permute_region_id(np.arange(d.index, d.index + d.count, dtype=np.uint64), size)
        j = j[j < size[0] * size[1]].astype(np.int64)
        np.add.at(seen, j, 1)
    assert (seen == 1).all()


def test_unequal_workloads():
    size = (640, 360)
    ds = D.device_distributions(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, [0.5, 0.3, 0.2])
    total = D.calculate_shuffled_strips_pixels_per_device(size, 1.0)
    assert ds[0].index == 0 and ds[-1].index + ds[-1].count == total
    assert abs(ds[0].count / total - 0.5) < 1e-3


def test_scanline_and_duplicate():
    d = D.get_device_distribution_params((100, 37), abi.DISTRIBUTION_SCANLINE, 0, 0, 2, 4, False)
    assert (d.index, d.count) == (2, 4)
    assert D.get_ray_count(d) == (100, (37 - 2 + 3) // 4)
    assert D.target_shape(d, 3) == (3, 9, 100, 4)
    d = D.make_distribution((100, 37), abi.DISTRIBUTION_DUPLICATE)
    assert D.get_ray_count(d) == (100, 37) and D.target_shape(d, 2) == (2, 37, 100, 4)


def test_load_balancer_matches_the_reference_update():
    lb = LoadBalancer([0, 0])
    assert lb.workloads == [0.5, 0.5]
    w = lb.update([2.0, 1.0])     # device 1 is twice as fast
    assert w[1] > w[0] and abs(sum(w) - 1.0) < 1e-9
    assert abs(w[0] - (0.5 * 0.9 + (0.25 / 0.75) * 0.1)) < 1e-9


def test_oracle_scanline_scatter_matches_numpy():
    """orc_unshuffle under the scanline distribution (stitch_scanline.comp:21-50): row ly of device `index` of
    `count` lands on image row ly * count + index; blend_ratio < 1 mixes with the previous content."""
    from tauray_b200 import testscenes
    from oracle.oracle_py import Oracle
    size = (10, 7)
    views = 2
    o = Oracle(abi.default_options(max_ray_depth=1))
    o.set_scene(testscenes.cornell(subdiv=1))
    cams = testscenes.cornell_cameras(size[0] / size[1])
    o.set_cameras(np.concatenate([cams] * views))
    rng = np.random.RandomState(5)
    full = rng.rand(views, size[1], size[0], 4).astype(np.float32)
    want = full.copy()
    for index in range(3):
        o.set_distribution(D.make_distribution(size, abi.DISTRIBUTION_SCANLINE, index, 3, False))
        rows = len(range(index, size[1], 3))
        strip = rng.rand(views, rows, size[0], 4).astype(np.float32)
        blend = 1.0 if index != 1 else 0.25
        o.unshuffle(strip, index, 3, full, blend)
        if blend == 1.0:
            want[:, index::3] = strip
        else:
            a = np.float32(blend)
            old = want[:, index::3]
            t = (old * (np.float32(1.0) - a)).astype(np.float32)
            want[:, index::3] = (strip.astype(np.float64) * np.float64(a) + t.astype(np.float64)).astype(np.float32)
    assert full.tobytes() == want.tobytes()


def test_strip_boundaries_follow_the_reference_float32_arithmetic(tmp_path):
    """calculate_shuffled_strips_pixels_per_device multiplies in binary32 (src/distribution_strategy.cc:79-83):
    at interior split points that differs from a double product by one pixel (1280x720 over 3 devices: 307200,
    not 307201). Checked against the C++ mirror of the same function, compiled here."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "strips.cc"
    src.write_text("""
#include "tauray_b200/host/tauray_b200.hh"
#include <cstdio>
int main()
{
    const unsigned sizes[][2] = {{1280, 720}, {640, 360}, {3840, 2160}, {1920, 1080}, {960, 540}};
    for(auto& s: sizes)
        for(unsigned world: {2u, 3u, 5u, 6u, 7u, 8u})
            for(unsigned r = 0; r <= world; ++r)
            {
                tr_b200::uvec2 sz; sz.x = s[0]; sz.y = s[1];
                double off = 0; for(unsigned k = 0; k < r; ++k) off += 1.0 / world;
                printf("%u %u %u %u %u\\n", s[0], s[1], world, r, tr_b200::calculate_shuffled_strips_pixels_per_device(sz, (float)off));
            }
}
""".replace("\\\\n", "\\n"))
    exe = tmp_path / "strips"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", root, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    rows = [tuple(int(x) for x in line.split()) for line in out.splitlines()]
    assert len(rows) > 100
    for w, h, world, r, want in rows:
        off = 0.0
        for _ in range(r):
            off += 1.0 / world
        assert D.calculate_shuffled_strips_pixels_per_device((w, h), off) == want, (w, h, world, r)
    assert D.calculate_shuffled_strips_pixels_per_device((1280, 720), 1.0 / 3.0) == 307200
