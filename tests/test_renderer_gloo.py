"""The N>1 path on CPU: world_size 2 (and 3 with unequal workloads) over gloo. Every rank must end the frame with
the full stitched image, bit-identical to the single-device render (the sampler key excludes the rank)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from tauray_b200 import abi, distribution, testscenes
from oracle.oracle_py import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def single_device_image(size):
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    o = Oracle(opt, threads=4)
    cams = testscenes.cornell_cameras(size[0] / size[1])
    o.set_scene(testscenes.cornell())
    o.set_cameras(np.concatenate([cams, cams]))
    o.set_distribution(distribution.make_distribution(size, abi.DISTRIBUTION_DUPLICATE))
    o.render()
    return o.read_color()


@pytest.mark.parametrize("world,workloads,schedule", [(2, "", ""), (3, "0.5,0.3,0.2", ""), (2, "scanline", ""), (3, "scanline_primary0", ""),
                                                      (2, "", "begin_end")])
def test_multi_rank_stitch_equals_single_device(tmp_path, oracle_lib, world, workloads, schedule):
    """`begin_end`: the frames go through RtRenderer.render_begin / render_end (two frames in flight: the stitch of a frame
    follows its end, the next frame is begun before); the last frame starts from a clean accumulation like the others."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_gloo_worker.py"), str(tmp_path), workloads, schedule]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    want = single_device_image((96, 54))
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, "rank%d.npy" % rank))
        assert got.shape == want.shape
        assert got.tobytes() == want.tobytes(), "rank %d image differs from the single-device image" % rank


@pytest.mark.parametrize("world,mode", [(2, "rebalance_primary0"), (3, "rebalance_none")])
def test_workload_change_under_accumulation(tmp_path, oracle_lib, world, mode):
    """set_device_workloads while accumulating (src/rt_renderer.cc:136-183): non-primary devices restart, the stitch
    blends their first strip with 1/(accumulated_frames+1) (stitch_shuffled_strips.comp:50-57), the primary keeps
    accumulating in place. Expected images come from single-device renders only (tests/_rebalance_common.py)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(ROOT, "tests", "_gloo_worker.py"), str(tmp_path), mode]
    env = dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    for rank in range(world):
        bad = np.load(os.path.join(tmp_path, "rebalance_rank%d.npy" % rank))
        assert list(bad[:3]) == [0, 0, 0], "rank %d: mismatching values before/after/next = %s" % (rank, bad[:3])
        assert (bad[3] > 0) == mode.endswith("primary0")
