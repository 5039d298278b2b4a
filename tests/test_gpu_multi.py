"""N>1 on real GPUs (needs >= 2 devices; skipped otherwise): one process per GPU over NCCL, fused NVLink peer-store
stitch and the all-gather baseline, equal and unequal workloads -- every rank's image == the single-device image; plus a workload change under
accumulation (blend-on-rebalance) for both stitch modes, with and without a primary (display) rank."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_ranks_nccl_and_peer_stitch():
    n = gpu_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert r.stdout.count("identical") == (3 + 1 + 1 + 1 + 2 + 4) * world and "DIFFERENT" not in r.stdout, r.stdout[-3000:]
