#!/usr/bin/env python3
"""Development aid: python tools/drain_probe.py <world> -- prints, for one frame of one rank's share of the bench frame, the
start and end of every kernel (%globaltimer), and per traversal launch when the queue ran empty, when the last warp left
and a histogram of the warps' exit times (tr_render.cu, TR_DEBUG_DRAIN). TR_TUNE_SPLIT_STAGE=1 adds tail splitting."""
import os, sys
# the probes live in the SPLIT instantiation of the traversal kernel; its tail stage stays off unless asked for, so that
# what is timed is the shipped schedule
os.environ.setdefault("TR_DEBUG_DRAIN", "1")
os.environ.setdefault("TR_TUNE_SPLIT", "3")
os.environ.setdefault("TR_TUNE_SPLIT_STAGE", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tauray_b200 import abi, distribution, synthetic
from tauray_b200.library import PathTracer
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
_, W, H, SPP, _, _ = bench.CONFIGS["c3"]
size = (W, H)
opt = bench.bench_options(abi, SPP)
scene = synthetic.atrium(bench.TRIANGLES)
cams = synthetic.atrium_cameras(size[0] / size[1])
dists = distribution.device_distributions(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, [1.0 / world] * world)
with PathTracer(opt) as pt:
    pt.set_scene(scene); pt.set_cameras(cams); pt.set_distribution(dists[world // 2])
    for i in range(3):
        sys.stderr.write("frame %d\n" % i); sys.stderr.flush()
        pt.reset_accumulation(); pt.render(); pt.sync()
