#!/usr/bin/env python3
"""Development aid: do two frames in flight pay? Two contexts on ONE GPU render the same share of the frame on their own
streams, alternately and without a sync between them; compared with one context rendering the same number of frames.
usage: two_ctx_bench.py <world> [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tauray_b200 import abi, distribution, synthetic
from tauray_b200.library import PathTracer
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20
_, W, H, SPP, _, _ = bench.CONFIGS["c3"]
size = (W, H)
opt = bench.bench_options(abi, SPP)
scene = synthetic.atrium(bench.TRIANGLES)
cams = synthetic.atrium_cameras(size[0] / size[1])
dists = distribution.device_distributions(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, [1.0 / world] * world)
pts = []
for k in range(2):
    pt = PathTracer(opt)
    pt.set_scene(scene); pt.set_cameras(cams); pt.set_distribution(dists[world // 2])
    pts.append(pt)

def run(contexts, n):
    for pt in contexts:
        for _ in range(3):
            pt.reset_accumulation(); pt.render()
    for pt in contexts: pt.sync()
    t0 = time.perf_counter()
    for i in range(n):
        pt = contexts[i % len(contexts)]
        pt.reset_accumulation(); pt.render()
    for pt in contexts: pt.sync()
    return 1e3 * (time.perf_counter() - t0) / n

one = run(pts[:1], 2 * frames)
two = run(pts, 2 * frames)
print("world %d: one context %.3f ms/frame, two contexts in flight %.3f ms/frame (x%.3f)" % (world, one, two, one / two))
for pt in pts: pt.close()
