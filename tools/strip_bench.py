#!/usr/bin/env python3
"""Development aid: time ONE rank's share of the N-GPU bench workload on one GPU (no stitch), to study how the
kernels behave on small per-rank queues. usage: strip_bench.py <world> [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tauray_b200 import abi, distribution, synthetic
from tauray_b200.library import PathTracer
import bench

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 10
_, W, H, SPP, _, _ = bench.CONFIGS["c3"]
size = (W, H)
opt = bench.bench_options(abi, SPP)
scene = synthetic.atrium(bench.TRIANGLES)
cams = synthetic.atrium_cameras(size[0] / size[1])
dists = distribution.device_distributions(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, [1.0 / world] * world)
with PathTracer(opt) as pt:
    pt.set_scene(scene); pt.set_cameras(cams); pt.set_distribution(dists[world // 2])
    for timing in (0, abi.INSTRUMENT_TIMING):
        pt.set_instrumentation(timing)
        for _ in range(3):
            pt.reset_accumulation(); pt.render()
        pt.sync()
        t0 = time.perf_counter()
        for _ in range(frames):
            pt.reset_accumulation(); pt.render()
        pt.sync()
        ms = 1e3 * (time.perf_counter() - t0) / frames
        st = pt.stats()
        if timing == 0:
            # the same with two frames in flight (tr_pt_render_begin / tr_pt_render_end)
            pt.reset_accumulation(); pt.render_begin()
            for _ in range(3):
                pt.reset_accumulation(); pt.render_begin(); pt.render_end()
            pt.sync()
            t0 = time.perf_counter()
            for _ in range(frames):
                pt.reset_accumulation(); pt.render_begin(); pt.render_end()
            pt.sync()
            ms2 = 1e3 * (time.perf_counter() - t0) / frames
            pt.render_end(); pt.sync()
            print("world %d: %.3f ms/frame (ideal %.3f of the full frame's), two frames in flight %.3f" % (world, ms, 1.0 / world, ms2), end="  ")
        else:
            print("closest %.3f shadow %.3f shade %.3f other %.3f" % (st["closest_ms"], st["shadow_ms"], st["shade_ms"], st["other_ms"]))
