#!/usr/bin/env python3
"""Development aid: acceleration-structure build time (tr_pt_stats.bvh_build_ms) on the bench scene, cold and warm."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tauray_b200 import abi, synthetic
from tauray_b200.library import PathTracer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_500_000
scene = synthetic.atrium(n)
desc = scene.desc()
with PathTracer(abi.default_options()) as pt:
    for i in range(4):
        t0 = time.perf_counter()
        pt.set_scene(desc)
        t1 = time.perf_counter()
        st = pt.stats()
        print("set_scene %d: wall %.1f ms, bvh_build_ms %.2f, nodes %d, depth %d" % (i, (t1 - t0) * 1e3, st["bvh_build_ms"], st["bvh_node_count"], st["bvh_depth"]), flush=True)
