#!/usr/bin/env python3
"""Collect the reference's own fixtures for the hot path into tests/golden/ (run where /root/reference exists).

* tests/golden/test.glb            <- /root/reference/test/test.glb  (the reference's only test scene; data, not source)
* tests/golden/references/*.exr    <- /root/reference/test/references/validate_{albedo,world-normal,view-normal,
                                      world-pos,view-pos,distance,path-tracer}.exr  (test/CMakeLists.txt:29-37)
* tests/golden/sobol_table_ref.npy <- the 256 x 4 direction numbers parsed from shader/sobol_lookup_table.glsl
* tests/golden/oracle_kat.npz      <- known-answer vectors produced by the ORACLE (regression pins for the parts
                                      the reference's tests leave unpinned: sampler draws, tri-lights, a tiny render)
"""
import os
import re
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(os.path.join(OUT, "references"), exist_ok=True)
    shutil.copyfile(os.path.join(REF, "test", "test.glb"), os.path.join(OUT, "test.glb"))
    for name in ("albedo", "world-normal", "view-normal", "world-pos", "view-pos", "distance", "path-tracer"):
        shutil.copyfile(os.path.join(REF, "test", "references", "validate_%s.exr" % name),
                        os.path.join(OUT, "references", "validate_%s.exr" % name))
    src = open(os.path.join(REF, "shader", "sobol_lookup_table.glsl")).read()
    rows = re.findall(r"uvec4\(0x([0-9A-Fa-f]+)u, 0x([0-9A-Fa-f]+)u, 0x([0-9A-Fa-f]+)u, 0x([0-9A-Fa-f]+)u\)", src)
    np.save(os.path.join(OUT, "sobol_table_ref.npy"), np.array([[int(x, 16) for x in r] for r in rows], dtype=np.uint32))

    from tauray_b200 import abi, distribution, testscenes
    from oracle.oracle_py import Oracle
    kat = {}
    coords = np.random.RandomState(7).randint(0, 4096, size=(64, 4)).astype(np.uint32)
    kat["coords"] = coords
    for sampler, tag in ((abi.SAMPLER_SOBOL_OWEN, "sobol"), (abi.SAMPLER_UNIFORM_RANDOM, "random")):
        o = Oracle(abi.default_options(max_ray_depth=4, local_sampler=sampler, rng_seed=0))
        for d in (-1, 0, 1, 2, 7, 9):
            kat["%s_draw_%d" % (tag, d)] = o.sample_test(coords, d)
    o = Oracle(abi.default_options(max_ray_depth=4, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN))
    sc = testscenes.cornell()
    o.set_scene(sc)
    o.set_cameras(testscenes.cornell_cameras(1.0))
    o.set_distribution(distribution.make_distribution((32, 32), abi.DISTRIBUTION_DUPLICATE))
    o.render()
    kat["cornell_32x32_2spp"] = o.read_color()
    kat["cornell_tri_lights"] = o.tri_lights().view(np.uint8)
    np.savez_compressed(os.path.join(OUT, "oracle_kat.npz"), **kat)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
