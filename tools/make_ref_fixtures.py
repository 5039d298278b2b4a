#!/usr/bin/env python3
"""Writes tests/golden/ref_glsl_*.npz from the live oracle/_ref library (the reference's own shader text compiled for
the CPU, oracle/ref_glsl/build_ref.py): inputs + reference outputs of the function table, and the reference pipeline's
frames of tests/test_ref_glsl.py's image cases. Needs /root/reference; the fixtures let the same parity checks run where
it is absent. usage: python tools/make_ref_fixtures.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ref_glsl import ref_py  # noqa: E402
import ref_cases as rc  # noqa: E402
import test_ref_glsl as T  # noqa: E402


def main():
    assert ref_py.available(), "oracle/_ref cannot be built: /root/reference is missing"
    golden = os.path.join(ROOT, "tests", "golden")
    n = 192
    for variant in ("v0", "v2", "v3"):
        out = {}
        for fn, (name, kinds, gen) in sorted(rc.FUNCTIONS.items()):
            rng = np.random.RandomState(5000 + fn)
            rec = gen(rng, n, variant) if gen else T.tangent_space_records(rng, n, lambda f, r: ref_py.function(variant, f, r))
            out["in_%d" % fn] = rec
            out["out_%d" % fn] = ref_py.function(variant, fn, rec)
        np.savez_compressed(os.path.join(golden, "ref_glsl_functions_%s.npz" % variant), **out)
    images = {}
    for name in sorted(T.IMAGE_CASES):
        variant, scene, cams, size, opt = T.image_case(name)
        o, _, _ = T.oracle_frames(scene, cams, size, opt, 0)
        col, _, _, _ = ref_py.RefPathTracer(variant, scene, cams, o).render(size, opt, frames=2)
        images[name] = col
    np.savez_compressed(os.path.join(golden, "ref_glsl_images.npz"), **images)
    print("written")


if __name__ == "__main__":
    main()
