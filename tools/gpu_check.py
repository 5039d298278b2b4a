#!/usr/bin/env python3
"""Ad-hoc GPU-vs-oracle parity probe (development aid; the real checks live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tauray_b200 import abi, testscenes, distribution
from tauray_b200.library import PathTracer
from oracle.oracle_py import Oracle


def compare(name, a, b):
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    px_same = same.reshape(-1, 4).all(axis=1)
    denom = np.maximum(np.abs(b), 1e-6)
    rel = np.abs(a - b) / denom
    rel = np.where(np.isfinite(rel), rel, 0)
    print("%-28s bit-exact pixels %.6f  max rel %.3e  within1e-5 %.6f  mean a %.6f b %.6f" % (
        name, px_same.mean(), rel.max(), (rel.reshape(-1, 4).max(axis=1) <= 1e-5).mean(), np.nanmean(a), np.nanmean(b)))
    return px_same.mean()


def main():
    size = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (160, 120)
    sc = testscenes.cornell()
    cams = testscenes.cornell_cameras(size[0] / size[1])
    for sampler in (abi.SAMPLER_SOBOL_OWEN, abi.SAMPLER_UNIFORM_RANDOM):
        opt = abi.default_options(max_ray_depth=4, samples_per_pixel=2, local_sampler=sampler)
        dist = distribution.make_distribution(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, 0,
                                              distribution.calculate_shuffled_strips_pixels_per_device(size, 1.0), True)
        o = Oracle(opt)
        o.set_scene(sc); o.set_cameras(cams); o.set_distribution(dist)
        t = time.time(); o.render(); t_or = time.time() - t
        ref = o.read_color()
        so = o.stats()
        for mode in (abi.KERNEL_MEGAKERNEL, abi.KERNEL_WAVEFRONT):
            opt2 = abi.copy_options(opt); opt2.kernel_mode = mode
            g = PathTracer(opt2)
            g.set_scene(sc); g.set_cameras(cams); g.set_distribution(dist)
            g.render(); g.sync()
            img = g.read_color()
            sg = g.stats()
            compare("sampler %d mode %d" % (sampler, mode), img, ref)
            print("   rays gpu %d+%d oracle %d+%d  gpu ms %.3f launches %d oracle s %.2f" % (
                sg["closest_rays"], sg["shadow_rays"] + sg["culled_shadow_rays"], so["closest_rays"], so["shadow_rays"], sg["render_ms"], sg["kernel_launches"], t_or))
            if mode == abi.KERNEL_WAVEFRONT:
                rays = np.random.RandomState(1).randn(4096, 8).astype(np.float32)
                rays[:, 3] = 0; rays[:, 7] = np.inf
                rays[:, 4:7] /= np.linalg.norm(rays[:, 4:7], axis=1, keepdims=True)
                hg = g.trace_rays(rays); ho = o.trace_rays(rays)
                print("   trace_rays identical:", bool((hg.tobytes() == ho.tobytes())), (hg["instance_id"] >= 0).mean())
                coords = np.random.RandomState(2).randint(0, 2000, size=(1000, 4)).astype(np.uint32)
                for d in (-1, 0, 1, 5, 9):
                    print("   sample draw", d, bool((g.sample_test(coords, d) == o.sample_test(coords, d)).all()))
                print("   tri lights identical:", g.tri_lights().tobytes() == o.tri_lights().tobytes())
            g.close()


if __name__ == "__main__":
    main()
