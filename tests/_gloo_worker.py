"""world_size-N gloo worker for tests/test_renderer_gloo.py: the host-side N>1 path of RtRenderer (strip ranges,
all-gather, un-shuffle) on CPU tensors, with the ORACLE standing in for the GPU tracer (test-only injection)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch.distributed as td  # noqa: E402

from tauray_b200 import abi, testscenes  # noqa: E402
from tauray_b200.renderer import RtRenderer  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402


class OracleTracer:
    """PathTracer-shaped adapter over the oracle."""

    def __init__(self, opt):
        self.o = Oracle(opt, threads=2)

    def __getattr__(self, name):
        return getattr(self.o, name)

    def render(self, stream=None):
        self.o.render()

    # frames in flight (tr_pt_render_begin / tr_pt_render_end): the stand-in only keeps the order of the frames
    def render_begin(self, stream=None):
        self.begun = getattr(self, "begun", 0) + 1
        assert self.begun <= 2

    def render_end(self, stream=None):
        assert getattr(self, "begun", 0) > 0
        self.begun -= 1
        self.o.render()

    def close(self):
        self.o.close()


def rebalance(out_dir, primary):
    """Workload change under accumulation (src/rt_renderer.cc:136-183) against single-device expectations."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _rebalance_common as rc
    from helpers import full_dist
    rank, world = td.get_rank(), td.get_world_size()
    size = (64, 36)
    f1 = 2
    opt = abi.default_options(max_ray_depth=2, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    scene = testscenes.cornell()
    cams = testscenes.cornell_cameras(size[0] / size[1])
    ren = RtRenderer(opt, size, rank=rank, world=world, stitch="nccl", tracer_factory=OracleTracer, backend_is_cuda=False,
                     primary=primary, accumulate=True)
    ren.set_scene(scene, cams)
    for _ in range(f1):
        ren.render()
    before = ren.image().copy()
    ren.set_device_workloads([0.25, 0.75] if world == 2 else [0.2, 0.5, 0.3])
    ren.render()
    img_a = ren.image().copy()
    ren.render()
    img_b = ren.image().copy()
    single = Oracle(opt, threads=2)
    single.set_scene(scene)
    single.set_cameras(cams)
    m_f1, a_pri, a_oth, b_pri, b_oth = rc.expected_images(single, full_dist(size), f1)
    mask = rc.primary_mask(size, ren.dists, primary)
    bad = [rc.mismatches(before, m_f1), rc.mismatches(img_a, rc.compose(mask, a_pri, a_oth)),
           rc.mismatches(img_b, rc.compose(mask, b_pri, b_oth))]
    np.save(os.path.join(out_dir, "rebalance_rank%d.npy" % rank), np.asarray(bad + [int(mask.sum())]))
    ren.close()


def main():
    out_dir = sys.argv[1]
    td.init_process_group("gloo")
    if len(sys.argv) > 2 and sys.argv[2].startswith("rebalance"):
        rebalance(out_dir, 0 if sys.argv[2].endswith("primary0") else None)
        td.destroy_process_group()
        return
    scanline = len(sys.argv) > 2 and sys.argv[2].startswith("scanline")
    primary = 0 if len(sys.argv) > 2 and sys.argv[2] == "scanline_primary0" else None
    workloads = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 and sys.argv[2] and not scanline else None
    rank, world = td.get_rank(), td.get_world_size()
    size = (96, 54)
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    ren = RtRenderer(opt, size, rank=rank, world=world, stitch="nccl", tracer_factory=OracleTracer, backend_is_cuda=False,
                     strategy=abi.DISTRIBUTION_SCANLINE if scanline else abi.DISTRIBUTION_SHUFFLED_STRIPS, primary=primary)
    cams = testscenes.cornell_cameras(size[0] / size[1])
    cams = np.concatenate([cams, cams])            # two views
    ren.set_scene(testscenes.cornell(), cams)
    if workloads:
        ren.set_device_workloads(workloads)
    if len(sys.argv) > 3 and sys.argv[3] == "begin_end":
        # three frames through RtRenderer.render_begin / render_end: begin(k + 1) before end(k), the stitch after every end
        ren.render_begin()
        ren.render_begin()
        ren.render_end()
        ren.reset_accumulation()
        ren.render_end()
    else:
        ren.render()
    img = ren.image()
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), img)
    ren.close()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
