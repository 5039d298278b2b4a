"""world_size-N NCCL worker for tests/test_gpu_multi.py: real PathTracers, one per GPU, both stitch modes; every rank
checks that its stitched full image equals a single-device render bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

from tauray_b200 import abi, distribution, synthetic  # noqa: E402
from tauray_b200.library import PathTracer  # noqa: E402
from tauray_b200.renderer import RtRenderer  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = run_checks(rank, world, local)
    td.destroy_process_group()
    sys.exit(0 if ok else 1)


def run_checks(rank, world, local):
    """All multi-GPU checks on an initialised NCCL process group (also behind `bench.py --gpus N --selftest`)."""
    size = (640, 360)
    opt = abi.default_options(max_ray_depth=4, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    scene = synthetic.atrium(60000)
    cams = synthetic.atrium_cameras(16 / 9)
    cams = np.concatenate([cams, cams])
    with PathTracer(opt, device=local) as single:
        single.set_scene(scene)
        single.set_cameras(cams)
        single.set_distribution(distribution.make_distribution(
            size, abi.DISTRIBUTION_SHUFFLED_STRIPS, 0, distribution.calculate_shuffled_strips_pixels_per_device(size, 1.0), True))
        single.render()
        single.sync()
        want = single.read_color()
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    ok = True
    for stitch, workloads in (("peer", None), ("nccl", None), ("peer", [0.7] + [0.3 / (world - 1)] * (world - 1))):
        ren = RtRenderer(opt, size, rank=rank, world=world, device=local, stitch=stitch)
        ren.set_scene(scene, cams)
        if workloads:
            ren.set_device_workloads(workloads)
        for _ in range(2):                      # second frame: accumulation reset + re-stitch over the same buffers
            ren.reset_accumulation()
            ren.render(stream.cuda_stream)
        torch.cuda.synchronize(local)
        td.barrier()
        got = ren.image()
        same = got.tobytes() == want.tobytes()
        print("rank %d stitch %s workloads %s: %s" % (rank, stitch, workloads, "identical" if same else "DIFFERENT"), flush=True)
        ok = ok and same
        td.barrier()
        ren.close()
    # ---- consecutive frames with different cameras, no barrier between them, odd ranks slow to read back: a fast
    # rank's frame k+1 must not tear the image a slow rank is still copying out (double-buffered full images)
    cam_seq = []
    for k in range(4):
        ck = cams.copy()
        ck["current"]["origin"][:, 0] += 0.35 * k
        ck["current"]["view_inverse"][:, 3, 0] += 0.35 * k    # column-major wire layout: [column][row]
        cam_seq.append(ck)
    wants = []
    with PathTracer(opt, device=local) as single:
        single.set_scene(scene)
        single.set_distribution(distribution.make_distribution(
            size, abi.DISTRIBUTION_SHUFFLED_STRIPS, 0, distribution.calculate_shuffled_strips_pixels_per_device(size, 1.0), True))
        for ck in cam_seq:
            single.set_cameras(ck)
            single.reset_accumulation()
            single.render()
            single.sync()
            wants.append(single.read_color().copy())
    assert wants[0].tobytes() != wants[1].tobytes()
    ren = RtRenderer(opt, size, rank=rank, world=world, device=local, stitch="peer")
    ren.set_scene(scene, cams)
    same = True
    for k, ck in enumerate(cam_seq):
        ren.tracer.set_cameras(ck)
        ren.reset_accumulation()
        ren.render(stream.cuda_stream)
        if rank % 2 == 1:
            torch.cuda._sleep(400_000_000)          # ~0.2 s on the render stream before this rank reads frame k
        got = ren.image()
        same = same and got.tobytes() == wants[k].tobytes()
    print("rank %d consecutive frames while a peer reads back: %s" % (rank, "identical" if same else "DIFFERENT"), flush=True)
    ok = ok and same
    torch.cuda.synchronize(local)
    td.barrier()
    # ---- frames in flight: every rank reads the WHOLE stitched image of frame k back asynchronously (copy stream, pinned
    # host memory, read_back_async) while frame k + 1 renders and the peers store into the other half; odd ranks' copies
    # start late. Each host image must be the single-device frame.
    hosts = [torch.zeros(wants[0].shape, dtype=torch.float32).pin_memory() for _ in cam_seq]
    for k, ck in enumerate(cam_seq):
        ren.tracer.set_cameras(ck)
        ren.reset_accumulation()
        ren.render(stream.cuda_stream)
        if rank % 2 == 1:
            torch.cuda._sleep(100_000_000)
        ren.read_back_async(hosts[k].data_ptr(), 0, hosts[k].numel() * 4, stream.cuda_stream)
    torch.cuda.synchronize(local)
    same = all(hosts[k].numpy().tobytes() == wants[k].tobytes() for k in range(len(cam_seq)))
    print("rank %d asynchronous read-back with two frames in flight: %s" % (rank, "identical" if same else "DIFFERENT"), flush=True)
    ok = ok and same
    td.barrier()
    # ---- frames in flight on the device too: begin(k + 1) before end(k) (tr_pt_render_begin / tr_pt_render_end), camera
    # tables in stream order, read-backs asynchronous, odd ranks' copies late
    hosts = [torch.zeros(wants[0].shape, dtype=torch.float32).pin_memory() for _ in cam_seq]
    pinned = [torch.from_numpy(np.frombuffer(ck.tobytes(), dtype=np.uint8).copy()).pin_memory() for ck in cam_seq]

    def begin(k):
        ren.tracer.set_cameras_async(pinned[k].numpy().view(abi.CAMERA_PAIR), stream.cuda_stream)
        ren.reset_accumulation()
        ren.render_begin(stream.cuda_stream)

    begin(0)
    for k in range(len(cam_seq)):
        if k + 1 < len(cam_seq):
            begin(k + 1)
        ren.render_end(stream.cuda_stream)
        if rank % 2 == 1:
            torch.cuda._sleep(100_000_000)
        ren.read_back_async(hosts[k].data_ptr(), 0, hosts[k].numel() * 4, stream.cuda_stream)
    torch.cuda.synchronize(local)
    same = all(hosts[k].numpy().tobytes() == wants[k].tobytes() for k in range(len(cam_seq)))
    print("rank %d frames in flight on the device (begin / end): %s" % (rank, "identical" if same else "DIFFERENT"), flush=True)
    ok = ok and same
    td.barrier()
    ren.close()
    # ---- aux G-buffer layers across devices (src/rt_renderer.cc:370-397): every rank's stitched layers == one device's
    layers = [abi.GBUFFER_DIFFUSE, abi.GBUFFER_REFLECTION, abi.GBUFFER_ALBEDO, abi.GBUFFER_MATERIAL, abi.GBUFFER_NORMAL,
              abi.GBUFFER_POS, abi.GBUFFER_SCREEN_MOTION, abi.GBUFFER_INSTANCE_ID]
    with PathTracer(opt, device=local) as single:
        single.set_scene(scene)
        single.set_cameras(cams)
        single.set_distribution(distribution.make_distribution(
            size, abi.DISTRIBUTION_SHUFFLED_STRIPS, 0, distribution.calculate_shuffled_strips_pixels_per_device(size, 1.0), True))
        single.set_gbuffer_layers(layers)
        for _ in range(3):
            single.render()
        single.sync()
        want_col = single.read_color().copy()
        want_layers = {l: single.read_gbuffer_layer(l).copy() for l in layers}
    for primary in (None, 0):
        ren = RtRenderer(opt, size, rank=rank, world=world, device=local, stitch="peer", primary=primary, accumulate=True)
        ren.set_scene(scene, cams)
        ren.set_gbuffer_layers(layers)
        for _ in range(3):                       # three accumulated frames: running means + carried-over first-sample layers
            ren.render(stream.cuda_stream)
        got_col = ren.image()
        bad = [] if got_col.tobytes() == want_col.tobytes() else ["color"]
        for l in layers:
            if ren.layer_image(l).tobytes() != want_layers[l].tobytes():
                bad.append(l)
        print("rank %d aux layers primary %s: %s %s" % (rank, primary, "identical" if not bad else "DIFFERENT", bad), flush=True)
        ok = ok and not bad
        torch.cuda.synchronize(local)
        td.barrier()
        ren.close()
    # ---- workload change under accumulation (src/rt_renderer.cc:136-183, stitch_shuffled_strips.comp:50-57)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _rebalance_common as rc
    f1 = 2
    size2 = (320, 180)
    full = distribution.make_distribution(size2, abi.DISTRIBUTION_SHUFFLED_STRIPS, 0,
                                          distribution.calculate_shuffled_strips_pixels_per_device(size2, 1.0), True)
    with PathTracer(opt, device=local) as single:
        single.set_scene(scene)
        single.set_cameras(cams)
        m_f1, a_pri, a_oth, b_pri, b_oth = rc.expected_images(single, full, f1)
    for stitch in ("peer", "nccl"):
        for primary in (0, None):
            ren = RtRenderer(opt, size2, rank=rank, world=world, device=local, stitch=stitch, primary=primary, accumulate=True)
            ren.set_scene(scene, cams)
            for _ in range(f1):
                ren.render(stream.cuda_stream)
            torch.cuda.synchronize(local)
            td.barrier()
            before = ren.image().copy()
            ren.set_device_workloads([0.3] + [0.7 / (world - 1)] * (world - 1))
            ren.render(stream.cuda_stream)
            torch.cuda.synchronize(local)
            td.barrier()
            img_a = ren.image().copy()
            ren.render(stream.cuda_stream)
            torch.cuda.synchronize(local)
            td.barrier()
            img_b = ren.image().copy()
            mask = rc.primary_mask(size2, ren.dists, primary)
            bad = [rc.mismatches(before, m_f1), rc.mismatches(img_a, rc.compose(mask, a_pri, a_oth)),
                   rc.mismatches(img_b, rc.compose(mask, b_pri, b_oth))]
            # the expected blend emulates the fp32 fma through fp64 (double rounding: at most a stray value)
            same = bad[0] == 0 and bad[1] <= 2 and bad[2] == 0
            print("rank %d rebalance stitch %s primary %s: %s %s" % (rank, stitch, primary, "identical" if same else "DIFFERENT", bad), flush=True)
            ok = ok and same
            td.barrier()
            ren.close()
    return ok


if __name__ == "__main__":
    main()
