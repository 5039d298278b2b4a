#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs.

Metric : Mrays/s (and ms/frame) of the path-tracing hot path, sampler sobol-owen, 4 bounces (`max-ray-depth 4`), on the
         synthetic 1.5 M-triangle atrium (tauray_b200/synthetic.py); N GPUs shard the pixels with the reference's
         shuffled strips and the frame ends with the stitched rgba32f image on every rank.
--config c3 (default) : BASELINE configs[2], 1920x1080, 4 spp per frame, accumulation reset every frame.
         c4           : BASELINE configs[3], the same frame accumulated (no reset): `--steps 1024` is the 4096-spp image.
         c5           : BASELINE configs[4], light-field camera grid 5x9 = 45 views of 960x540, 2 spp.
Step   : one frame = all passes of the hot path over the whole image (+ stitch for N > 1).
value  : rays of all ranks / max-over-ranks CUDA-event time, scene + BVH resident in HBM.
e2e    : the same through the C ABI with HOST buffers: per step the camera table goes host->device (pinned) and the
         finished image device->host (pinned), inside the timed region. For N > 1 every rank copies its own 1/N byte
         range of the stitched image into ONE shared pinned host buffer (tr_pt_read_full_image_range).
parity : (N = 1) the GPU frame against the CPU oracle's frame of the same workload, bit for bit; `image_sha256` of the
         stitched image on every N; for N > 1 rank 0 also renders the whole frame alone and the run FAILS if the stitched
         image of any rank differs from it.
--impl reference : the reference's own implementation of this path cannot be built here (no Vulkan/glm/SDL2,
         BASELINE.md section 2); per the task contract the CPU restatement in oracle/ (kind "port") is timed on the box's
         host cores instead.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DEPTH, TRIANGLES = 4, 1_500_000
CONFIGS = {
    # name: (workload, width, height, spp per frame, views, reset accumulation every step)
    "c3": ("atrium-1.5Mtri-1920x1080-4spp-4bounce", 1920, 1080, 4, 1, True),
    "c4": ("atrium-1.5Mtri-1920x1080-4096spp-accumulation-4bounce", 1920, 1080, 4, 1, False),
    "c5": ("atrium-1.5Mtri-lightfield-5x9-960x540-2spp-4bounce", 960, 540, 2, 45, True),
}
METRICS = {
    "c3": "Mrays/s at 1920x1080 4spp 4-bounce (ms/frame in ms_per_step)",
    "c4": "Mrays/s at 1920x1080 4spp-per-frame accumulation towards 4096 spp, 4-bounce (ms/frame in ms_per_step)",
    "c5": "Mrays/s at 45 views x 960x540 2spp 4-bounce (ms/frame in ms_per_step)",
}
NCU_FULL_CSV = os.path.join(ROOT, "profiles", "r2_ncu_full_top_kernels.csv")


def bench_options(abi, spp):
    return abi.default_options(max_ray_depth=DEPTH, samples_per_pixel=spp, samples_per_pass=1,
                               local_sampler=abi.SAMPLER_SOBOL_OWEN, film=abi.FILM_POINT, min_ray_dist=1e-4,
                               mis_mode=abi.MIS_POWER_HEURISTIC, tri_light_mode=abi.TRI_LIGHT_SOLID_ANGLE)


def bench_cameras(synthetic, cfg):
    _, w, h, _, views, _ = CONFIGS[cfg]
    return synthetic.atrium_camera_grid(w / h) if views > 1 else synthetic.atrium_cameras(w / h)


def ncu_per_launch(kernel_substr, columns):
    """Per-launch averages of `columns` for the kernels whose name contains `kernel_substr`, from the committed
    `ncu --set full` capture of this same command (profiles/r2_ncu_full_top_kernels.csv; tools/ncu_table.py wrote it)."""
    try:
        import csv
        rows = list(csv.reader(open(NCU_FULL_CSV)))
        hdr, units = rows[0], rows[1]
        k = hdr.index("Kernel Name")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "inst": 1.0, "": 1.0}
        out = {}
        for c in columns:
            i = hdr.index(c)
            vals = [float(x[i].replace(",", "")) * scale.get(units[i], 1.0) for x in rows[2:] if kernel_substr in x[k] and x[i] not in ("", "n/a")]
            out[c] = sum(vals) / len(vals) if vals else None
        return out
    except Exception:
        return {c: None for c in columns}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def reference_sample(cfg):
    """The bounded sample of the workload one CPU step renders: (views used, description)."""
    _, w, h, spp, views, _ = CONFIGS[cfg]
    if cfg == "c5":
        return 5, "per step: 1 spp over the first 5 of the 45 views (%d paths = 1/18 of the 2-spp frame), same scene/options" % (5 * w * h)
    return 1, "per step: 1 spp over all %dx%d shuffled-strip slots (%d paths = 1/%d of the %d-spp frame), same scene/options" % (w, h, w * h, spp, spp)


def make_oracle(abi, distribution, synthetic, scene, cfg, spp, cores, views=None):
    from oracle.oracle_py import Oracle
    _, w, h, _, _, _ = CONFIGS[cfg]
    ora = Oracle(bench_options(abi, spp), threads=cores)
    ora.set_scene(scene)
    cams = bench_cameras(synthetic, cfg)
    ora.set_cameras(cams if views is None else cams[:views])
    slots = distribution.calculate_shuffled_strips_pixels_per_device((w, h), 1.0)
    ora.set_distribution(distribution.make_distribution((w, h), abi.DISTRIBUTION_SHUFFLED_STRIPS, 0, slots, True))
    return ora


def run_reference(args):
    """--impl reference: the CPU restatement (oracle/, kind "port") on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from tauray_b200 import abi, distribution, synthetic
    cores = os.cpu_count() or 1
    cfg = args.config
    workload = CONFIGS[cfg][0]
    scene = synthetic.atrium(TRIANGLES)
    views, sample = reference_sample(cfg)
    ora = make_oracle(abi, distribution, synthetic, scene, cfg, 1, cores, views=views if cfg == "c5" else None)
    for _ in range(args.warmup):
        ora.render()
    rays = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ora.render()
        st = ora.stats()
        rays += st["closest_rays"] + st["shadow_rays"]
    dt = time.perf_counter() - t0
    value = rays / dt / 1e6
    line = {
        "impl": "reference", "metric": METRICS[cfg], "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "triangles": TRIANGLES, "sampler": "sobol-owen", "note":
                   "reference (Vulkan RT / lavapipe raster) cannot be built in this image; CPU restatement timed instead"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def sha256_of(arr):
    return hashlib.sha256(memoryview(arr).cast("B")).hexdigest()


class SharedHostImage:
    """One pinned host buffer shared by all ranks of the node (a /dev/shm file, cudaHostRegister'ed in every process):
    every rank copies its own byte range of the stitched image into it over its own PCIe link."""

    def __init__(self, np, torch, td, rank, world, image_bytes, images=1):
        self.np, self.torch = np, torch
        self.image_bytes = image_bytes
        nbytes = image_bytes * images          # `images` host images back to back (frames in flight)
        name = [None]
        if rank == 0:
            name[0] = "/dev/shm/tauray_b200_%d_%d" % (os.getpid(), int(time.time() * 1e3) & 0xFFFFFF)
            with open(name[0], "wb") as f:
                f.truncate(nbytes)
        if td is not None:
            td.broadcast_object_list(name, src=0)
        self.path = name[0]
        self.owner = rank == 0
        self.buf = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(nbytes,))
        self.ptr = self.buf.ctypes.data
        self.registered = False
        try:
            rc = torch.cuda.cudart().cudaHostRegister(self.ptr, nbytes, 0)
            self.registered = int(rc) == 0
        except Exception:
            self.registered = False
        per = (image_bytes // world + 15) & ~15
        self.offset = min(rank * per, image_bytes)
        self.count = min(per, image_bytes - self.offset) if rank + 1 < world else image_bytes - self.offset

    def close(self):
        try:
            if self.registered:
                self.torch.cuda.cudart().cudaHostUnregister(self.ptr)
        except Exception:
            pass
        del self.buf
        if self.owner:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tauray_b200", choices=["tauray_b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--stitch", default=None, choices=[None, "peer", "nccl"])
    ap.add_argument("--kernel", default="wavefront", choices=["wavefront", "megakernel"])
    ap.add_argument("--triangles", type=int, default=TRIANGLES)
    ap.add_argument("--frames-in-flight", type=int, default=2, choices=[1, 2],
                    help="2: tr_pt_render_begin(k + 1) is issued before tr_pt_render_end(k) (two frames' traversal work queued); 1: frame by frame")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the oracle leg (cpu_baseline + parity)")
    ap.add_argument("--no-checks", action="store_true", help="N > 1: skip the single-rank image check and the NCCL-stitch comparison")
    ap.add_argument("--selftest", action="store_true", help="N > 1: run the whole multi-GPU test worker (tests/_nccl_worker.py) and exit")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "tauray_b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from tauray_b200 import abi, distribution, synthetic
    from tauray_b200.library import PathTracer
    from tauray_b200.renderer import RtRenderer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE %d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: tauray_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as td
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        td = None
    if args.selftest:
        if td is None:
            raise SystemExit("--selftest needs --gpus >= 2 under torchrun")
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _nccl_worker
        ok = _nccl_worker.run_checks(rank, world, local_rank)
        td.destroy_process_group()
        if rank == 0:
            print(json.dumps({"selftest": "ok" if ok else "FAILED", "n_gpus": world}))
        return 0 if ok else 1

    cfg = args.config
    workload, WIDTH, HEIGHT, SPP, VIEWS, reset_each_step = CONFIGS[cfg]
    size = (WIDTH, HEIGHT)
    opt = bench_options(abi, SPP)
    opt.kernel_mode = abi.KERNEL_WAVEFRONT if args.kernel == "wavefront" else abi.KERNEL_MEGAKERNEL
    t0 = time.perf_counter()
    scene = synthetic.atrium(args.triangles)
    cams = bench_cameras(synthetic, cfg)
    t_gen = time.perf_counter() - t0
    ren = RtRenderer(opt, size, rank=rank, world=world, device=local_rank, stitch=args.stitch)
    t0 = time.perf_counter()
    ren.set_scene(scene, cams)
    t_scene = time.perf_counter() - t0
    build_stats = ren.stats()          # the static scene's acceleration structure (quality build)
    # a real (non-NULL) stream: kernels, NCCL ops and the timing events all go on it
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize(local_rank)
        if td is not None:
            td.barrier()
            torch.cuda.synchronize(local_rank)

    def reduce_over_ranks(x, op):
        if td is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_over_ranks(x, td.ReduceOp.MAX if td else None)

    def sum_over_ranks(x):
        return reduce_over_ranks(x, td.ReduceOp.SUM if td else None)

    def frame(r=ren):
        if reset_each_step:
            r.reset_accumulation()
        r.render(sptr)

    # two frames in flight on the device: begin(k + 1) before end(k) (tr_pt_render_begin / tr_pt_render_end)
    fif = args.frames_in_flight == 2 and args.kernel == "wavefront"

    def frame_begin(r=ren):
        if reset_each_step:
            r.reset_accumulation()
        r.render_begin(sptr)

    def frame_end(r=ren):
        r.render_end(sptr)

    # ---- value: device-resident, CUDA events on the launching stream around exactly K frames ----
    ren.tracer.set_instrumentation(0)

    def frames(k, r=ren):
        """EXACTLY k frames, begun and ended inside the call: with frames in flight the pipeline starts empty and is
        drained (begin(0); { begin(i + 1); end(i) }), so the fill and the drain are part of what is timed."""
        if not fif:
            for _ in range(k):
                frame(r)
            return
        frame_begin(r)
        for i in range(k):
            if i + 1 < k:
                frame_begin(r)
            frame_end(r)

    frames(args.warmup)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e0.record(stream)
    frames(args.steps)
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    clock_info = clocks.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    st = ren.stats()   # counters of the last frame (every frame traces the same number of paths)
    rays_frame = st["closest_rays"] + st["shadow_rays"]
    launches = st["kernel_launches"] * args.steps
    ms_max = max_over_ranks(ms_total)
    rays_all = sum_over_ranks(float(rays_frame * args.steps))
    culled_all = sum_over_ranks(float(st["culled_shadow_rays"]))
    closest_all = sum_over_ranks(float(st["closest_rays"]))
    value = rays_all / (ms_max * 1e-3) / 1e6
    ms_per_step = ms_max / args.steps
    accumulated_spp = (args.warmup + args.steps) * SPP if not reset_each_step else SPP

    # ---- the image of one frame: checksum on every N, and (N > 1) the same frame rendered by rank 0 alone ----
    ren.reset_accumulation()
    ren.render(sptr)
    torch.cuda.synchronize(local_rank)
    image = ren.image()
    image_sha = sha256_of(image)
    st_img = ren.stats()
    img_closest = sum_over_ranks(float(st_img["closest_rays"]))
    img_shadow_incl_culled = sum_over_ranks(float(st_img["shadow_rays"] + st_img["culled_shadow_rays"]))
    stitch_check = None
    if td is not None:
        shas = [None] * world
        td.all_gather_object(shas, image_sha)
        stitch_check = {"all_ranks_equal": len(set(shas)) == 1}
        if not args.no_checks:
            single_sha = [None]
            if rank == 0:
                with PathTracer(opt, device=local_rank) as single:
                    single.set_scene(scene)
                    single.set_cameras(cams)
                    single.set_distribution(distribution.make_distribution(
                        size, abi.DISTRIBUTION_SHUFFLED_STRIPS, 0, distribution.calculate_shuffled_strips_pixels_per_device(size, 1.0), True))
                    single.render()
                    single.sync()
                    single_sha[0] = sha256_of(single.read_color())
            td.broadcast_object_list(single_sha, src=0)
            stitch_check["single_rank_sha256"] = single_sha[0]
            stitch_check["equal_single_rank"] = all(s == single_sha[0] for s in shas)

    # ---- per-kernel-class durations for the roofline: the same K frames again with one CUDA event before every
    # launch (TR_INSTRUMENT_TIMING). The events need a single stream, so this pass runs without the second stream
    # that overlaps the shadow traversal of bounce b with the closest-hit traversal of bounce b + 1.
    cls_ms = {"closest": 0.0, "shadow": 0.0, "shade": 0.0, "other": 0.0}
    cls_n = {"closest": 0, "shadow": 0, "shade": 0}
    ren.tracer.set_instrumentation(abi.INSTRUMENT_TIMING)
    for _ in range(args.steps):
        frame()
    barrier()
    st_t = ren.stats()
    for k in cls_ms:
        cls_ms[k] = st_t[k + "_ms"]
    for k in cls_n:
        cls_n[k] = st_t[k + "_launches"]

    # ---- e2e: host buffers through the C ABI, H2D cameras + D2H image inside the timed region ----
    ren.tracer.set_instrumentation(0)
    cam_host = torch.from_numpy(np.frombuffer(cams.tobytes(), dtype=np.uint8).copy()).pin_memory()
    cam_np = cam_host.numpy().view(abi.CAMERA_PAIR)
    img_bytes = VIEWS * HEIGHT * WIDTH * 16
    shared = None
    if world == 1:
        img_host = torch.empty((VIEWS, HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory()
        img_np = img_host.numpy()
    elif ren.stitch == "peer":
        shared = SharedHostImage(np, torch, td, rank, world, img_bytes, images=2)
    else:
        img_host = torch.empty((VIEWS, HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory()

    # Two frames in flight, like the reference's headless loop (MAX_FRAMES_IN_FLIGHT = 2, src/headless.cc:125-171): the
    # camera table goes up in stream order (tr_pt_set_cameras_async), the frame renders, and its staging copy runs on a
    # copy stream into one of two pinned host images beside the next frame; before a host image is reused the host waits
    # for the copy that filled it two frames ago (that is the throttle, and where a consumer would read the picture).
    # N = 1: two device targets (tr_pt_set_color_buffer) + tr_pt_read_color_async; N > 1: every rank copies its 1/N byte
    # range of the double-buffered stitched image (RtRenderer.read_back_async). Every step's camera upload and image
    # read-back are inside the timed region; the host images are checked afterwards.
    # (an accumulating run keeps ONE target: its staging copy still runs beside the next frame's traversal work when the frames
    # overlap on the device, and the next accumulation kernel waits for it on the device)
    pipelined = world == 1 and (reset_each_step or fif)
    two_targets = pipelined and reset_each_step
    in_flight = pipelined or shared is not None
    e2e_count = [0]
    copied = [None, None]
    if pipelined:
        targets = [torch.empty((VIEWS, HEIGHT, WIDTH, 4), dtype=torch.float32, device="cuda") for _ in range(2)] if two_targets else None
        hosts = [torch.empty((VIEWS, HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=local_rank)

    def e2e_step():
        if not in_flight:
            ren.tracer.set_cameras(cam_np)                  # H2D (pinned), blocking
            frame()
            if world == 1:
                ren.tracer.read_color_into(img_np.ctypes.data, img_np.nbytes)   # D2H (pinned) + sync
            elif rank == 0:
                img_host.copy_(ren._full, non_blocking=False)
            return
        k = e2e_count[0]
        e2e_count[0] += 1
        b = k & 1
        if copied[b] is not None:
            copied[b].synchronize()                         # host image b is free again (frame k - 2 has been delivered)
        if fif_e2e:
            if k + 1 < e2e_total[0]:
                e2e_issue(k + 1)                            # the NEXT frame's camera upload + traversal work ...
            frame_end()                                     # ... then this frame's accumulation / stitch
        else:
            ren.tracer.set_cameras_async(cam_np, sptr)      # H2D (pinned), in stream order
            if two_targets:
                ren.tracer.set_color_buffer(targets[b].data_ptr(), targets[b].numel() * 4)
            frame()
        if pipelined:
            ren.tracer.read_color_async(hosts[b].data_ptr(), hosts[b].numel() * 4, copy_stream.cuda_stream)   # D2H (pinned)
            copied[b] = torch.cuda.Event()
            copied[b].record(copy_stream)
        else:
            copied[b] = ren.read_back_async(shared.ptr + b * img_bytes + shared.offset, shared.offset, shared.count, sptr)

    fif_e2e = fif and in_flight

    def e2e_issue(k):
        ren.tracer.set_cameras_async(cam_np, sptr)          # H2D (pinned), in stream order
        if two_targets:
            ren.tracer.set_color_buffer(targets[k & 1].data_ptr(), targets[k & 1].numel() * 4)
        frame_begin()

    e2e_total = [0]

    def e2e_run(k):
        """EXACTLY k end-to-end steps; with frames in flight on the device the pipeline starts empty and ends drained."""
        e2e_count[0] = 0
        e2e_total[0] = k
        if fif_e2e:
            e2e_issue(0)
        for _ in range(k):
            e2e_step()
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.steps)
    barrier()                                               # synchronises the device: the last staging copy included
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = rays_all / e2e_s / 1e6
    h2d = cam_host.numel() * world
    d2h = img_bytes
    e2e_image_ok = None
    if two_targets:
        # both host images hold a finished frame of the checked camera, rendered from a clean accumulation
        e2e_image_ok = all(sha256_of(h.numpy()) == image_sha for h in hosts)
        ren.tracer.set_color_buffer(None, 0)
    elif pipelined:
        # accumulating run: the host image of the last step must be the colour target as it stands now
        e2e_image_ok = sha256_of(hosts[(args.steps - 1) & 1].numpy()) == sha256_of(ren.tracer.read_color())
    if shared is not None:
        # one more frame from a clean accumulation so that the shared host image must equal the checked one
        ren.reset_accumulation()
        ren.render(sptr)
        ren.read_back_async(shared.ptr + shared.offset, shared.offset, shared.count, sptr)
        barrier()
        if rank == 0:
            e2e_image_ok = sha256_of(shared.buf[:img_bytes]) == image_sha
            if reset_each_step:      # the other host image holds a frame of the timed loop: the same picture
                e2e_image_ok = e2e_image_ok and sha256_of(shared.buf[img_bytes:]) == image_sha
        pinned = shared.registered
        shared.close()
    else:
        pinned = True

    # ---- N > 1: the NCCL all-gather + un-shuffle stitch (the baseline the north star names) beside the fused peer stores
    stitch_compare = None
    if td is not None and not args.no_checks and ren.stitch == "peer":
        ren2 = RtRenderer(opt, size, rank=rank, world=world, device=local_rank, stitch="nccl")
        ren2.set_scene(scene, cams)
        # the same schedule as the timed loop above (frames in flight or frame by frame)
        frames(args.warmup, ren2)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        frames(args.steps, ren2)
        a1.record(stream)
        barrier()
        nccl_ms = max_over_ranks(a0.elapsed_time(a1)) / args.steps
        ren2.reset_accumulation()
        ren2.render(sptr)
        torch.cuda.synchronize(local_rank)
        same = sha256_of(ren2.image()) == image_sha
        flags = [None] * world
        td.all_gather_object(flags, same)
        stitch_compare = {"peer_ms_per_step": ms_per_step, "nccl_allgather_ms_per_step": nccl_ms, "nccl_image_equal": all(flags)}
        ren2.close()

    # ---- roofline of the dominant kernel: instrumented (visit-counting) frame, untimed ----
    ren.tracer.set_instrumentation(abi.INSTRUMENT_VISITS)
    ren.reset_accumulation()
    ren.render(sptr)
    torch.cuda.synchronize(local_rank)
    sv = ren.stats()
    ren.tracer.set_instrumentation(0)
    peak, peak_kind = measured_peaks()
    dominant = max(("closest", "shadow", "shade"), key=lambda k: cls_ms[k])
    nodes_c = sv["closest_node_visits"] / max(sv["closest_rays"], 1)
    tris_c = sv["closest_triangle_tests"] / max(sv["closest_rays"], 1)
    nodes_s = sv["shadow_node_visits"] / max(sv["shadow_rays"], 1)
    tris_s = sv["shadow_triangle_tests"] / max(sv["shadow_rays"], 1)
    # algorithmic bytes per ray: queue id 4 + ray 32 + result 16 + nodes*80 + tris*48
    b_closest = 4 + 32 + 16 + nodes_c * 80 + tris_c * 48
    b_shadow = 4 + 16 + 64 + 32 + nodes_s * 80 + tris_s * 48     # id + origin + job + accumulators RMW
    b_shade = 4 + 16 + 6 * 16 * 2 + 12 + 96 + 112 + 64             # id, hit, state r/w, indices, vertices, instance, job
    per_frame_bytes = {"closest": b_closest * st["closest_rays"], "shadow": b_shadow * st["shadow_rays"],
                       "shade": b_shade * st["closest_rays"]}
    dom_ms = cls_ms[dominant]          # per frame: the launch timer is reset by every tr_pt_render
    n_dom = max(cls_n.get(dominant, 1), 1)
    avg_launch_ms = dom_ms / n_dom
    hbm_achieved = per_frame_bytes[dominant] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    ncu_name = {"closest": "k_wf_trace<0", "shadow": "k_wf_trace<1", "shade": "k_wf_shade"}[dominant]
    ncu = ncu_per_launch(ncu_name, ["dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                                    "smsp__thread_inst_executed_per_inst_executed.ratio"]) if (world == 1 and cfg == "c3") else {}
    traffic = None
    if ncu.get("dram__bytes_read.sum") is not None and ncu.get("dram__bytes_write.sum") is not None:
        traffic = ncu["dram__bytes_read.sum"] + ncu["dram__bytes_write.sum"]
    sm_mhz = (clock_info or {}).get("sm_mhz") or 1965.0
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    issue_peak = sm_count * 4 * sm_mhz * 1e6 / 1e9                    # warp instructions per ns: 4 schedulers per SM, 1 per clock
    warp_inst = ncu.get("smsp__inst_executed.sum")
    issue_achieved = warp_inst / (avg_launch_ms * 1e-3) / 1e9 if (warp_inst and avg_launch_ms > 0) else None
    roofline = {
        # The traversal kernels are bound by instruction issue (ALU pipe), not by bytes: nodes and triangles are L1/L2 hits
        # and DRAM sits at a few % of peak. The contract's HBM figure (ALGORITHMIC bytes / launch time) is kept under "hbm".
        "bound": "issue", "kernel": {"closest": "k_wf_trace_closest", "shadow": "k_wf_trace_shadow", "shade": "k_wf_shade"}[dominant],
        "achieved": issue_achieved, "peak": issue_peak, "unit": "G warp-instructions/s",
        "frac": (issue_achieved / issue_peak) if issue_achieved else None,
        "traffic": traffic,
        "issue_note": "achieved = warp instructions per launch (smsp__inst_executed.sum of the committed ncu capture of this command, "
                      "profiles/r2_ncu_full_top_kernels.csv) / the launch duration measured live with CUDA events; peak = SMs x 4 schedulers x "
                      "the SM clock sampled during the run",
        "lanes_per_instruction": ncu.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "traffic": traffic,
                "peak_source": peak_kind,
                "note": "achieved = ALGORITHMIC bytes (DESIGN.md section 6) / CUDA-event time; most node/triangle fetches hit L1/L2, so real "
                        "DRAM traffic (`traffic`, bytes per launch from ncu) is far smaller"},
        "timing_note": "launch durations come from a second pass over the same K frames with one CUDA event before every launch (single "
                       "stream); `value` / `ms_per_step` are timed without those events, with the shadow traversal of bounce b overlapping the "
                       "closest-hit traversal of bounce b+1 on a second stream",
        "launches_per_frame": cls_n.get(dominant), "avg_launch_ms": avg_launch_ms,
        "algorithmic_bytes_per_launch": per_frame_bytes[dominant] / n_dom,
        "bytes_per_ray": {"closest": b_closest, "shadow": b_shadow}, "nodes_per_ray": {"closest": nodes_c, "shadow": nodes_s},
        "tris_per_ray": {"closest": tris_c, "shadow": tris_s},
        "class_ms_per_frame": cls_ms,
    }
    if args.kernel != "wavefront":
        roofline["kernel"] = "k_megakernel"
        roofline["traffic"] = None

    # ---- cpu_baseline + parity (rank 0, N = 1 only): the oracle renders the workload's frame (c5: a bounded sample) ----
    cpu_baseline = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        if cfg == "c5":
            views_used, sample = reference_sample(cfg)
            ora = make_oracle(abi, distribution, synthetic, scene, cfg, SPP, cores, views=views_used)
            sample = "the %d-spp frame of the first %d of the 45 views (%d paths), same scene/options" % (SPP, views_used, views_used * WIDTH * HEIGHT * SPP)
        else:
            views_used = VIEWS
            ora = make_oracle(abi, distribution, synthetic, scene, cfg, SPP, cores)
            sample = "the whole %d-spp frame of the workload (%d paths), same scene/options" % (SPP, WIDTH * HEIGHT * SPP)
        t0 = time.perf_counter()
        ora.render()
        dt = time.perf_counter() - t0
        so = ora.stats()
        cpu_baseline = {"value": (so["closest_rays"] + so["shadow_rays"]) / dt / 1e6, "unit": "Mrays/s", "cores": cores,
                        "kind": "port", "sample": sample + "; %.1f s" % dt}
        want = ora.read_color()
        got = image[:views_used]
        a, b = got.astype(np.float64), want.astype(np.float64)
        finite = np.isfinite(a) & np.isfinite(b)
        rel = np.abs(a - b)[finite] / np.maximum(np.abs(b)[finite], 1e-30)
        rel = rel[np.abs(b)[finite] > 0]
        bit_exact = got.tobytes() == want.tobytes()
        parity = {"against": "oracle (CPU restatement) rendering the same frame", "bit_exact": bool(bit_exact),
                  "max_rel": float(rel.max()) if rel.size else 0.0, "tolerance_rel": 1e-5,
                  "pixels": int(got.shape[0] * got.shape[1] * got.shape[2]), "nan_pattern_equal": bool((np.isnan(got) == np.isnan(want)).all()),
                  "ray_counts_equal": bool(views_used != VIEWS or (so["closest_rays"], so["shadow_rays"]) ==
                                           (int(img_closest), int(img_shadow_incl_culled)))}
        ora.close()

    # animated-scene update (tr_pt_update_instances: world-space vertices, tri lights and the fast radix-tree rebuild), last
    # so that every timed frame above ran on the static scene's quality build; same records, timed once after a warm-up
    inst_table = scene.instance_table()
    # first the refit (tr_pt_set_update_mode: the quality tree keeps its topology, boxes recomputed bottom-up) ...
    ren.tracer.set_update_mode(abi.UPDATE_REFIT)
    ren.update_instances(inst_table)
    torch.cuda.synchronize(local_rank)
    t0 = time.perf_counter()
    ren.update_instances(inst_table)
    torch.cuda.synchronize(local_rank)
    t_refit = time.perf_counter() - t0
    refit_stats = ren.stats()

    def frames_ms(k=3):
        frame(); barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(k):
            frame()
        ev1.record(stream)
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1)) / k
    ms_on_refitted = frames_ms()
    # ... then the rebuild
    ren.tracer.set_update_mode(abi.UPDATE_REBUILD)
    ren.update_instances(inst_table)
    torch.cuda.synchronize(local_rank)
    t0 = time.perf_counter()
    ren.update_instances(inst_table)
    torch.cuda.synchronize(local_rank)
    t_update = time.perf_counter() - t0
    update_stats = ren.stats()
    ms_on_rebuilt = frames_ms()

    failed = None
    if stitch_check and (not stitch_check["all_ranks_equal"] or stitch_check.get("equal_single_rank") is False):
        failed = "the stitched image differs between ranks or from the single-rank render"
    if parity and not (parity["bit_exact"] or parity["max_rel"] <= 1e-5):
        failed = "the GPU frame differs from the oracle's beyond 1e-5 relative"
    if e2e_image_ok is False:
        failed = "the image gathered into the shared host buffer differs from the stitched image"

    if rank == 0:
        line = {
            "metric": METRICS[cfg], "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload, "triangles": args.triangles, "sampler": "sobol-owen", "kernel": args.kernel,
                       "views": VIEWS, "size": [WIDTH, HEIGHT], "spp_per_frame": SPP, "accumulated_spp_at_end": accumulated_spp,
                       "frames_in_flight": ("2: tr_pt_render_begin(k + 1) is issued before tr_pt_render_end(k); the timed region begins and ends exactly K frames (pipeline fill and drain included)"
                                            if fif else "1: frame by frame"),
                       "rays_definition": "traversal queries actually traced (closest + shadow); NEE shadow rays whose BSDF weight is "
                                          "exactly 0 are proven moot and not traced",
                       "culled_shadow_rays_per_frame": culled_all,
                       "distribution": "shuffled-strips", "stitch": ren.stitch, "rays_per_frame": rays_all / args.steps,
                       "l2": "inputs exceed L2: BVH+triangles+vertices ~%.0f MB, wavefront state ~%.0f MB per rank" % (
                           (st["bvh_node_count"] * 80 + st["triangle_count"] * 48 * 2 + st["triangle_count"] * 12) / 1e6,
                           WIDTH * HEIGHT * VIEWS / world * min(SPP, 8) * 236 / 1e6),
                       "bvh_builder": "binned SAH, on device (static scene); radix tree for per-frame updates",
                       "bvh_build_ms": build_stats["bvh_build_ms"], "bvh_nodes": build_stats["bvh_node_count"], "bvh_depth": build_stats["bvh_depth"],
                       "bvh_update_rebuild_ms": update_stats["bvh_build_ms"],
                       "bvh_update_refit_ms": refit_stats["bvh_build_ms"], "scene_update_refit_ms": 1e3 * t_refit,
                       "ms_per_frame_after_refit": ms_on_refitted, "ms_per_frame_after_rebuild": ms_on_rebuilt,
                       "scene_upload_s": t_scene, "scene_update_ms": 1e3 * t_update, "scene_generate_s": t_gen, "wall_s_timed": wall},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
            "image_sha256": image_sha,
            "stitch_check": stitch_check,
            "stitch_compare": stitch_compare,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "readback": (("two frames in flight: two device targets, staging copies on a copy stream into two pinned host images (tr_pt_read_color_async)"
                                  if two_targets else "two frames in flight: one accumulating device target, staging copies on a copy stream into two pinned host images (tr_pt_read_color_async)")
                                 if pipelined else "one pinned host image") if world == 1 else
                                ("two frames in flight: every rank copies its 1/N byte range of the double-buffered stitched image into one of two shared pinned host images on a copy stream%s" % ("" if pinned else " (NOT pinned: cudaHostRegister failed)")
                                 if ren.stitch == "peer" else "rank 0 copies the stitched image"),
                    "host_image_equal": e2e_image_ok},
            "gpu_launches": int(launches),
            "clocks": clock_info,
        }
        if failed:
            line["failed"] = failed
        print(json.dumps(line))
    ren.close()
    if td is not None:
        td.destroy_process_group()
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
