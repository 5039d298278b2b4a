#!/usr/bin/env python3
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

Metric : Mrays/s (and ms/frame) of the path-tracing hot path at 1920x1080, 4 spp, 4 bounces
         (`max-ray-depth 4`), sampler sobol-owen, on the synthetic 1.5 M-triangle atrium
         (tauray_b200/synthetic.py) -- BASELINE.json configs[2]; N GPUs shard the pixels with the
         reference's shuffled strips and the frame ends with the stitched rgba32f image on every rank.
Step   : one frame = all 4 passes of the hot path over the whole image (+ stitch for N > 1).
value  : rays of all ranks / max-over-ranks CUDA-event time, scene + BVH resident in HBM.
e2e    : the same through the C ABI with HOST buffers: per step the camera table goes host->device
         (pinned) and the finished image device->host (pinned), inside the timed region.
--impl reference : the reference's own implementation of this path cannot be built here (no
         Vulkan/glm/SDL2); per the task contract the CPU restatement in oracle/
         (kind "port") is timed on the box's host cores instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SPP, DEPTH, TRIANGLES = 1920, 1080, 4, 4, 1_500_000
WORKLOAD = "atrium-1.5Mtri-1920x1080-4spp-4bounce"
METRIC = "Mrays/s at 1920x1080 4spp 4-bounce (ms/frame in ms_per_step)"


def bench_options(abi, spp=SPP):
    return abi.default_options(max_ray_depth=DEPTH, samples_per_pixel=spp, samples_per_pass=1,
                               local_sampler=abi.SAMPLER_SOBOL_OWEN, film=abi.FILM_POINT, min_ray_dist=1e-4,
                               mis_mode=abi.MIS_POWER_HEURISTIC, tri_light_mode=abi.TRI_LIGHT_SOLID_ANGLE)


def ncu_traffic_per_launch(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r1_ncu_full_top_kernels.csv; same command line as this bench)."""
    path = os.path.join(ROOT, "profiles", "r1_ncu_full_top_kernels.csv")
    try:
        import csv
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        k, r, w = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        vals = [float(x[r]) * scale.get(units[r], 1.0) + float(x[w]) * scale.get(units[w], 1.0)
                for x in rows[2:] if kernel_substr in x[k]]
        return sum(vals) / len(vals) if vals else None
    except Exception:
        return None


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


def run_reference(args):
    """--impl reference: the CPU restatement (oracle/, kind "port") on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from tauray_b200 import abi, distribution, synthetic
    from oracle.oracle_py import Oracle
    cores = os.cpu_count() or 1
    opt = bench_options(abi, spp=1)
    scene = synthetic.atrium(TRIANGLES)
    ora = Oracle(opt, threads=cores)
    ora.set_scene(scene)
    ora.set_cameras(synthetic.atrium_cameras(WIDTH / HEIGHT))
    size = (WIDTH, HEIGHT)
    slots = distribution.calculate_shuffled_strips_pixels_per_device(size, 1.0)
    ora.set_distribution(distribution.make_distribution(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, 0, slots, True))
    # a step = 1 spp over the whole image = 1/4 of the 4-spp frame
    sample_slots = slots
    for _ in range(args.warmup):
        ora.render(sample_slots)
    rays = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ora.render(sample_slots)
        st = ora.stats()
        rays += st["closest_rays"] + st["shadow_rays"]
    dt = time.perf_counter() - t0
    value = rays / dt / 1e6
    sample = "per step: 1 spp over all 1920x1080 shuffled-strip slots (%d paths = 1/4 of the 4-spp frame), same scene/options" % sample_slots
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "triangles": TRIANGLES, "sampler": "sobol-owen", "note":
                   "reference (Vulkan RT / lavapipe raster) cannot be built in this image; CPU restatement timed instead"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tauray_b200", choices=["tauray_b200", "reference"])
    ap.add_argument("--stitch", default=None, choices=[None, "peer", "nccl"])
    ap.add_argument("--kernel", default="wavefront", choices=["wavefront", "megakernel"])
    ap.add_argument("--triangles", type=int, default=TRIANGLES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "tauray_b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from tauray_b200 import abi, distribution, synthetic
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

    size = (WIDTH, HEIGHT)
    opt = bench_options(abi)
    opt.kernel_mode = abi.KERNEL_WAVEFRONT if args.kernel == "wavefront" else abi.KERNEL_MEGAKERNEL
    t0 = time.perf_counter()
    scene = synthetic.atrium(args.triangles)
    cams = synthetic.atrium_cameras(WIDTH / HEIGHT)
    t_gen = time.perf_counter() - t0
    ren = RtRenderer(opt, size, rank=rank, world=world, device=local_rank, stitch=args.stitch)
    t0 = time.perf_counter()
    ren.set_scene(scene, cams)
    t_scene = time.perf_counter() - t0
    # animated-scene update (tr_pt_update_instances): same records, so the image is unchanged; timed once
    inst_table = scene.instance_table()
    ren.update_instances(inst_table)
    t0 = time.perf_counter()
    ren.update_instances(inst_table)
    t_update = time.perf_counter() - t0
    # a real (non-NULL) stream: kernels, NCCL ops and the timing events all go on it
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize(local_rank)
        if td is not None:
            td.barrier()
            torch.cuda.synchronize(local_rank)

    def max_over_ranks(x):
        if td is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if td is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return float(t.item())

    # ---- value: device-resident, CUDA events on the launching stream around exactly K frames ----
    ren.tracer.set_instrumentation(0)
    for _ in range(args.warmup):
        ren.reset_accumulation()
        ren.render(sptr)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rays_total = 0
    launches = 0
    cls_ms = {"closest": 0.0, "shadow": 0.0, "shade": 0.0, "other": 0.0}
    cls_n = {"closest": 0, "shadow": 0, "shade": 0}
    frame_stats = []
    wall0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        ren.reset_accumulation()
        ren.render(sptr)
        frame_stats.append(None)
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    clock_info = clocks.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    st = ren.stats()   # counters of the last frame (every frame traces the same rays)
    rays_frame = st["closest_rays"] + st["shadow_rays"]
    rays_total = rays_frame * args.steps
    launches = st["kernel_launches"] * args.steps
    ms_max = max_over_ranks(ms_total)
    rays_all = sum_over_ranks(float(rays_total))
    culled_all = sum_over_ranks(float(st["culled_shadow_rays"]))
    value = rays_all / (ms_max * 1e-3) / 1e6
    ms_per_step = ms_max / args.steps
    # ---- per-kernel-class durations for the roofline: the same K frames again with one CUDA event before every
    # launch (TR_INSTRUMENT_TIMING). The events need a single stream, so this pass runs without the second stream
    # that overlaps the shadow traversal of bounce b with the closest-hit traversal of bounce b + 1.
    ren.tracer.set_instrumentation(abi.INSTRUMENT_TIMING)
    for _ in range(args.steps):
        ren.reset_accumulation()
        ren.render(sptr)
    barrier()
    st_t = ren.stats()
    for k in cls_ms:
        cls_ms[k] = st_t[k + "_ms"]
    for k in cls_n:
        cls_n[k] = st_t[k + "_launches"]

    # ---- e2e: host buffers through the C ABI, H2D cameras + D2H image inside the timed region ----
    ren.tracer.set_instrumentation(0)
    cam_host = torch.from_numpy(np.frombuffer(cams.tobytes(), dtype=np.uint8).copy()).pin_memory()
    img_host = torch.empty((len(cams), HEIGHT, WIDTH, 4), dtype=torch.float32).pin_memory()
    img_np = img_host.numpy()
    cam_np = cam_host.numpy().view(abi.CAMERA_PAIR)

    def e2e_step():
        ren.tracer.set_cameras(cam_np)                      # H2D (pinned)
        ren.reset_accumulation()
        ren.render(sptr)
        if world == 1:
            ren.tracer.read_color_into(img_np.ctypes.data, img_np.nbytes)   # D2H (pinned) + sync
        elif ren.stitch == "peer":
            ren.tracer.read_full_image(img_np)
        else:
            img_host.copy_(ren._full, non_blocking=False)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = rays_all / e2e_s / 1e6
    h2d = cam_host.numel()
    d2h = img_host.numel() * 4

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
    dom_ms = cls_ms[dominant]
    achieved = per_frame_bytes[dominant] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": {"closest": "k_wf_trace_closest", "shadow": "k_wf_trace_shadow", "shade": "k_wf_shade"}[dominant],
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_kind,
        "note": "achieved = ALGORITHMIC bytes (DESIGN.md section 6) / CUDA-event time; most node/triangle fetches hit L1/L2, so real "
                "DRAM traffic (`traffic`, bytes per launch from ncu) is far smaller: the kernel is bound by the ALU pipe, issue slots and warp "
                "divergence. Launch durations come from a second pass over the same K frames with one CUDA event before every launch "
                "(single stream); `value` / `ms_per_step` are timed without those events, with the shadow traversal of bounce b "
                "overlapping the closest-hit traversal of bounce b+1 on a second stream",
        "launches_per_frame": cls_n.get(dominant), "avg_launch_ms": dom_ms / max(cls_n.get(dominant, 1), 1),
        "bytes_per_ray": {"closest": b_closest, "shadow": b_shadow}, "nodes_per_ray": {"closest": nodes_c, "shadow": nodes_s},
        "tris_per_ray": {"closest": tris_c, "shadow": tris_s},
        "class_ms_per_frame": cls_ms,
    }
    ncu_name = {"closest": "k_wf_trace<0", "shadow": "k_wf_trace<1", "shade": "k_wf_shade"}[dominant]
    roofline["traffic"] = ncu_traffic_per_launch(ncu_name) if world == 1 else None
    roofline["algorithmic_bytes_per_launch"] = per_frame_bytes[dominant] / max(cls_n.get(dominant, 1), 1)
    if args.kernel != "wavefront":
        roofline["kernel"] = "k_megakernel"
        roofline["traffic"] = None

    # ---- cpu_baseline (rank 0, N = 1 only): the oracle on a bounded sample ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.oracle_py import Oracle
        cores = os.cpu_count() or 1
        ora = Oracle(bench_options(abi, spp=1), threads=cores)
        ora.set_scene(scene)
        ora.set_cameras(cams)
        slots = distribution.calculate_shuffled_strips_pixels_per_device(size, 1.0)
        ora.set_distribution(distribution.make_distribution(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, 0, slots, True))
        sample_slots = slots
        t0 = time.perf_counter()
        ora.render(sample_slots)
        dt = time.perf_counter() - t0
        so = ora.stats()
        cpu_baseline = {"value": (so["closest_rays"] + so["shadow_rays"]) / dt / 1e6, "unit": "Mrays/s", "cores": cores,
                        "kind": "port", "sample": "1 spp over all shuffled-strip slots = 1/4 of the 4-spp frame (%d paths), same scene/options; %.1f s" % (sample_slots, dt)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "triangles": args.triangles, "sampler": "sobol-owen", "kernel": args.kernel,
                       "rays_definition": "traversal queries actually traced (closest + shadow); NEE shadow rays whose BSDF weight is "
                                          "exactly 0 are proven moot and not traced",
                       "culled_shadow_rays_per_frame": culled_all,
                       "distribution": "shuffled-strips", "stitch": ren.stitch, "rays_per_frame": rays_all / args.steps,
                       "l2": "inputs exceed L2: BVH+triangles+vertices ~%.0f MB, wavefront state ~%.0f MB per rank" % (
                           (st["bvh_node_count"] * 80 + st["triangle_count"] * 48 * 2 + st["triangle_count"] * 12) / 1e6,
                           WIDTH * HEIGHT / world * SPP * 236 / 1e6),
                       "bvh_build_ms": st["bvh_build_ms"], "bvh_nodes": st["bvh_node_count"], "bvh_depth": st["bvh_depth"],
                       "scene_upload_s": t_scene, "scene_update_ms": 1e3 * t_update, "scene_generate_s": t_gen, "wall_s_timed": wall},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": int(launches),
            "clocks": clock_info,
        }
        print(json.dumps(line))
    ren.close()
    if td is not None:
        td.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
