"""Host mirror of Tauray's multi-device renderer for the path tracer: `rt_renderer<path_tracer_stage>`
(src/rt_renderer.{hh,cc}) + `stitch_stage` + `device_transfer` + `load_balancer`.

One process per GPU (torch.distributed). Every rank owns one `PathTracer` (= one
path_tracer_stage), renders its shuffled-strips range [index, index + count) for all views and all
samples, and the frame ends with the stitched rgba32f image resident on EVERY rank:

  stitch="peer"  (default for world > 1): the tracer's last accumulation pass stores each finished
                 pixel straight into every rank's full image over NVLink peer memory (images
                 exchanged once as CUDA IPC handles) -- render, gather and un-shuffle in one kernel;
                 the frame is closed by a device-side barrier (a 1-element NCCL all-reduce).
  stitch="nccl"  the portable baseline: ncclAllGather of the strips + one un-shuffle kernel per
                 source rank (stitch_shuffled_strips.comp:21-63).

Reference call sequence mirrored by `render()`: src/rt_renderer.cc:84-133; workload split:
src/rt_renderer.cc:136-183,245-279; load balancing: src/load_balancer.cc:12-32.
"""
import numpy as np

from . import abi
from . import distribution as dist_mod


class LoadBalancer:
    """load_balancer (src/load_balancer.{hh,cc}): EMA 0.1 towards the normalised device speed."""

    def __init__(self, initial_weights):
        self.workloads = self.normalize_workloads(list(initial_weights))

    @staticmethod
    def normalize_workloads(w):
        s = float(sum(w))
        add = 0.0
        if s == 0:
            add = 1.0
            s = float(len(w))
        return [(max(x, 0.0) + add) / s for x in w]

    def update(self, times_ms):
        speeds = [max(w / t, 0.0) if t > 0 else 0.0 for w, t in zip(self.workloads, times_ms)]
        total = sum(speeds)
        if total > 0 and np.isfinite(total):
            self.workloads = [w * 0.9 + (s / total) * 0.1 for w, s in zip(self.workloads, speeds)]
        return self.workloads


class RtRenderer:
    """`renderer` interface of the reference (src/renderer.hh:9-20): set_scene, reset_accumulation,
    render, set_device_workloads -- for one rank of a `world`-GPU job."""

    def __init__(self, options, size, rank=0, world=1, device=0, stitch=None,
                 strategy=abi.DISTRIBUTION_SHUFFLED_STRIPS, tracer_factory=None, process_group=None,
                 backend_is_cuda=True):
        self.options = abi.copy_options(options)
        self.size = (int(size[0]), int(size[1]))
        self.rank, self.world, self.device = rank, world, device
        self.strategy = strategy
        self.cuda = backend_is_cuda
        self.pg = process_group
        if stitch is None:
            stitch = "none" if world == 1 else ("peer" if backend_is_cuda else "nccl")
        if stitch not in ("none", "peer", "nccl"):
            raise ValueError("stitch must be none|peer|nccl")
        if world > 1 and strategy != abi.DISTRIBUTION_SHUFFLED_STRIPS:
            raise NotImplementedError("multi-device stitching implements shuffled strips (the reference's default)")
        self.stitch = stitch
        if tracer_factory is None:
            from .library import PathTracer
            tracer_factory = lambda opt: PathTracer(opt, device=device)  # noqa: E731
        self.tracer = tracer_factory(self.options)
        self.views = 0
        self.workloads = dist_mod.even_workloads(world)
        self.dists = None
        self._strip = None
        self._gathered = None
        self._full = None
        self._full_ptr = None
        self._peer_ptrs = []
        self._flag = None

    # -- reference interface ---------------------------------------------------------------
    def set_scene(self, scene, cameras):
        self.tracer.set_scene(scene)
        self.tracer.set_cameras(cameras)
        self.views = len(cameras)
        self._apply_distribution()

    def reset_accumulation(self, reset_sample_counter=True):
        self.tracer.reset_accumulation(reset_sample_counter)

    def set_device_workloads(self, ratios):
        """rt_renderer::set_device_workloads (src/rt_renderer.cc:136-183)."""
        assert len(ratios) == self.world
        self.workloads = list(ratios)
        self._apply_distribution()

    # -- distribution / buffers ------------------------------------------------------------
    def _apply_distribution(self):
        if self.world == 1:
            count = dist_mod.calculate_shuffled_strips_pixels_per_device(self.size, 1.0) \
                if self.strategy == abi.DISTRIBUTION_SHUFFLED_STRIPS else 1
            self.dists = [dist_mod.make_distribution(self.size, self.strategy, 0, count, True)]
        else:
            self.dists = dist_mod.device_distributions(self.size, self.strategy, self.workloads, primary_device=None)
        self.tracer.set_distribution(self.dists[self.rank])
        if self.world > 1 and self.views:
            self._alloc_stitch_buffers()

    def _alloc_stitch_buffers(self):
        import torch
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        W, H = self.size
        self.max_count = max(d.count for d in self.dists)
        my = self.dists[self.rank]
        if self.cuda:
            # strip target: caller-owned torch memory so NCCL can gather straight out of it
            self._strip = torch.zeros((self.views, my.count, 4), dtype=torch.float32, device=dev)
            self.tracer.set_color_buffer(self._strip.data_ptr(), self._strip.numel() * 4)
        if self.stitch == "nccl":
            self._send = torch.zeros((self.views, self.max_count, 4), dtype=torch.float32, device=dev)
            self._gathered = torch.zeros((self.world, self.views, self.max_count, 4), dtype=torch.float32, device=dev)
            self._full = torch.zeros((self.views, H, W, 4), dtype=torch.float32, device=dev)
            self._full_ptr = self._full.data_ptr() if self.cuda else None
        elif self.stitch == "peer":
            import torch.distributed as td
            ptr, nbytes = self.tracer.alloc_full_image()
            self._full_ptr, self._full_bytes = ptr, nbytes
            handles = [None] * self.world
            td.all_gather_object(handles, self.tracer.ipc_get_handle(), group=self.pg)
            for p in self._peer_ptrs:
                self.tracer.ipc_close_handle(p)
            self._peer_ptrs = []
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(ptr)
                else:
                    p = self.tracer.ipc_open_handle(h)
                    self._peer_ptrs.append(p)
                    ptrs.append(p)
            self.tracer.set_peer_images(ptrs)
            self._flag = torch.zeros(1, dtype=torch.float32, device=dev)

    # -- frame -----------------------------------------------------------------------------
    def render(self, stream=None):
        """rt_renderer::render (src/rt_renderer.cc:84-133): trace, then transfer + stitch."""
        if stream is None and self.cuda and self.world > 1:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream  # order with the NCCL ops
        self.tracer.render(stream)
        if self.world == 1 or self.stitch == "none":
            return
        import torch.distributed as td
        if self.stitch == "peer":
            # all peers' stores into my image are complete once every rank's kernels finished:
            # a 1-element all-reduce on the same stream is the frame barrier.
            td.all_reduce(self._flag, group=self.pg)
            return
        my = self.dists[self.rank]
        if self.cuda:
            self._send[:, :my.count].copy_(self._strip)
        else:
            import torch
            self._send[:, :my.count].copy_(torch.from_numpy(self.tracer.read_color()))
        # concatenated-along-dim-0 form: accepted by both NCCL and gloo
        td.all_gather_into_tensor(self._gathered.view(self.world * self.views, self.max_count, 4), self._send, group=self.pg)
        for r, d in enumerate(self.dists):
            src = self._gathered[r, :, :d.count].contiguous() if d.count != self.max_count else self._gathered[r]
            if self.cuda:
                self.tracer.unshuffle(src.data_ptr(), d.index, d.count, self._full_ptr, stream)
            else:
                self._unshuffle_host(src.numpy(), d.index, d.count)

    def _unshuffle_host(self, strip, index, count):
        """stitch_shuffled_strips.comp:21-63 on the host (gloo tests)."""
        W, H = self.size
        j = dist_mod.permute_region_id(np.arange(index, index + count, dtype=np.uint64), self.size)
        ok = j < W * H
        full = self._full.numpy().reshape(self.views, W * H, 4)
        full[:, j[ok].astype(np.int64)] = strip[:, ok]

    # -- results ---------------------------------------------------------------------------
    def image(self):
        """The stitched image as float32 (views, H, W, 4) on the host (headless staging copy)."""
        W, H = self.size
        if self.world == 1 or self.stitch == "none":
            out = self.tracer.read_color()
            return out
        if self.stitch == "nccl":
            return self._full.cpu().numpy()
        out = np.empty((self.views, H, W, 4), dtype=np.float32)
        return self.tracer.read_full_image(out)

    def full_image_ptr(self):
        return self._full_ptr

    def stats(self):
        return self.tracer.stats()

    def close(self):
        for p in self._peer_ptrs:
            try:
                self.tracer.ipc_close_handle(p)
            except Exception:
                pass
        self._peer_ptrs = []
        self.tracer.close()
