"""Host mirror of Tauray's multi-device renderer for the path tracer: `rt_renderer<path_tracer_stage>`
(src/rt_renderer.{hh,cc}) + `stitch_stage` + `device_transfer` + `load_balancer`.

One process per GPU (torch.distributed). Every rank owns one `PathTracer` (= one
path_tracer_stage), renders its shuffled-strips range [index, index + count) for all views and all
samples, and the frame ends with the stitched rgba32f image resident on EVERY rank:

  stitch="peer"  (default for world > 1): the tracer's last accumulation pass stores each finished
                 pixel straight into every rank's full image over NVLink peer memory (images
                 exchanged once as CUDA IPC handles) -- render, gather and un-shuffle in one kernel;
                 the frame is closed by a device-side barrier (a 1-element NCCL all-reduce). The
                 full images are double-buffered (frames alternate between two halves), so a fast
                 rank's next frame never stores into the image a slower rank is still reading; the
                 enabled aux G-buffer layers are stitched the same way.
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
                 backend_is_cuda=True, primary=None, accumulate=False):
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
        if world > 1 and strategy == abi.DISTRIBUTION_DUPLICATE:
            raise NotImplementedError("the duplicate distribution renders the whole image on every device: nothing to stitch")
        self.stitch = stitch
        # `primary`: the rank that plays the reference's display device -- it renders in place into its full
        # image and keeps accumulating across workload changes (src/rt_renderer.cc:165-175). None: no rank does
        # (headless; every rank renders a strip). `accumulate` = rt_renderer::options::accumulate.
        self.primary = primary if world > 1 else None
        self.accumulate = bool(accumulate)
        self.accumulated_frames = 0
        self._blend = 1.0
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
        self._copy_stream = None          # read_back_async
        self._copy_done = None
        self._full_key = None
        self.layers = []

    def set_gbuffer_layers(self, layers):
        """The other gbuffer_target entries (diffuse, reflection, albedo, ...): rendered by every rank and, for the
        peer stitch, gathered into every rank's full image like the colour (src/rt_renderer.cc:370-397). Collective."""
        self.layers = [int(l) for l in layers]
        self.tracer.set_gbuffer_layers(self.layers)
        if self.world > 1 and self.views and self.dists is not None:
            self._alloc_stitch_buffers()

    def layer_image(self, layer):
        """The stitched aux layer (views, H, W, channels) on the host."""
        if self.world == 1 or self.stitch == "none":
            W, H = self.size
            return self.tracer.read_gbuffer_layer(layer).reshape(self.views, H, W, -1)
        if self.stitch != "peer":
            raise NotImplementedError("aux layers are stitched by the peer stitch only")
        return self.tracer.read_full_layer(layer, self.size, self.views)

    # -- reference interface ---------------------------------------------------------------
    def set_scene(self, scene, cameras):
        self.tracer.set_scene(scene)
        self.tracer.set_cameras(cameras)
        self.views = len(cameras)
        self._apply_distribution()

    def update_instances(self, instances):
        """scene_stage::update for an animated scene (src/scene_stage.cc:1572-1675): new instance records ->
        world-space vertices, tri lights and the acceleration structure, rebuilt on the device."""
        self.tracer.update_instances(instances)

    def reset_accumulation(self, reset_sample_counter=True):
        """rt_renderer::reset_accumulation (src/rt_renderer.cc:66-81)."""
        self.tracer.reset_accumulation(reset_sample_counter)
        self.accumulated_frames = 0
        self._set_blend(1.0)

    def set_device_workloads(self, ratios):
        """rt_renderer::set_device_workloads (src/rt_renderer.cc:136-183): new strips for everyone; only the
        primary device renders in place, so every other device restarts its accumulation and its next strip is
        blended into the stitched image with 1/(accumulated_frames+1) instead."""
        assert len(ratios) == self.world
        if self.strategy != abi.DISTRIBUTION_SHUFFLED_STRIPS:
            return
        cumulative, clamped = 0.0, []
        for r in ratios:
            r = min(max(float(r), 0.0), 1.0 - cumulative)
            clamped.append(r)
            cumulative += r
        self.workloads = clamped
        self._apply_distribution()
        if self.world > 1:
            if self.rank != self.primary:
                self.tracer.reset_accumulation(False)
            if self.accumulate:
                self._set_blend(1.0 / (self.accumulated_frames + 1))

    def _set_blend(self, ratio):
        self._blend = float(ratio)
        if self.world > 1 and self.stitch == "peer" and hasattr(self.tracer, "set_stitch_blend"):
            # the primary's pixels are written in place and never blended (stitch_stage.cc:160: img_idx != primary_index)
            self.tracer.set_stitch_blend(1.0 if self.rank == self.primary else self._blend)

    # -- distribution / buffers ------------------------------------------------------------
    def _apply_distribution(self):
        if self.world == 1:
            count = dist_mod.calculate_shuffled_strips_pixels_per_device(self.size, 1.0) \
                if self.strategy == abi.DISTRIBUTION_SHUFFLED_STRIPS else 1
            self.dists = [dist_mod.make_distribution(self.size, self.strategy, 0, count, True)]
        else:
            self.dists = dist_mod.device_distributions(self.size, self.strategy, self.workloads, primary_device=self.primary)
        self.tracer.set_distribution(self.dists[self.rank])
        if self.world > 1 and self.views:
            self._alloc_stitch_buffers()

    def _alloc_stitch_buffers(self):
        import torch
        dev = torch.device("cuda", self.device) if self.cuda else torch.device("cpu")
        W, H = self.size
        # pixels per view in each device's own target: `count` strip slots, or this device's scanlines
        self.pixels = [int(np.prod(dist_mod.get_ray_count(d))) for d in self.dists]
        self.max_count = max(self.pixels)
        my = self.dists[self.rank]
        my_pixels = self.pixels[self.rank]
        in_place = self.rank == self.primary          # the display device renders straight into the stitched image
        if self.cuda and not in_place and self.stitch == "nccl":
            # strip target: caller-owned torch memory so NCCL can gather straight out of it
            self._strip = torch.zeros((self.views, my_pixels, 4), dtype=torch.float32, device=dev)
            self.tracer.set_color_buffer(self._strip.data_ptr(), self._strip.numel() * 4)
        elif self.cuda and self.stitch == "peer":
            self.tracer.set_color_buffer(None, 0)
        if self.stitch == "nccl":
            self._send = torch.zeros((self.views, self.max_count, 4), dtype=torch.float32, device=dev)
            self._gathered = torch.zeros((self.world, self.views, self.max_count, 4), dtype=torch.float32, device=dev)
            if self._full is None or tuple(self._full.shape) != (self.views, H, W, 4):
                self._full = torch.zeros((self.views, H, W, 4), dtype=torch.float32, device=dev)
            self._full_ptr = self._full.data_ptr() if self.cuda else None
            if self.cuda and in_place:
                self.tracer.set_color_buffer(self._full_ptr, self._full.numel() * 4)
        elif self.stitch == "peer":
            import torch.distributed as td
            ptr, nbytes = self.tracer.alloc_full_image()
            info = self.tracer.full_image_info()
            # same size and layers => the library keeps its allocation on every rank: the (collective) handle
            # exchange is needed only when the arena was reallocated
            if (ptr, info.generation) != self._full_key:
                self._full_key = (ptr, info.generation)
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
                self.tracer.set_stitch_double_buffer(True)
                self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
            # the display device renders in place into its (double-buffered) full image
            self.tracer.set_render_in_full_image(in_place)

    # -- frame -----------------------------------------------------------------------------
    def render(self, stream=None):
        """rt_renderer::render (src/rt_renderer.cc:84-133): trace, then transfer + stitch."""
        if stream is None and self.cuda and self.world > 1:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream  # order with the NCCL ops
        in_place = self.world > 1 and self.rank == self.primary
        if in_place and not self.cuda and hasattr(self.tracer, "write_color"):
            self.tracer.write_color(self._full.numpy())      # CPU stand-in for "the target is the stitched image"
        self.tracer.render(stream)
        if self.world > 1 and self.stitch != "none":
            if self.cuda and stream is not None:
                # the torch / NCCL work of the stitch goes on the stream the frame was rendered on
                import torch
                with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=torch.device("cuda", self.device))):
                    self._stitch(stream, in_place)
            else:
                self._stitch(stream, in_place)
            # "Reset temporary blending from stitching" (src/rt_renderer.cc:126-127)
            if self._blend != 1.0:
                self._set_blend(1.0)
        self.accumulated_frames += 1

    def render_begin(self, stream=None):
        """Frames in flight on the device (tr_pt_render_begin): camera rays, traversal and shading of the NEXT frame on a
        stream of the tracer's own, behind what `stream` holds now. Call order begin(0); { begin(k + 1); end(k) }."""
        if stream is None and self.cuda and self.world > 1:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        self.tracer.render_begin(stream)

    def render_end(self, stream=None):
        """The rest of the oldest begun frame on `stream`: accumulation + peer stores (tr_pt_render_end), then the frame's
        barrier / stitch exactly like render(). The next frame's stores into the stitched images come behind this barrier
        on the same stream, so the double-buffer argument of render() holds unchanged."""
        if stream is None and self.cuda and self.world > 1:
            import torch
            stream = torch.cuda.current_stream(self.device).cuda_stream
        in_place = self.world > 1 and self.rank == self.primary
        self.tracer.render_end(stream)
        if self.world > 1 and self.stitch != "none":
            if self.cuda and stream is not None:
                import torch
                with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=torch.device("cuda", self.device))):
                    self._stitch(stream, in_place)
            else:
                self._stitch(stream, in_place)
            if self._blend != 1.0:
                self._set_blend(1.0)
        self.accumulated_frames += 1

    def _stitch(self, stream, in_place):
        import torch
        import torch.distributed as td
        if self.stitch == "peer":
            # an asynchronous read-back of the previous frame (read_back_async) must have left its half of the image
            # before this rank reports the frame done: past this barrier the peers start storing into that half again
            if self._copy_done is not None:
                torch.cuda.current_stream(self.device).wait_event(self._copy_done)
                self._copy_done = None
            # all peers' stores into my image are complete once every rank's kernels finished:
            # a 1-element all-reduce on the same stream is the frame barrier.
            td.all_reduce(self._flag, group=self.pg)
            return
        my = self.dists[self.rank]
        my_pixels = self.pixels[self.rank]
        if not in_place:
            if self.cuda:
                self._send[:, :my_pixels].copy_(self._strip)
            else:
                self._send[:, :my_pixels].copy_(torch.from_numpy(self.tracer.read_color()).reshape(self.views, my_pixels, 4))
        elif not self.cuda:
            # the primary's pixels, rendered in place in the tracer's own full-size target
            img = torch.from_numpy(self.tracer.read_color()).reshape(self.views, -1, 4)
            j = torch.from_numpy(self._pixel_ids(my))
            self._full.view(self.views, -1, 4)[:, j] = img[:, j]
        # concatenated-along-dim-0 form: accepted by both NCCL and gloo
        td.all_gather_into_tensor(self._gathered.view(self.world * self.views, self.max_count, 4), self._send, group=self.pg)
        if self.primary is None or in_place:
            for r, d in enumerate(self.dists):
                if r == self.primary:
                    continue                              # in place already, never blended (stitch_stage.cc:160)
                n = self.pixels[r]
                src = self._gathered[r, :, :n].contiguous() if n != self.max_count else self._gathered[r]
                if self.cuda:
                    self.tracer.unshuffle(src.data_ptr(), d.index, d.count, self._full_ptr, stream, self._blend)
                else:
                    self._unshuffle_host(src.numpy(), d.index, d.count, self._blend)
        if self.primary is not None:
            # only the display device stitches in the reference; every rank gets the image here
            td.broadcast(self._full, src=self.primary, group=self.pg)

    def _pixel_ids(self, d):
        """Linear pixel index (y * W + x) of every pixel device distribution `d` renders, in target order."""
        W, H = self.size
        if self.strategy == abi.DISTRIBUTION_SCANLINE:
            rows = np.arange(d.index, H, d.count, dtype=np.int64)
            return (rows[:, None] * W + np.arange(W, dtype=np.int64)[None, :]).reshape(-1)
        j = dist_mod.permute_region_id(np.arange(d.index, d.index + d.count, dtype=np.uint64), self.size)
        return j[j < W * H].astype(np.int64)

    def _unshuffle_host(self, strip, index, count, blend=1.0):
        """stitch_shuffled_strips.comp:21-63 / stitch_scanline.comp:21-50 on the host (gloo tests)."""
        W, H = self.size
        if self.strategy == abi.DISTRIBUTION_SCANLINE:
            rows = np.arange(index, H, count, dtype=np.int64)
            j = (rows[:, None] * W + np.arange(W, dtype=np.int64)[None, :]).reshape(-1).astype(np.uint64)
        else:
            j = dist_mod.permute_region_id(np.arange(index, index + count, dtype=np.uint64), self.size)
        ok = j < W * H
        jj = j[ok].astype(np.int64)
        full = self._full.numpy().reshape(self.views, W * H, 4)
        new = strip[:, ok]
        if blend < 1.0:
            # mix(old, new, a) = fma(new, a, old * (1 - a)) in fp32 (TR math); the product new*a is exact in fp64
            a = np.float32(blend)
            old = full[:, jj]
            t = (old * (np.float32(1.0) - a)).astype(np.float32)
            new = (new.astype(np.float64) * np.float64(a) + t.astype(np.float64)).astype(np.float32)
        full[:, jj] = new

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

    def read_back_async(self, host_ptr, byte_offset, nbytes, stream=None):
        """Frames in flight (the reference keeps MAX_FRAMES_IN_FLIGHT staging copies going, src/headless.cc:125-171): copy
        bytes [byte_offset, byte_offset + nbytes) of the stitched colour image of the frame rendered last into pinned host
        memory on a copy stream, beside the next frame. The copy starts after the frame's barrier; the next frame writes the
        other half of the double-buffered image, and this rank passes the NEXT barrier only when the copy has finished.
        `stream` = the stream the frame was rendered on (default: torch's current stream). Returns the torch event that
        marks the host bytes valid."""
        import torch
        assert self.cuda and self.world > 1 and self.stitch == "peer"
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        rs = torch.cuda.current_stream(self.device) if stream is None else \
            torch.cuda.ExternalStream(stream, device=torch.device("cuda", self.device))
        frame_done = torch.cuda.Event()
        frame_done.record(rs)
        self._copy_stream.wait_event(frame_done)
        self.tracer.read_full_image_range_async(host_ptr, byte_offset, nbytes, self._copy_stream.cuda_stream)
        self._copy_done = torch.cuda.Event()
        self._copy_done.record(self._copy_stream)
        return self._copy_done

    def full_image_ptr(self):
        """Device pointer of the stitched colour image of the last frame."""
        if self.stitch == "peer" and self.world > 1:
            info = self.tracer.full_image_info()
            return info.base + info.current_half * info.half_bytes
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
