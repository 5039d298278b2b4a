"""ctypes binding of libtauray_b200.so (include/tauray_b200.h).

`PathTracer` is the Python spelling of one `tr_pt_ctx`, i.e. of one `path_tracer_stage` of the
reference (src/path_tracer_stage.hh:10-52): construct with options, `set_scene`, `set_cameras`,
`set_distribution`, then `render()` once per frame. The library is loaded from the package
directory (built in-tree by `__graft_entry__.build()` / `make -C tauray_b200/csrc`); a missing
library or a missing GPU is an error -- there is no fallback path.
"""
import ctypes as C
import os

import numpy as np

from . import abi
from .distribution import target_shape

_HERE = os.path.dirname(os.path.abspath(__file__))
# TAURAY_B200_LIB: development override used to A/B differently tuned builds of the same library
LIB_PATH = os.environ.get("TAURAY_B200_LIB") or os.path.join(_HERE, "libtauray_b200.so")
_lib = None

EXPORTS = [
    "tr_pt_default_options", "tr_pt_create", "tr_pt_destroy", "tr_pt_set_scene", "tr_pt_update_instances", "tr_pt_set_cameras",
    "tr_pt_set_distribution", "tr_pt_reset_accumulation", "tr_pt_render", "tr_pt_render_feature", "tr_pt_sync",
    "tr_pt_get_color_buffer", "tr_pt_read_color", "tr_pt_get_accumulation_state", "tr_pt_restore_accumulation", "tr_pt_unshuffle", "tr_pt_set_peer_images", "tr_pt_set_stitch_blend", "tr_pt_get_stats",
    "tr_pt_set_gbuffer_layers", "tr_pt_get_gbuffer_layer", "tr_pt_read_gbuffer_layer", "tr_pt_write_gbuffer_layer",
    "tr_pt_set_instrumentation", "tr_pt_read_tri_lights", "tr_pt_read_alias_table", "tr_pt_trace_rays", "tr_pt_sample_test",
    "tr_pt_last_error", "tr_pt_abi_version", "tr_pt_alloc_full_image", "tr_pt_ipc_get_handle",
    "tr_pt_ipc_open_handle", "tr_pt_ipc_close_handle", "tr_pt_set_color_buffer", "tr_pt_read_full_image", "tr_pt_enable_peer_access",
    "tr_pt_get_full_image_info", "tr_pt_set_stitch_double_buffer", "tr_pt_set_render_in_full_image", "tr_pt_read_full_image_range",
    "tr_pt_read_full_layer", "tr_pt_read_color_async", "tr_pt_read_full_image_range_async", "tr_pt_set_update_mode", "tr_pt_set_cameras_async", "tr_pt_render_begin", "tr_pt_render_end",
]


class TaurayB200Error(RuntimeError):
    def __init__(self, what, code, message):
        super().__init__("%s failed with %d: %s" % (what, code, message))
        self.code = code


def lib():
    """Load libtauray_b200.so and declare every prototype of include/tauray_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(tauray_b200 has no CPU fallback)" % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    l.tr_pt_default_options.argtypes = [C.POINTER(abi.Options)]
    l.tr_pt_default_options.restype = None
    l.tr_pt_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(abi.Options)]
    l.tr_pt_destroy.argtypes = [vp]
    l.tr_pt_destroy.restype = None
    l.tr_pt_set_scene.argtypes = [vp, C.POINTER(abi.SceneDesc)]
    l.tr_pt_update_instances.argtypes = [vp, vp, C.c_uint32]
    l.tr_pt_set_cameras.argtypes = [vp, vp, C.c_uint32]
    l.tr_pt_set_distribution.argtypes = [vp, C.POINTER(abi.Distribution)]
    l.tr_pt_reset_accumulation.argtypes = [vp, C.c_int]
    l.tr_pt_render.argtypes = [vp, vp]
    l.tr_pt_render_begin.argtypes = [vp, vp]
    l.tr_pt_render_end.argtypes = [vp, vp]
    l.tr_pt_render_feature.argtypes = [vp, C.c_int, C.POINTER(C.c_float), vp]
    l.tr_pt_sync.argtypes = [vp]
    l.tr_pt_get_color_buffer.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    l.tr_pt_get_accumulation_state.argtypes = [vp, C.POINTER(abi.AccumulationState)]
    l.tr_pt_restore_accumulation.argtypes = [vp, vp, C.c_size_t, C.POINTER(abi.AccumulationState)]
    l.tr_pt_read_color.argtypes = [vp, vp, C.c_size_t]
    l.tr_pt_set_update_mode.argtypes = [vp, C.c_int]
    l.tr_pt_set_cameras_async.argtypes = [vp, vp, C.c_uint32, vp]
    l.tr_pt_read_color_async.argtypes = [vp, vp, C.c_size_t, vp]
    l.tr_pt_unshuffle.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_float, vp]
    l.tr_pt_set_stitch_blend.argtypes = [vp, C.c_float]
    l.tr_pt_set_peer_images.argtypes = [vp, C.POINTER(vp), C.c_int]
    l.tr_pt_get_stats.argtypes = [vp, C.POINTER(abi.Stats)]
    l.tr_pt_set_instrumentation.argtypes = [vp, C.c_int]
    l.tr_pt_read_tri_lights.argtypes = [vp, vp, C.c_uint32]
    l.tr_pt_set_gbuffer_layers.argtypes = [vp, C.c_uint32]
    l.tr_pt_get_gbuffer_layer.argtypes = [vp, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    l.tr_pt_read_gbuffer_layer.argtypes = [vp, C.c_int, vp, C.c_size_t]
    l.tr_pt_write_gbuffer_layer.argtypes = [vp, C.c_int, vp, C.c_size_t]
    l.tr_pt_read_alias_table.argtypes = [vp, vp, C.c_uint32]
    l.tr_pt_trace_rays.argtypes = [vp, vp, C.c_uint32, vp]
    l.tr_pt_sample_test.argtypes = [vp, vp, C.c_uint32, C.c_int, vp]
    l.tr_pt_alloc_full_image.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    l.tr_pt_read_full_image.argtypes = [vp, vp, C.c_size_t]
    l.tr_pt_enable_peer_access.argtypes = [vp, C.c_int]
    l.tr_pt_get_full_image_info.argtypes = [vp, C.POINTER(abi.FullImageInfo)]
    l.tr_pt_set_stitch_double_buffer.argtypes = [vp, C.c_int]
    l.tr_pt_set_render_in_full_image.argtypes = [vp, C.c_int]
    l.tr_pt_read_full_image_range.argtypes = [vp, vp, C.c_size_t, C.c_size_t]
    l.tr_pt_read_full_image_range_async.argtypes = [vp, vp, C.c_size_t, C.c_size_t, vp]
    l.tr_pt_read_full_layer.argtypes = [vp, C.c_int, vp, C.c_size_t]
    l.tr_pt_ipc_get_handle.argtypes = [vp, C.c_char_p]
    l.tr_pt_ipc_open_handle.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    l.tr_pt_ipc_close_handle.argtypes = [vp, vp]
    l.tr_pt_set_color_buffer.argtypes = [vp, vp, C.c_size_t]
    l.tr_pt_last_error.argtypes = [vp]
    l.tr_pt_last_error.restype = C.c_char_p
    l.tr_pt_abi_version.argtypes = []
    if l.tr_pt_abi_version() != abi.ABI_VERSION:
        raise ImportError("libtauray_b200.so has ABI %d, expected %d" % (l.tr_pt_abi_version(), abi.ABI_VERSION))
    _lib = l
    return l


class PathTracer:
    """One path_tracer_stage on one GPU."""

    def __init__(self, options, device=0):
        self._l = lib()
        self._ctx = C.c_void_p()
        self.options = abi.copy_options(options)
        self.device = device
        rc = self._l.tr_pt_create(C.byref(self._ctx), device, C.byref(self.options))
        if rc:
            raise TaurayB200Error("tr_pt_create", rc, self._l.tr_pt_last_error(None).decode())
        self._dist = None
        self._views = 0

    def close(self):
        if getattr(self, "_ctx", None):
            self._l.tr_pt_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc, what):
        if rc < 0:
            raise TaurayB200Error(what, rc, self._l.tr_pt_last_error(self._ctx).decode())
        return rc

    # -- reference-shaped calls -----------------------------------------------------------
    def set_scene(self, scene):
        d = scene.desc() if hasattr(scene, "desc") else scene
        self._check(self._l.tr_pt_set_scene(self._ctx, C.byref(d)), "tr_pt_set_scene")

    def update_instances(self, instances):
        """Per-frame scene update of an animated scene: `instances` = the new instance records
        (abi.INSTANCE array, e.g. Scene.instance_table())."""
        inst = np.ascontiguousarray(instances, dtype=abi.INSTANCE)
        self._check(self._l.tr_pt_update_instances(self._ctx, inst.ctypes.data, inst.shape[0]), "tr_pt_update_instances")

    def set_cameras_async(self, pinned_pairs, stream):
        """The camera table in stream order (frames in flight): `pinned_pairs` = abi.CAMERA_PAIR array in pinned host
        memory that stays untouched until the copy has run on `stream`, the stream the next renders use."""
        assert pinned_pairs.dtype == abi.CAMERA_PAIR and pinned_pairs.flags["C_CONTIGUOUS"]
        self._views = pinned_pairs.shape[0]
        self._check(self._l.tr_pt_set_cameras_async(self._ctx, pinned_pairs.ctypes.data, pinned_pairs.shape[0], stream), "tr_pt_set_cameras_async")

    def set_update_mode(self, mode):
        """abi.UPDATE_REBUILD (a new radix tree per update) or abi.UPDATE_REFIT (the scene's tree keeps its topology)."""
        self._check(self._l.tr_pt_set_update_mode(self._ctx, mode), "tr_pt_set_update_mode")

    def set_cameras(self, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=abi.CAMERA_PAIR)
        self._views = pairs.shape[0]
        self._check(self._l.tr_pt_set_cameras(self._ctx, pairs.ctypes.data, pairs.shape[0]), "tr_pt_set_cameras")

    def set_distribution(self, dist):
        """rt_camera_stage::reset_distribution_params."""
        self._dist = dist
        self._check(self._l.tr_pt_set_distribution(self._ctx, C.byref(dist)), "tr_pt_set_distribution")

    reset_distribution_params = set_distribution

    def reset_accumulation(self, reset_sample_counter=True):
        """reset_accumulated_samples (+ reset_sample_counter), src/rt_renderer.cc:67-81."""
        self._check(self._l.tr_pt_reset_accumulation(self._ctx, 1 if reset_sample_counter else 0),
                    "tr_pt_reset_accumulation")

    def render(self, stream=None):
        """stage::run: all passes of one frame, asynchronous."""
        self._check(self._l.tr_pt_render(self._ctx, stream), "tr_pt_render")

    def render_begin(self, stream=None):
        """Frames in flight: everything of the next frame but the accumulation, on a stream of the context's own."""
        self._check(self._l.tr_pt_render_begin(self._ctx, stream), "tr_pt_render_begin")

    def render_end(self, stream=None):
        """The accumulation (and stitch stores) of the oldest begun frame, on `stream`."""
        self._check(self._l.tr_pt_render_end(self._ctx, stream), "tr_pt_render_end")

    def render_feature(self, feature, default=(0.0, 0.0, 0.0, 0.0), stream=None):
        dv = (C.c_float * 4)(*default)
        self._check(self._l.tr_pt_render_feature(self._ctx, feature, dv, stream), "tr_pt_render_feature")

    def sync(self):
        self._check(self._l.tr_pt_sync(self._ctx), "tr_pt_sync")

    # -- results --------------------------------------------------------------------------
    def target_shape(self):
        return target_shape(self._dist, self._views)

    def read_color(self, out=None):
        shape = self.target_shape()
        if out is None:
            out = np.empty(shape, dtype=np.float32)
        assert out.nbytes == int(np.prod(shape)) * 4 and out.flags["C_CONTIGUOUS"]
        self._check(self._l.tr_pt_read_color(self._ctx, out.ctypes.data, out.nbytes), "tr_pt_read_color")
        return out

    def read_color_into(self, host_ptr, nbytes):
        self._check(self._l.tr_pt_read_color(self._ctx, host_ptr, nbytes), "tr_pt_read_color")

    def read_color_async(self, host_ptr, nbytes, copy_stream):
        """D2H of the colour target on `copy_stream` without blocking the host (frames in flight)."""
        self._check(self._l.tr_pt_read_color_async(self._ctx, host_ptr, nbytes, copy_stream), "tr_pt_read_color_async")

    def checkpoint(self):
        """Accumulation checkpoint: the colour target, the enabled aux G-buffer layers and the counters."""
        st = abi.AccumulationState()
        self._check(self._l.tr_pt_get_accumulation_state(self._ctx, C.byref(st)), "tr_pt_get_accumulation_state")
        ck = {"color": self.read_color().copy()}
        for name, _ in st._fields_:
            ck[name] = int(getattr(st, name))
        ck["layers"] = {l: self.read_gbuffer_layer(l).copy() for l in range(1, abi.GBUFFER_LAYER_COUNT) if (st.gbuffer_mask >> l) & 1}
        return ck

    def restore(self, ck):
        """Resume from PathTracer.checkpoint() (scene, cameras, distribution and layers must already be set)."""
        color = np.ascontiguousarray(ck["color"], dtype=np.float32)
        st = abi.AccumulationState(int(ck["frame_counter"]), int(ck["accumulated_samples"]),
                                   int(ck.get("samples_per_pixel", self.options.samples_per_pixel)),
                                   int(ck.get("gbuffer_mask", 0)), int(ck.get("color_pixels", color.size // 4)))
        self._check(self._l.tr_pt_restore_accumulation(self._ctx, color.ctypes.data, color.nbytes, C.byref(st)),
                    "tr_pt_restore_accumulation")
        for l, texels in ck.get("layers", {}).items():
            self.write_gbuffer_layer(l, texels)

    def write_gbuffer_layer(self, layer, texels):
        texels = np.ascontiguousarray(texels)
        self._check(self._l.tr_pt_write_gbuffer_layer(self._ctx, int(layer), texels.ctypes.data, texels.nbytes), "tr_pt_write_gbuffer_layer")

    def color_buffer(self):
        """(device pointer, bytes) of the colour target."""
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._l.tr_pt_get_color_buffer(self._ctx, C.byref(p), C.byref(n)), "tr_pt_get_color_buffer")
        return p.value, n.value

    def unshuffle(self, strip_ptr, index, count, full_image_ptr, stream=None, blend_ratio=1.0):
        self._check(self._l.tr_pt_unshuffle(self._ctx, strip_ptr, index, count, full_image_ptr, blend_ratio, stream), "tr_pt_unshuffle")

    def set_stitch_blend(self, blend_ratio):
        self._check(self._l.tr_pt_set_stitch_blend(self._ctx, float(blend_ratio)), "tr_pt_set_stitch_blend")

    def set_peer_images(self, ptrs):
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        self._check(self._l.tr_pt_set_peer_images(self._ctx, arr, len(ptrs)), "tr_pt_set_peer_images")

    def set_color_buffer(self, device_ptr, nbytes):
        """Render into caller-owned device memory (None restores the internal target)."""
        self._check(self._l.tr_pt_set_color_buffer(self._ctx, device_ptr, nbytes), "tr_pt_set_color_buffer")

    def alloc_full_image(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self._l.tr_pt_alloc_full_image(self._ctx, C.byref(p), C.byref(n)), "tr_pt_alloc_full_image")
        return p.value, n.value

    def read_full_image(self, out):
        self._check(self._l.tr_pt_read_full_image(self._ctx, out.ctypes.data, out.nbytes), "tr_pt_read_full_image")
        return out

    def full_image_info(self):
        info = abi.FullImageInfo()
        self._check(self._l.tr_pt_get_full_image_info(self._ctx, C.byref(info)), "tr_pt_get_full_image_info")
        return info

    def set_stitch_double_buffer(self, enable=True):
        self._check(self._l.tr_pt_set_stitch_double_buffer(self._ctx, 1 if enable else 0), "tr_pt_set_stitch_double_buffer")

    def set_render_in_full_image(self, enable=True):
        self._check(self._l.tr_pt_set_render_in_full_image(self._ctx, 1 if enable else 0), "tr_pt_set_render_in_full_image")

    def read_full_image_range(self, host_ptr, byte_offset, nbytes):
        """D2H of bytes [byte_offset, byte_offset + nbytes) of the stitched colour image to host_ptr."""
        self._check(self._l.tr_pt_read_full_image_range(self._ctx, host_ptr, byte_offset, nbytes), "tr_pt_read_full_image_range")

    def read_full_image_range_async(self, host_ptr, byte_offset, nbytes, copy_stream):
        """The same on `copy_stream` without blocking the host; the caller orders the stream after the frame's barrier."""
        self._check(self._l.tr_pt_read_full_image_range_async(self._ctx, host_ptr, byte_offset, nbytes, copy_stream),
                    "tr_pt_read_full_image_range_async")

    def read_full_layer(self, layer, size, views):
        """Stitched aux layer as (views, H, W, channels) in the layer's own texel format."""
        dt, ch = abi.GBUFFER_TEXEL[int(layer)]
        out = np.zeros((views, size[1], size[0], ch), dtype=dt)
        self._check(self._l.tr_pt_read_full_layer(self._ctx, int(layer), out.ctypes.data, out.nbytes), "tr_pt_read_full_layer")
        return out

    def ipc_get_handle(self):
        buf = C.create_string_buffer(64)
        self._check(self._l.tr_pt_ipc_get_handle(self._ctx, buf), "tr_pt_ipc_get_handle")
        return buf.raw

    def ipc_open_handle(self, handle):
        p = C.c_void_p()
        self._check(self._l.tr_pt_ipc_open_handle(self._ctx, handle, C.byref(p)), "tr_pt_ipc_open_handle")
        return p.value

    def ipc_close_handle(self, ptr):
        self._check(self._l.tr_pt_ipc_close_handle(self._ctx, ptr), "tr_pt_ipc_close_handle")

    def stats(self):
        s = abi.Stats()
        self._check(self._l.tr_pt_get_stats(self._ctx, C.byref(s)), "tr_pt_get_stats")
        return s.as_dict()

    def set_gbuffer_layers(self, layers):
        """Enable aux G-buffer layers (iterable of abi.GBUFFER_* ids)."""
        mask = 0
        for l in layers:
            mask |= 1 << int(l)
        self._check(self._l.tr_pt_set_gbuffer_layers(self._ctx, mask), "tr_pt_set_gbuffer_layers")

    def gbuffer_layer_ptr(self, layer):
        ptr, n = C.c_void_p(), C.c_size_t()
        self._check(self._l.tr_pt_get_gbuffer_layer(self._ctx, int(layer), C.byref(ptr), C.byref(n)), "tr_pt_get_gbuffer_layer")
        return ptr.value, n.value

    def read_gbuffer_layer(self, layer):
        """Layer texels as (views, layer pixels, channels) in the layer's own format (abi.GBUFFER_TEXEL)."""
        _, nbytes = self.gbuffer_layer_ptr(layer)
        dt, ch = abi.GBUFFER_TEXEL[int(layer)]
        out = np.zeros(nbytes // np.dtype(dt).itemsize, dtype=dt)
        self._check(self._l.tr_pt_read_gbuffer_layer(self._ctx, int(layer), out.ctypes.data, nbytes), "tr_pt_read_gbuffer_layer")
        return out.reshape(self._views, -1, ch)

    def set_instrumentation(self, flags):
        self._check(self._l.tr_pt_set_instrumentation(self._ctx, int(flags)), "tr_pt_set_instrumentation")

    # -- test hooks -----------------------------------------------------------------------
    def tri_lights(self):
        n = self._check(self._l.tr_pt_read_tri_lights(self._ctx, None, 0), "tr_pt_read_tri_lights")
        out = np.zeros(max(n, 1), dtype=abi.TRI_LIGHT)
        self._check(self._l.tr_pt_read_tri_lights(self._ctx, out.ctypes.data, n), "tr_pt_read_tri_lights")
        return out[:n]

    def alias_table(self):
        n = self._check(self._l.tr_pt_read_alias_table(self._ctx, None, 0), "tr_pt_read_alias_table")
        out = np.zeros(max(n, 1), dtype=abi.ALIAS_TABLE_ENTRY)
        self._check(self._l.tr_pt_read_alias_table(self._ctx, out.ctypes.data, n), "tr_pt_read_alias_table")
        return out[:n]

    def trace_rays(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = np.zeros(rays.shape[0], dtype=abi.HIT)
        self._check(self._l.tr_pt_trace_rays(self._ctx, rays.ctypes.data, rays.shape[0], out.ctypes.data), "tr_pt_trace_rays")
        return out

    def sample_test(self, coords, draw_index):
        coords = np.ascontiguousarray(coords, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros_like(coords)
        self._check(self._l.tr_pt_sample_test(self._ctx, coords.ctypes.data, coords.shape[0], draw_index, out.ctypes.data), "tr_pt_sample_test")
        return out
