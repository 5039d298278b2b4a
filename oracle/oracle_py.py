"""ctypes loader for the CPU parity oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by bench.py's
cpu_baseline / --impl reference legs -- never by the tauray_b200 package (oracle/README.md).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from tauray_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """Compile liboracle.so with the committed Makefile (no-op when up to date)."""
    if force and os.path.exists(_LIB_PATH):
        os.remove(_LIB_PATH)
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        l = C.CDLL(_LIB_PATH)
        vp = C.c_void_p
        l.orc_create.argtypes = [C.POINTER(vp), C.POINTER(abi.Options), C.c_int]
        l.orc_destroy.argtypes = [vp]
        l.orc_destroy.restype = None
        l.orc_set_scene.argtypes = [vp, C.POINTER(abi.SceneDesc)]
        l.orc_set_cameras.argtypes = [vp, vp, C.c_uint32]
        l.orc_set_distribution.argtypes = [vp, C.POINTER(abi.Distribution)]
        l.orc_reset_accumulation.argtypes = [vp, C.c_int]
        l.orc_render.argtypes = [vp, C.c_uint64]
        l.orc_render_feature.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
        l.orc_read_color.argtypes = [vp, vp, C.c_size_t]
        l.orc_write_color.argtypes = [vp, vp, C.c_size_t]
        l.orc_get_stats.argtypes = [vp, C.POINTER(abi.Stats)]
        l.orc_read_tri_lights.argtypes = [vp, vp, C.c_uint32]
        l.orc_read_alias_table.argtypes = [vp, vp, C.c_uint32]
        l.orc_set_gbuffer_layers.argtypes = [vp, C.c_uint32]
        l.orc_read_gbuffer_layer.argtypes = [vp, C.c_int, vp, C.c_size_t]
        l.orc_trace_rays.argtypes = [vp, vp, C.c_uint32, vp]
        l.orc_sample_test.argtypes = [vp, vp, C.c_uint32, C.c_int, vp]
        l.orc_unshuffle.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_float]
        l.orc_math.argtypes = [C.c_int, vp, vp, vp, C.c_uint32]
        l.orc_function.argtypes = [vp, C.c_int, vp, vp, C.c_uint32]
        l.orc_trace_one.argtypes = [vp, vp, vp]
        l.orc_last_error.argtypes = [vp]
        l.orc_last_error.restype = C.c_char_p
        _lib = l
    return _lib


class OracleError(RuntimeError):
    pass


class Oracle:
    """Same call sequence as tauray_b200.PathTracer, on the host cores."""

    def __init__(self, options, threads=0):
        self._l = lib()
        self._ctx = C.c_void_p()
        self.options = abi.copy_options(options)
        rc = self._l.orc_create(C.byref(self._ctx), C.byref(self.options), threads)
        if rc:
            raise OracleError("orc_create: %d %s" % (rc, self._l.orc_last_error(None).decode()))
        self._dist = None
        self._views = 0

    def close(self):
        if self._ctx:
            self._l.orc_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            raise OracleError("%s: %d %s" % (what, rc, self._l.orc_last_error(self._ctx).decode()))
        return rc

    def set_scene(self, scene):
        d = scene.desc()
        self._check(self._l.orc_set_scene(self._ctx, C.byref(d)), "orc_set_scene")

    def set_cameras(self, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=abi.CAMERA_PAIR)
        self._views = pairs.shape[0]
        self._check(self._l.orc_set_cameras(self._ctx, pairs.ctypes.data, pairs.shape[0]), "orc_set_cameras")

    def set_distribution(self, dist):
        self._dist = dist
        self._check(self._l.orc_set_distribution(self._ctx, C.byref(dist)), "orc_set_distribution")

    def reset_accumulation(self, reset_sample_counter=True):
        self._check(self._l.orc_reset_accumulation(self._ctx, 1 if reset_sample_counter else 0), "orc_reset")

    def render(self, max_slots=0):
        self._check(self._l.orc_render(self._ctx, max_slots), "orc_render")

    def render_feature(self, feature, default=(0, 0, 0, 0)):
        dv = (C.c_float * 4)(*default)
        self._check(self._l.orc_render_feature(self._ctx, feature, dv), "orc_render_feature")

    def read_color(self):
        from tauray_b200.distribution import target_shape
        shape = target_shape(self._dist, self._views)
        out = np.empty(shape, dtype=np.float32)
        self._check(self._l.orc_read_color(self._ctx, out.ctypes.data, out.nbytes), "orc_read_color")
        return out

    def write_color(self, img):
        img = np.ascontiguousarray(img, dtype=np.float32)
        self._check(self._l.orc_write_color(self._ctx, img.ctypes.data, img.nbytes), "orc_write_color")

    def stats(self):
        s = abi.Stats()
        self._check(self._l.orc_get_stats(self._ctx, C.byref(s)), "orc_get_stats")
        return s.as_dict()

    def tri_lights(self):
        n = self._check(self._l.orc_read_tri_lights(self._ctx, None, 0), "orc_read_tri_lights")
        out = np.zeros(max(n, 1), dtype=abi.TRI_LIGHT)
        self._l.orc_read_tri_lights(self._ctx, out.ctypes.data, n)
        return out[:n]

    def set_gbuffer_layers(self, layers):
        mask = 0
        for l in layers:
            mask |= 1 << int(l)
        self._check(self._l.orc_set_gbuffer_layers(self._ctx, mask), "orc_set_gbuffer_layers")

    def read_gbuffer_layer(self, layer):
        dt, ch = abi.GBUFFER_TEXEL[int(layer)]
        texels = self.read_color().size // 4
        out = np.zeros(texels * ch, dtype=dt)
        self._check(self._l.orc_read_gbuffer_layer(self._ctx, int(layer), out.ctypes.data, out.nbytes), "orc_read_gbuffer_layer")
        return out.reshape(self._views, -1, ch)

    def alias_table(self):
        n = self._check(self._l.orc_read_alias_table(self._ctx, None, 0), "orc_read_alias_table")
        out = np.zeros(max(n, 1), dtype=abi.ALIAS_TABLE_ENTRY)
        self._l.orc_read_alias_table(self._ctx, out.ctypes.data, n)
        return out[:n]

    def trace_rays(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        out = np.zeros(rays.shape[0], dtype=abi.HIT)
        self._check(self._l.orc_trace_rays(self._ctx, rays.ctypes.data, rays.shape[0], out.ctypes.data), "orc_trace_rays")
        return out

    def sample_test(self, coords, draw_index):
        coords = np.ascontiguousarray(coords, dtype=np.uint32).reshape(-1, 4)
        out = np.zeros_like(coords)
        self._check(self._l.orc_sample_test(self._ctx, coords.ctypes.data, coords.shape[0], draw_index, out.ctypes.data), "orc_sample_test")
        return out

    def function(self, fn, records, words_out):
        """The restatement's version of function `fn` of oracle/ref_glsl/ref_unit_fn.cc's table over packed records."""
        rec = np.ascontiguousarray(records, dtype=np.uint32)
        out = np.zeros((rec.shape[0], words_out), dtype=np.uint32)
        self._check(self._l.orc_function(self._ctx, int(fn), rec.ctypes.data, out.ctypes.data, rec.shape[0]), "orc_function")
        return out

    def unshuffle(self, strip, index, count, full_image, blend_ratio=1.0):
        strip = np.ascontiguousarray(strip, dtype=np.float32)
        self._check(self._l.orc_unshuffle(self._ctx, strip.ctypes.data, index, count, full_image.ctypes.data, blend_ratio), "orc_unshuffle")


def math_op(op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(a if b is None else b, dtype=np.float32)
    out = np.empty_like(a)
    rc = lib().orc_math(op, a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size)
    if rc:
        raise OracleError("orc_math %d" % rc)
    return out
