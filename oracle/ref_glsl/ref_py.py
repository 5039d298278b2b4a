"""ctypes binding of oracle/_ref/libref_glsl_<variant>.so -- TEST INFRASTRUCTURE ONLY.

The libraries are the reference's own shader code (path_tracer.rgen and everything it includes, pre_transform.comp,
extract_tri_lights.comp) compiled for the CPU by build_ref.py from /root/reference. They exist in this container (and
travel to the GPU box as built files); `available()` says whether they are there. Only tests may use this module.
"""
import ctypes as C
import os

import numpy as np

from . import build_ref

_libs = {}


class RefTexture(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("format", C.c_uint32), ("reserved", C.c_uint32)]


class RefScene(C.Structure):
    _fields_ = [("instance_count", C.c_uint32), ("point_light_count", C.c_uint32), ("directional_light_count", C.c_uint32),
                ("tri_light_count", C.c_uint32), ("texture_count", C.c_uint32), ("view_count", C.c_uint32),
                ("instances", C.c_void_p), ("world_vertices", C.POINTER(C.c_void_p)), ("indices", C.POINTER(C.c_void_p)),
                ("opaque", C.c_void_p), ("point_lights", C.c_void_p), ("directional_lights", C.c_void_p), ("tri_lights", C.c_void_p),
                ("textures", C.POINTER(RefTexture)), ("environment_map", RefTexture), ("alias_table", C.c_void_p), ("cameras", C.c_void_p),
                ("environment_factor", C.c_float * 4), ("environment_proj", C.c_int32), ("reserved", C.c_uint32),
                ("trace", C.c_void_p), ("trace_user", C.c_void_p)]


class RefFrame(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("views", C.c_uint32), ("previous_samples", C.c_uint32),
                ("samples_accumulated", C.c_uint32), ("sample_counter", C.c_uint32), ("rng_seed", C.c_uint32),
                ("antialiasing", C.c_int32), ("min_ray_dist", C.c_float), ("indirect_clamping", C.c_float), ("film_radius", C.c_float),
                ("russian_roulette_delta", C.c_float), ("regularization_gamma", C.c_float), ("threads", C.c_int32)]


class RefCounters(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64)]


def available():
    try:
        return len(build_ref.build()) == len(build_ref.VARIANTS)
    except Exception as e:           # a failed build must not take test collection down with it
        import sys
        sys.stderr.write("oracle/_ref unavailable: %s\n" % e)
        return False


def lib(variant):
    if variant not in _libs:
        paths = build_ref.build()
        if variant not in paths or not os.path.exists(paths[variant]):
            raise ImportError("oracle/_ref/libref_glsl_%s.so is missing (it is built from /root/reference)" % variant)
        l = C.CDLL(paths[variant])
        l.ref_function.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32]
        l.ref_function_words.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        l.ref_pt_render.argtypes = [C.POINTER(RefScene), C.POINTER(RefFrame), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(RefCounters)]
        l.ref_pre_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        l.ref_extract_tri_lights.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        l.ref_variant_name.restype = C.c_char_p
        _libs[variant] = l
    return _libs[variant]


def function_words(variant, fn):
    wi, wo = C.c_int(), C.c_int()
    if lib(variant).ref_function_words(fn, C.byref(wi), C.byref(wo)) != 0:
        raise ValueError("unknown function id %d" % fn)
    return wi.value, wo.value


def function(variant, fn, records):
    """records: (n, words_in) uint32 (floats as their bit patterns) -> (n, words_out) uint32."""
    wi, wo = function_words(variant, fn)
    rec = np.ascontiguousarray(records, dtype=np.uint32).reshape(-1, wi)
    out = np.zeros((rec.shape[0], wo), dtype=np.uint32)
    if lib(variant).ref_function(fn, rec.ctypes.data, out.ctypes.data, rec.shape[0]) != 0:
        raise RuntimeError("ref_function(%d) failed" % fn)
    return out


class RefPathTracer:
    """The reference's path tracer pipeline on the CPU. `scene`: tauray_b200.scene.Scene; the ray caster for triangle
    candidates is the oracle's BVH (`oracle`: an oracle_py.Oracle with the same scene set)."""

    def __init__(self, variant, scene, cameras, oracle):
        from tauray_b200 import abi
        from .. import oracle_py
        self.l = lib(variant)
        self.abi = abi
        self.scene = scene
        self.oracle = oracle
        self.cameras = np.ascontiguousarray(cameras, dtype=abi.CAMERA_PAIR)
        self.inst = scene.instance_table()
        n = len(scene.meshes)
        # pre_transform.comp per instance
        self.world = []
        for i, (v, idx) in enumerate(scene.meshes):
            vin = np.ascontiguousarray(v, dtype=abi.VERTEX)
            vout = np.zeros_like(vin)
            self.l.ref_pre_transform(self.inst[i:i + 1].ctypes.data, vin.ctypes.data, vout.ctypes.data, vin.shape[0])
            self.world.append(vout)
        self.index = [np.ascontiguousarray(idx, dtype=np.uint32) for _, idx in scene.meshes]
        # extract_tri_lights.comp per emissive instance
        total = 0
        for i in range(n):
            if self.inst[i]["light_base_id"] >= 0:
                total = max(total, int(self.inst[i]["light_base_id"]) + self.index[i].size // 3)
        self.tri_lights = np.zeros(max(total, 1), dtype=abi.TRI_LIGHT)
        for i in range(n):
            if self.inst[i]["light_base_id"] >= 0:
                self.l.ref_extract_tri_lights(self.inst.ctypes.data, i, self.world[i].ctypes.data, self.index[i].ctypes.data,
                                              self.index[i].size // 3, self.tri_lights.ctypes.data)
        self.tri_light_count = total
        s = RefScene()
        s.instance_count = n
        s.point_light_count = scene.point_lights.shape[0]
        s.directional_light_count = scene.directional_lights.shape[0]
        s.tri_light_count = total
        s.texture_count = len(scene.textures)
        s.view_count = self.cameras.shape[0]
        s.instances = self.inst.ctypes.data
        self._wv = (C.c_void_p * max(n, 1))(*[w.ctypes.data for w in self.world])
        self._ix = (C.c_void_p * max(n, 1))(*[x.ctypes.data for x in self.index])
        s.world_vertices = self._wv
        s.indices = self._ix
        self._opaque = np.asarray(scene.opaque, dtype=np.uint32)
        s.opaque = self._opaque.ctypes.data
        s.point_lights = scene.point_lights.ctypes.data if scene.point_lights.size else None
        s.directional_lights = scene.directional_lights.ctypes.data if scene.directional_lights.size else None
        s.tri_lights = self.tri_lights.ctypes.data
        fmt = {np.dtype(np.uint8): abi.TEXTURE_RGBA8_UNORM, np.dtype(np.float16): abi.TEXTURE_RGBA16F, np.dtype(np.float32): abi.TEXTURE_RGBA32F}
        self._tex = (RefTexture * max(len(scene.textures), 1))()
        for i, t in enumerate(scene.textures):
            self._tex[i].data = t.ctypes.data
            self._tex[i].width, self._tex[i].height, self._tex[i].format = t.shape[1], t.shape[0], fmt[t.dtype]
        s.textures = self._tex
        s.environment_factor = (C.c_float * 4)(*scene.environment_factor)
        s.environment_proj = scene.environment_proj
        if scene.environment_proj >= 0:
            e = np.ascontiguousarray(scene.environment_map)
            self._env = e
            s.environment_map.data = e.ctypes.data
            s.environment_map.width, s.environment_map.height, s.environment_map.format = e.shape[1], e.shape[0], fmt[e.dtype]
            # the alias table is built on the host in the reference (src/environment_map.cc:39-140); taken from the oracle
            self._alias = oracle.alias_table()
            s.alias_table = self._alias.ctypes.data
        s.cameras = self.cameras.ctypes.data
        s.trace = C.cast(oracle_py.lib().orc_trace_one, C.c_void_p)
        s.trace_user = oracle._ctx
        self.s = s

    def render(self, size, options, frames=1, threads=0):
        """Accumulates `frames` frames of samples_per_pixel samples like rt_stage / rt_camera_stage drive the pipeline
        (sample_counter = frame * spp, previous_samples = pass * samples_per_pass, samples_accumulated). Returns the
        colour, diffuse and reflection targets (views, H, W, 4) float32 and the ray counters of the last frame."""
        abi = self.abi
        W, H = size
        V = self.cameras.shape[0]
        color = np.zeros((V, H, W, 4), dtype=np.float32)
        diffuse = np.zeros_like(color)
        reflection = np.zeros_like(color)
        f = RefFrame()
        f.width, f.height, f.views = W, H, V
        f.antialiasing = 1 if options.film != abi.FILM_POINT else 0
        f.min_ray_dist = options.min_ray_dist
        f.indirect_clamping = options.indirect_clamping
        f.film_radius = options.film_radius
        f.russian_roulette_delta = options.russian_roulette_delta
        f.regularization_gamma = options.regularization_gamma
        f.threads = threads
        seed = options.rng_seed & 0xFFFFFFFF
        if seed != 0:                                   # src/rt_stage.cc:76-85: rng_seed = pcg(seed) when non-zero
            x = (seed * 747796405 + 2891336453) & 0xFFFFFFFF
            x = ((((x >> ((x >> 28) + 4)) ^ x) * 277803737) & 0xFFFFFFFF)
            seed = ((x >> 22) ^ x) & 0xFFFFFFFF
        f.rng_seed = seed
        spp, per_pass = options.samples_per_pixel, options.samples_per_pass
        counters = RefCounters()
        closest = shadow = 0
        for frame in range(frames):
            f.sample_counter = frame * spp
            f.samples_accumulated = frame * spp
            closest = shadow = 0
            for p in range(spp // per_pass):
                f.previous_samples = p * per_pass
                rc = self.l.ref_pt_render(C.byref(self.s), C.byref(f), color.ctypes.data, diffuse.ctypes.data, reflection.ctypes.data, C.byref(counters))
                if rc != 0:
                    raise RuntimeError("ref_pt_render failed")
                closest += counters.closest_rays
                shadow += counters.shadow_rays
        return color, diffuse, reflection, {"closest_rays": closest, "shadow_rays": shadow}
