"""ctypes / numpy mirrors of include/tauray_b200.h.

The numpy dtypes are the reference's GPU wire formats (GLSL ``scalar`` layout) cited in the
header: ``mesh::vertex`` (src/mesh.hh:19-25), ``instance_buffer`` / ``material_buffer``
(src/scene_stage.cc:17-44), the light entries (src/scene_stage.cc:46-111) and
``matrix_camera_data_buffer`` (src/camera.cc:397-407). Matrices are column-major, so a numpy
``(4, 4)`` field holds the *transpose* of the usual row-major math matrix; use
:func:`mat_to_wire`.
"""
import ctypes as C

import numpy as np

ABI_VERSION = 2

# enums --------------------------------------------------------------------------------------
SAMPLER_UNIFORM_RANDOM, SAMPLER_SOBOL_OWEN, SAMPLER_SOBOL_Z_ORDER_2D, SAMPLER_SOBOL_Z_ORDER_3D = range(4)
FILM_POINT, FILM_BOX, FILM_BLACKMAN_HARRIS = range(3)
MIS_DISABLED, MIS_BALANCE_HEURISTIC, MIS_POWER_HEURISTIC = range(3)
BOUNCE_HEMISPHERE, BOUNCE_COSINE_HEMISPHERE, BOUNCE_MATERIAL = range(3)
TRI_LIGHT_AREA, TRI_LIGHT_SOLID_ANGLE, TRI_LIGHT_HYBRID = range(3)
DISTRIBUTION_DUPLICATE, DISTRIBUTION_SCANLINE, DISTRIBUTION_SHUFFLED_STRIPS = range(3)
PROJECTION_PERSPECTIVE, PROJECTION_ORTHOGRAPHIC = range(2)
FEATURE_ALBEDO, FEATURE_WORLD_NORMAL, FEATURE_VIEW_NORMAL, FEATURE_WORLD_POS, FEATURE_VIEW_POS, FEATURE_DISTANCE = range(6)
FEATURE_INSTANCE_ID = 9
KERNEL_WAVEFRONT, KERNEL_MEGAKERNEL = range(2)
ESTIMATOR_PATH_TRACER, ESTIMATOR_DIRECT = range(2)   # direct = direct_stage, shader/direct.rgen
INSTRUMENT_VISITS, INSTRUMENT_TIMING = 1, 2
UPDATE_REBUILD, UPDATE_REFIT = 0, 1          # tr_pt_set_update_mode
TEXTURE_RGBA8_UNORM, TEXTURE_RGBA32F, TEXTURE_RGBA16F = range(3)
ENVMAP_NONE, ENVMAP_LAT_LONG = -1, 0
(GBUFFER_COLOR, GBUFFER_DIFFUSE, GBUFFER_REFLECTION, GBUFFER_ALBEDO, GBUFFER_MATERIAL, GBUFFER_NORMAL, GBUFFER_POS,
 GBUFFER_SCREEN_MOTION, GBUFFER_INSTANCE_ID) = range(9)
GBUFFER_LAYER_COUNT = 9
# numpy view of one texel of each layer (src/gbuffer.hh:18-51)
GBUFFER_TEXEL = {
    GBUFFER_COLOR: ("<f4", 4), GBUFFER_DIFFUSE: ("<f2", 4), GBUFFER_REFLECTION: ("<f2", 4), GBUFFER_ALBEDO: ("<f2", 4),
    GBUFFER_MATERIAL: ("u1", 4), GBUFFER_NORMAL: ("<i2", 2), GBUFFER_POS: ("<f4", 4), GBUFFER_SCREEN_MOTION: ("<f4", 4),
    GBUFFER_INSTANCE_ID: ("<i4", 1),
}

OK = 0
ERR_INVALID_ARGUMENT, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_SCENE, ERR_OUT_OF_MEMORY = -1, -2, -3, -4, -5, -6

# numpy table dtypes ---------------------------------------------------------------------------
VERTEX = np.dtype([("pos", "<f4", 3), ("normal", "<f4", 3), ("uv", "<f4", 2), ("tangent", "<f4", 4)])
MATERIAL = np.dtype([
    ("albedo_factor", "<f4", 4), ("metallic_roughness_factor", "<f4", 4), ("emission_factor", "<f4", 4),
    ("transmittance", "<f4"), ("ior", "<f4"), ("normal_factor", "<f4"), ("flags", "<u4"),
    ("albedo_tex_id", "<i4"), ("metallic_roughness_tex_id", "<i4"), ("normal_tex_id", "<i4"),
    ("emission_tex_id", "<i4"),
])
INSTANCE = np.dtype([
    ("light_base_id", "<i4"), ("sh_grid_index", "<i4"), ("pad", "<u4"), ("shadow_terminator_mul", "<f4"),
    ("model", "<f4", (4, 4)), ("model_normal", "<f4", (4, 4)), ("model_prev", "<f4", (4, 4)),
    ("mat", MATERIAL),
])
DIRECTIONAL_LIGHT = np.dtype([("color", "<f4", 3), ("shadow_map_index", "<i4"), ("dir", "<f4", 3), ("dir_cutoff", "<f4")])
POINT_LIGHT = np.dtype([
    ("color", "<f4", 3), ("dir", "<f4", 3), ("pos", "<f4", 3), ("radius", "<f4"), ("dir_cutoff", "<f4"),
    ("dir_falloff", "<f4"), ("cutoff_radius", "<f4"), ("spot_radius", "<f4"), ("shadow_map_index", "<i4"),
    ("padding", "<i4"),
])
TRI_LIGHT = np.dtype([
    ("pos", "<f4", (3, 3)), ("emission_factor", "<u4"), ("instance_id", "<u4"), ("primitive_id", "<u4"),
    ("uv", "<u4", 3), ("emission_tex_id", "<i4"),
])
CAMERA_DATA = np.dtype([
    ("view", "<f4", (4, 4)), ("view_inverse", "<f4", (4, 4)), ("view_proj", "<f4", (4, 4)),
    ("proj_inverse", "<f4", (4, 4)), ("origin", "<f4", 4), ("dof_params", "<f4", 4),
    ("projection_info", "<f4", 4), ("pan", "<f4", 4),
])
CAMERA_PAIR = np.dtype([("current", CAMERA_DATA), ("previous", CAMERA_DATA)])
HIT = np.dtype([("instance_id", "<i4"), ("primitive_id", "<i4"), ("u", "<f4"), ("v", "<f4"), ("t", "<f4")])
ALIAS_TABLE_ENTRY = np.dtype([("alias_id", "<u4"), ("probability", "<u4"), ("pdf", "<f4"), ("alias_pdf", "<f4")])

assert VERTEX.itemsize == 48 and MATERIAL.itemsize == 80 and INSTANCE.itemsize == 288
assert DIRECTIONAL_LIGHT.itemsize == 32 and POINT_LIGHT.itemsize == 64 and TRI_LIGHT.itemsize == 64
assert CAMERA_DATA.itemsize == 320 and CAMERA_PAIR.itemsize == 640 and HIT.itemsize == 20


def mat_to_wire(m):
    """Row-major math matrix (4, 4) -> column-major wire layout."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float32).T)


# ctypes structs -----------------------------------------------------------------------------
class MeshRef(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("indices", C.c_void_p), ("vertex_count", C.c_uint32),
                ("index_count", C.c_uint32), ("opaque", C.c_uint32), ("reserved", C.c_uint32)]


class Texture(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32),
                ("format", C.c_uint32), ("reserved", C.c_uint32)]


class SceneDesc(C.Structure):
    _fields_ = [("instance_count", C.c_uint32), ("point_light_count", C.c_uint32),
                ("directional_light_count", C.c_uint32), ("texture_count", C.c_uint32),
                ("instances", C.c_void_p), ("meshes", C.POINTER(MeshRef)), ("point_lights", C.c_void_p),
                ("directional_lights", C.c_void_p), ("textures", C.POINTER(Texture)),
                ("environment_factor", C.c_float * 4), ("environment_proj", C.c_int32),
                ("reserved", C.c_uint32), ("environment_map", C.POINTER(Texture))]


class Options(C.Structure):
    """tr_pt_options; see include/tauray_b200.h for the reference citations of every field."""
    _fields_ = [
        ("max_ray_depth", C.c_int32), ("min_ray_dist", C.c_float), ("rng_seed", C.c_int32),
        ("local_sampler", C.c_int32),
        ("samples_per_pixel", C.c_int32), ("samples_per_pass", C.c_int32), ("projection", C.c_int32),
        ("transparent_background", C.c_int32), ("active_viewport_count", C.c_uint32),
        ("use_shadow_terminator_fix", C.c_int32), ("use_white_albedo_on_first_bounce", C.c_int32),
        ("hide_lights", C.c_int32), ("film", C.c_int32), ("mis_mode", C.c_int32),
        ("film_radius", C.c_float), ("russian_roulette_delta", C.c_float), ("indirect_clamping", C.c_float),
        ("regularization_gamma", C.c_float), ("depth_of_field", C.c_int32),
        ("sampling_weight_point_lights", C.c_float), ("sampling_weight_directional_lights", C.c_float),
        ("sampling_weight_envmap", C.c_float), ("sampling_weight_emissive_triangles", C.c_float),
        ("bounce_mode", C.c_int32), ("tri_light_mode", C.c_int32),
        ("kernel_mode", C.c_int32), ("estimator", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class Distribution(C.Structure):
    _fields_ = [("size", C.c_uint32 * 2), ("strategy", C.c_int32), ("index", C.c_uint32),
                ("count", C.c_uint32), ("primary", C.c_int32)]


class AccumulationState(C.Structure):
    """tr_pt_accumulation_state"""
    _fields_ = [("frame_counter", C.c_uint32), ("accumulated_samples", C.c_uint32), ("samples_per_pixel", C.c_uint32),
                ("gbuffer_mask", C.c_uint32), ("color_pixels", C.c_uint64)]


class FullImageInfo(C.Structure):
    """tr_pt_full_image_info"""
    _fields_ = [("base", C.c_void_p), ("arena_bytes", C.c_uint64), ("half_bytes", C.c_uint64),
                ("layer_offset", C.c_uint64 * 9), ("pixels", C.c_uint64), ("generation", C.c_uint32),
                ("current_half", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("paths", C.c_uint64),
                ("closest_node_visits", C.c_uint64), ("closest_triangle_tests", C.c_uint64),
                ("shadow_node_visits", C.c_uint64), ("shadow_triangle_tests", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("culled_shadow_rays", C.c_uint64),
                ("render_ms", C.c_float), ("closest_ms", C.c_float), ("shadow_ms", C.c_float),
                ("shade_ms", C.c_float), ("other_ms", C.c_float),
                ("closest_launches", C.c_uint32), ("shadow_launches", C.c_uint32), ("shade_launches", C.c_uint32),
                ("bvh_build_ms", C.c_float), ("bvh_node_count", C.c_uint32), ("bvh_depth", C.c_uint32),
                ("triangle_count", C.c_uint32), ("tri_light_count", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


def default_options(**overrides):
    """CLI defaults of the reference (src/options.hh; SURVEY.md section 5 'Config / flags')."""
    o = Options()
    o.max_ray_depth = 8
    o.min_ray_dist = 1e-4
    o.rng_seed = 0
    o.local_sampler = SAMPLER_UNIFORM_RANDOM
    o.samples_per_pixel = 1
    o.samples_per_pass = 1
    o.projection = PROJECTION_PERSPECTIVE
    o.transparent_background = 0
    o.active_viewport_count = 1
    o.film = FILM_POINT
    o.mis_mode = MIS_POWER_HEURISTIC
    o.film_radius = 0.5
    o.sampling_weight_point_lights = 1.0
    o.sampling_weight_directional_lights = 1.0
    o.sampling_weight_envmap = 1.0
    o.sampling_weight_emissive_triangles = 1.0
    o.bounce_mode = BOUNCE_MATERIAL
    o.tri_light_mode = TRI_LIGHT_SOLID_ANGLE
    o.kernel_mode = KERNEL_WAVEFRONT
    for k, v in overrides.items():
        if not hasattr(o, k):
            raise AttributeError("tr_pt_options has no field %r" % k)
        setattr(o, k, v)
    return o


def copy_options(o):
    c = Options()
    C.memmove(C.byref(c), C.byref(o), C.sizeof(Options))
    return c
