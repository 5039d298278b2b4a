"""Host-side scene tables: what Tauray's ``scene_stage`` flattens for the tracer.

This is the Python stand-in (tests / bench only) for the part of the reference that stays host
C++ in the north-star picture: it fills the same tables ``scene_stage::update`` fills
(src/scene_stage.cc:1066-1116 instances, :1277-1354 lights and metadata) and the camera uniform
``camera::write_uniform_buffer`` writes (src/camera.cc:431-462), and hands them to the C ABI.
"""
import ctypes as C
import math

import numpy as np

from . import abi


def identity4():
    return np.eye(4, dtype=np.float32)


def make_material(albedo=(1, 1, 1, 1), metallic=0.0, roughness=1.0, emission=(0, 0, 0), transmittance=0.0,
                  ior=1.45, double_sided=True, albedo_tex=-1, metallic_roughness_tex=-1, normal_tex=-1,
                  emission_tex=-1, normal_factor=1.0):
    """material defaults: src/material.hh:12-39; packing: src/scene_stage.cc:1098-1115."""
    m = np.zeros((), dtype=abi.MATERIAL)
    m["albedo_factor"] = albedo
    m["metallic_roughness_factor"] = (metallic, roughness, 0, 0)
    m["emission_factor"] = tuple(emission) + (0.0,)
    m["transmittance"] = transmittance
    m["ior"] = ior
    m["normal_factor"] = normal_factor
    m["flags"] = 1 if double_sided else 0
    m["albedo_tex_id"] = albedo_tex
    m["metallic_roughness_tex_id"] = metallic_roughness_tex
    m["normal_tex_id"] = normal_tex
    m["emission_tex_id"] = emission_tex
    return m


class Scene:
    """Flattened scene: one entry per (model, vertex group), in scene_stage's instance order."""

    def __init__(self):
        self.materials = []
        self.transforms = []
        self.prev_transforms = {}  # instance index -> previous-frame transform (default: the current one)
        self.meshes = []          # (vertices[VERTEX], indices[u32])
        self.opaque = []
        self.point_lights = np.zeros(0, dtype=abi.POINT_LIGHT)
        self.directional_lights = np.zeros(0, dtype=abi.DIRECTIONAL_LIGHT)
        self.textures = []        # (H, W, 4) uint8 or float32
        self.environment_factor = (0.0, 0.0, 0.0, 0.0)
        self.environment_proj = -1
        self.environment_map = None   # (H, W, 4) float16 / float32 lat-long texels
        self._keep = None

    # -- building -------------------------------------------------------------------------
    def set_environment_map(self, texels, factor=(1.0, 1.0, 1.0)):
        """environment_map(path, LAT_LONG, factor): texels (H, W, 3|4) -> RGBA16F like the reference's
        .hdr textures (src/texture.cc:498-510); scene_metadata gets vec4(factor, 1) (scene_stage.cc:1345)."""
        t = np.asarray(texels)
        if t.shape[2] == 3:
            t = np.concatenate([t, np.ones(t.shape[:2] + (1,), dtype=t.dtype)], axis=2)
        if t.dtype != np.float32:
            t = t.astype(np.float16)
        self.environment_map = np.ascontiguousarray(t)
        self.environment_factor = (float(factor[0]), float(factor[1]), float(factor[2]), 1.0)
        self.environment_proj = abi.ENVMAP_LAT_LONG

    def add_instance(self, vertices, indices, material, transform=None, opaque=None):
        vertices = np.ascontiguousarray(vertices, dtype=abi.VERTEX)
        indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        assert indices.size % 3 == 0
        if opaque is None:
            opaque = not self.potentially_transparent(material)
        self.meshes.append((vertices, indices))
        self.materials.append(material)
        self.transforms.append(identity4() if transform is None else np.asarray(transform, dtype=np.float32))
        self.opaque.append(1 if opaque else 0)
        return len(self.meshes) - 1

    def potentially_transparent(self, material):
        """material::potentially_transparent, src/material.cc:7-11."""
        if float(material["transmittance"]) > 0.0 or float(material["albedo_factor"][3]) < 1.0:
            return True
        tid = int(material["albedo_tex_id"])
        if tid >= 0:
            tex = self.textures[tid]
            if tex.dtype == np.uint8:
                return bool((tex[..., 3] < 255).any())
            return bool((tex[..., 3] < 1.0).any())
        return False

    def add_texture(self, pixels):
        pixels = np.ascontiguousarray(pixels)
        assert pixels.ndim == 3 and pixels.shape[2] == 4 and pixels.dtype in (np.uint8, np.float32)
        self.textures.append(pixels)
        return len(self.textures) - 1

    def add_directional_light(self, color, direction, angle_deg=0.0):
        """directional_light_entry, src/scene_stage.cc:46-62 (dir = normalised global -Z of the node)."""
        l = np.zeros(1, dtype=abi.DIRECTIONAL_LIGHT)
        d = np.asarray(direction, dtype=np.float64)
        d = d / np.linalg.norm(d)
        l["color"] = color
        l["shadow_map_index"] = -1
        l["dir"] = d.astype(np.float32)
        l["dir_cutoff"] = np.float32(math.cos(math.radians(angle_deg)))
        self.directional_lights = np.concatenate([self.directional_lights, l])

    def add_point_light(self, color, pos, radius=0.0, cutoff_radius=0.0):
        """point_light_entry(point_light), src/scene_stage.cc:66-73."""
        l = np.zeros(1, dtype=abi.POINT_LIGHT)
        l["color"] = color
        l["pos"] = pos
        l["radius"] = radius
        l["cutoff_radius"] = cutoff_radius
        l["spot_radius"] = -1.0
        l["shadow_map_index"] = -1
        self.point_lights = np.concatenate([self.point_lights, l])

    def add_spotlight(self, color, pos, direction, cutoff_angle_deg, falloff_exponent, radius=0.0,
                      cutoff_radius=0.0):
        """point_light_entry(spotlight), src/scene_stage.cc:75-87."""
        l = np.zeros(1, dtype=abi.POINT_LIGHT)
        d = np.asarray(direction, dtype=np.float64)
        d = d / np.linalg.norm(d)
        l["color"] = color
        l["dir"] = d.astype(np.float32)
        l["pos"] = pos
        l["radius"] = radius
        l["dir_cutoff"] = np.float32(math.cos(math.radians(cutoff_angle_deg)))
        l["dir_falloff"] = falloff_exponent
        l["cutoff_radius"] = cutoff_radius
        l["spot_radius"] = cutoff_radius * math.tan(math.radians(cutoff_angle_deg))
        l["shadow_map_index"] = -1
        self.point_lights = np.concatenate([self.point_lights, l])

    # -- flattening -----------------------------------------------------------------------
    @property
    def triangle_count(self):
        return sum(i.size // 3 for _, i in self.meshes)

    def instance_table(self):
        """instance_buffer fill loop, src/scene_stage.cc:1066-1116 (incl. the running
        tri_light_count that becomes light_base_id)."""
        n = len(self.meshes)
        inst = np.zeros(n, dtype=abi.INSTANCE)
        tri_light_count = 0
        for i in range(n):
            mat = self.materials[i]
            if np.any(np.asarray(mat["emission_factor"][:3]) != 0):
                inst[i]["light_base_id"] = tri_light_count
                tri_light_count += self.meshes[i][1].size // 3
            else:
                inst[i]["light_base_id"] = -1
            inst[i]["sh_grid_index"] = -1
            inst[i]["shadow_terminator_mul"] = 1.0
            m = self.transforms[i].astype(np.float64)
            inst[i]["model"] = abi.mat_to_wire(m)
            inst[i]["model_normal"] = abi.mat_to_wire(np.linalg.inv(m).T)
            inst[i]["model_prev"] = abi.mat_to_wire(np.asarray(self.prev_transforms.get(i, m), dtype=np.float64))
            inst[i]["mat"] = mat
        return inst

    def desc(self):
        """Build the tr_pt_scene_desc; the returned object keeps every buffer alive."""
        n = len(self.meshes)
        inst = self.instance_table()
        mesh_refs = (abi.MeshRef * max(n, 1))()
        for i, (v, idx) in enumerate(self.meshes):
            mesh_refs[i].vertices = v.ctypes.data
            mesh_refs[i].indices = idx.ctypes.data
            mesh_refs[i].vertex_count = v.shape[0]
            mesh_refs[i].index_count = idx.size
            mesh_refs[i].opaque = self.opaque[i]
        texs = (abi.Texture * max(len(self.textures), 1))()
        for i, t in enumerate(self.textures):
            texs[i].data = t.ctypes.data
            texs[i].width = t.shape[1]
            texs[i].height = t.shape[0]
            texs[i].format = {np.dtype(np.uint8): abi.TEXTURE_RGBA8_UNORM, np.dtype(np.float16): abi.TEXTURE_RGBA16F,
                              np.dtype(np.float32): abi.TEXTURE_RGBA32F}[t.dtype]
        d = abi.SceneDesc()
        d.instance_count = n
        d.point_light_count = self.point_lights.shape[0]
        d.directional_light_count = self.directional_lights.shape[0]
        d.texture_count = len(self.textures)
        d.instances = inst.ctypes.data
        d.meshes = mesh_refs
        d.point_lights = self.point_lights.ctypes.data if self.point_lights.size else None
        d.directional_lights = self.directional_lights.ctypes.data if self.directional_lights.size else None
        d.textures = texs
        d.environment_factor = (C.c_float * 4)(*self.environment_factor)
        d.environment_proj = self.environment_proj
        env = None
        if self.environment_proj >= 0:
            e = self.environment_map
            assert e is not None and e.ndim == 3 and e.shape[2] == 4 and e.dtype in (np.float16, np.float32)
            e = np.ascontiguousarray(e)
            env = abi.Texture()
            env.data = e.ctypes.data
            env.width, env.height = e.shape[1], e.shape[0]
            env.format = abi.TEXTURE_RGBA16F if e.dtype == np.float16 else abi.TEXTURE_RGBA32F
            d.environment_map = C.pointer(env)
            env = (env, e)
        d._keep = (inst, mesh_refs, texs, self, env)
        return d


# -- cameras --------------------------------------------------------------------------------------
def look_at_transform(eye, target, up=(0, 1, 0)):
    """Camera-to-world transform of a camera at `eye` looking down its -Z axis at `target`."""
    eye = np.asarray(eye, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, np.asarray(up, dtype=np.float64))
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = s, u, -f, eye
    return m


def perspective_camera(transform, fov_deg, aspect, near=0.1, far=100.0, focus=(1.0, 0.0, 0.0, 0.0),
                       pan=(0.0, 0.0)):
    """One tr_camera_data for a perspective camera.

    camera::refresh (src/camera.cc:358-379: glm::perspective with GLM_FORCE_DEPTH_ZERO_TO_ONE,
    right-handed) + camera::write_uniform_buffer (src/camera.cc:431-462) +
    camera::get_projection_info (src/camera.cc:323-356). `fov_deg` is the vertical field of view.
    """
    f32 = np.float32
    rad = math.radians(fov_deg)
    t = math.tan(rad / 2.0)
    proj = np.zeros((4, 4), dtype=np.float64)
    proj[0, 0] = 1.0 / (aspect * t)
    proj[1, 1] = 1.0 / t
    proj[2, 2] = far / (near - far)
    proj[3, 2] = -1.0
    proj[2, 3] = -(far * near) / (far - near)
    proj[0, 2] = pan[0]
    proj[1, 2] = pan[1]
    inv_view = np.asarray(transform, dtype=np.float64)
    view = np.linalg.inv(inv_view)
    cam = np.zeros((), dtype=abi.CAMERA_DATA)
    cam["view"] = abi.mat_to_wire(view)
    cam["view_inverse"] = abi.mat_to_wire(inv_view)
    cam["view_proj"] = abi.mat_to_wire(proj @ view)
    cam["proj_inverse"] = abi.mat_to_wire(np.linalg.inv(proj))
    cam["origin"] = (inv_view @ np.array([0, 0, 0, 1.0])).astype(f32)
    cam["dof_params"] = focus
    info_w = 2 * t
    cam["projection_info"] = (near * far / (near - far), (near + far) / (near - far), info_w * aspect, info_w)
    cam["pan"] = (pan[0], pan[1], 0, 0)
    return cam


def orthographic_camera(transform, left, right, bottom, top, near=0.1, far=100.0):
    """One tr_camera_data for an orthographic camera: camera::ortho + refresh (src/camera.cc:43-50,384-389:
    glm::ortho, right-handed, depth 0..1), get_projection_info (src/camera.cc:348-356)."""
    f32 = np.float32
    proj = np.zeros((4, 4), dtype=np.float64)
    proj[0, 0] = 2.0 / (right - left)
    proj[1, 1] = 2.0 / (top - bottom)
    proj[2, 2] = -1.0 / (far - near)
    proj[0, 3] = -(right + left) / (right - left)
    proj[1, 3] = -(top + bottom) / (top - bottom)
    proj[2, 3] = -near / (far - near)
    proj[3, 3] = 1.0
    inv_view = np.asarray(transform, dtype=np.float64)
    view = np.linalg.inv(inv_view)
    cam = np.zeros((), dtype=abi.CAMERA_DATA)
    cam["view"] = abi.mat_to_wire(view)
    cam["view_inverse"] = abi.mat_to_wire(inv_view)
    cam["view_proj"] = abi.mat_to_wire(proj @ view)
    cam["proj_inverse"] = abi.mat_to_wire(np.linalg.inv(proj))
    cam["origin"] = (inv_view @ np.array([0, 0, 0, 1.0])).astype(f32)
    cam["dof_params"] = (1.0, 0.0, 0.0, 0.0)
    cam["projection_info"] = (far - near, -far, right - left, top - bottom)
    cam["pan"] = (0, 0, 0, 0)
    return cam


def camera_pairs(cameras):
    """tr_camera_pair[] with previous = current (a still frame)."""
    pairs = np.zeros(len(cameras), dtype=abi.CAMERA_PAIR)
    for i, c in enumerate(cameras):
        pairs[i]["current"] = c
        pairs[i]["previous"] = c
    return pairs


def camera_grid(transform, fov_deg, aspect, grid_w, grid_h, dx, dy, near=0.1, far=100.0):
    """`--camera-grid=w,h,x,y` with recentering distance = infinity (src/tauray.cc:680-727): the
    grid cameras are offset copies of the scene camera in its local x/y plane."""
    cams = []
    base = np.asarray(transform, dtype=np.float64)
    for y in range(grid_h):
        for x in range(grid_w):
            off = np.eye(4)
            off[0, 3] = (x - (grid_w - 1) * 0.5) * dx
            off[1, 3] = ((grid_h - 1) * 0.5 - y) * dy
            cams.append(perspective_camera(base @ off, fov_deg, aspect, near, far))
    return cams
