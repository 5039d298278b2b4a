"""Minimal binary glTF 2.0 (.glb) reader -> Scene tables.

Restates the parts of the reference's loader that the hot path depends on
(src/gltf.cc:199-280 materials incl. TR_data / KHR_materials_transmission / ior /
emissive_strength, :319-503 node walk with KHR_lights_punctual + TR_data light angle/radius and
cameras, :654-751 primitives) and scene_stage's instance ordering (src/scene_stage.cc:664-819:
static meshes on static nodes in entity = node-visit order, one instance per vertex group).
Only what `test/test.glb`-class assets need: no skins, animations, sparse accessors or samplers
other than linear/repeat. Used by tests and tools, never by the render path.
"""
import json
import math
import struct

import numpy as np

from . import abi
from .scene import Scene, make_material, perspective_camera

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


def _read_chunks(data):
    magic, version, length = struct.unpack_from("<III", data, 0)
    if magic != 0x46546C67:
        raise ValueError("not a GLB file")
    off = 12
    js, binary = None, b""
    while off < length:
        clen, ctype = struct.unpack_from("<II", data, off)
        chunk = data[off + 8: off + 8 + clen]
        if ctype == 0x4E4F534A:
            js = json.loads(chunk.decode("utf-8"))
        elif ctype == 0x004E4942:
            binary = chunk
        off += 8 + clen
    return js, binary


def _accessor(js, binary, index):
    acc = js["accessors"][index]
    bv = js["bufferViews"][acc["bufferView"]]
    dt = np.dtype(_COMPONENT[acc["componentType"]])
    ncomp = _NCOMP[acc["type"]]
    start = bv.get("byteOffset", 0) + acc.get("byteOffset", 0)
    stride = bv.get("byteStride", 0) or dt.itemsize * ncomp
    count = acc["count"]
    raw = np.frombuffer(binary, dtype=np.uint8, count=stride * (count - 1) + dt.itemsize * ncomp, offset=start)
    out = np.lib.stride_tricks.as_strided(raw, shape=(count, dt.itemsize * ncomp), strides=(stride, 1))
    return np.ascontiguousarray(out).view(dt).reshape(count, ncomp)


def _quat_to_mat(q):
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ], dtype=np.float64)


def _node_matrix(node):
    if "matrix" in node:
        return np.asarray(node["matrix"], dtype=np.float64).reshape(4, 4).T
    m = np.eye(4)
    r = _quat_to_mat(node.get("rotation", [0, 0, 0, 1]))
    s = np.asarray(node.get("scale", [1, 1, 1]), dtype=np.float64)
    m[:3, :3] = r * s[None, :]
    m[:3, 3] = node.get("translation", [0, 0, 0])
    return m


def _global_direction(m, forward=(0, 0, -1)):
    """transformable::get_global_direction (src/transformable.cc:399-404) via
    get_matrix_orientation (src/math.cc:44-52)."""
    r = m[:3, :3].copy()
    r /= np.linalg.norm(r, axis=0, keepdims=True)
    d = r @ np.asarray(forward, dtype=np.float64)
    return d / np.linalg.norm(d)


def _decode_image(blob):
    import cv2
    img = cv2.imdecode(np.frombuffer(blob, dtype=np.uint8), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise ValueError("cannot decode embedded image")
    if img.ndim == 2:
        img = np.stack([img] * 3 + [np.full_like(img, 255)], axis=-1)
    elif img.shape[2] == 3:
        img = np.concatenate([img[..., ::-1], np.full(img.shape[:2] + (1,), 255, img.dtype)], axis=-1)
    else:
        img = img[..., [2, 1, 0, 3]]
    if img.dtype != np.uint8:
        img = (img.astype(np.float32) / 257.0 + 0.5).astype(np.uint8)
    return np.ascontiguousarray(img)


def _material(js, m, tex_of):
    pbr = m.get("pbrMetallicRoughness", {})
    ext = m.get("extensions", {})

    def tex(info):
        return tex_of(info["index"]) if info else -1

    emission = np.asarray(m.get("emissiveFactor", [0, 0, 0]), dtype=np.float64)
    discard_tr_emission = False
    if "KHR_materials_emissive_strength" in ext and "emissiveStrength" in ext["KHR_materials_emissive_strength"]:
        emission = emission * ext["KHR_materials_emissive_strength"]["emissiveStrength"]
        discard_tr_emission = True
    transmittance, ior = 0.0, 1.45
    tr = pbr.get("extensions", {}).get("TR_data")
    if tr:
        transmittance = tr.get("transmission", transmittance)
        ior = tr.get("ior", ior)
        if not discard_tr_emission and "emission" in tr:
            emission = np.asarray(tr["emission"][:3], dtype=np.float64)
    if "KHR_materials_transmission" in ext:
        transmittance = ext["KHR_materials_transmission"].get("transmissionFactor", transmittance)
    if "KHR_materials_ior" in ext:
        ior = ext["KHR_materials_ior"].get("ior", ior)
    return make_material(
        albedo=pbr.get("baseColorFactor", [1, 1, 1, 1]),
        metallic=pbr.get("metallicFactor", 1.0),
        roughness=pbr.get("roughnessFactor", 1.0),
        emission=emission, transmittance=transmittance, ior=ior,
        double_sided=m.get("doubleSided", False),
        albedo_tex=tex(pbr.get("baseColorTexture")),
        metallic_roughness_tex=tex(pbr.get("metallicRoughnessTexture")),
        normal_tex=tex(m.get("normalTexture")),
        emission_tex=tex(m.get("emissiveTexture")),
    )


class GlbScene:
    """Result of :func:`load_glb`: the flattened `scene`, the scene cameras and their transforms."""

    def __init__(self):
        self.scene = Scene()
        self.cameras = []      # dicts: transform, yfov_deg, aspect, near, far
        self.instance_names = []

    def camera_data(self, index=0, aspect=None):
        c = self.cameras[index]
        return perspective_camera(c["transform"], c["yfov_deg"], aspect or c["aspect"], c["near"], c["far"])


def load_glb(path):
    with open(path, "rb") as f:
        data = f.read()
    js, binary = _read_chunks(data)
    out = GlbScene()
    scene = out.scene

    images = {}

    def tex_of(texture_index):
        src = js["textures"][texture_index]["source"]
        if src not in images:
            bv = js["bufferViews"][js["images"][src]["bufferView"]]
            blob = binary[bv.get("byteOffset", 0): bv.get("byteOffset", 0) + bv["byteLength"]]
            images[src] = scene.add_texture(_decode_image(blob))
        return images[src]

    materials = [_material(js, m, tex_of) for m in js.get("materials", [])]
    default_material = make_material()

    meshes = []
    for mesh in js.get("meshes", []):
        groups = []
        for prim in mesh["primitives"]:
            attr = prim["attributes"]
            pos = _accessor(js, binary, attr["POSITION"]).astype(np.float32)
            v = np.zeros(pos.shape[0], dtype=abi.VERTEX)
            v["pos"] = pos
            if "NORMAL" in attr:
                v["normal"] = _accessor(js, binary, attr["NORMAL"]).astype(np.float32)
            if "TEXCOORD_0" in attr:
                v["uv"] = _accessor(js, binary, attr["TEXCOORD_0"]).astype(np.float32)
            if "TANGENT" in attr:
                v["tangent"] = _accessor(js, binary, attr["TANGENT"]).astype(np.float32)
            if "indices" in prim:
                idx = _accessor(js, binary, prim["indices"]).astype(np.uint32).reshape(-1)
            else:
                idx = np.arange(pos.shape[0], dtype=np.uint32)
            mat = materials[prim["material"]] if prim.get("material", -1) >= 0 else default_material
            groups.append((v, idx, mat))
        meshes.append((mesh.get("name", ""), groups))

    lights = js.get("extensions", {}).get("KHR_lights_punctual", {}).get("lights", [])

    # src/gltf.cc:319-345: meta (light angle/radius) is a reference shared by the whole traversal,
    # so a value set on one node stays for the nodes visited later.
    shared_meta = {}

    def walk_shared(node_index, parent):
        node = js["nodes"][node_index]
        tr = node.get("extensions", {}).get("TR_data")
        if tr and "light" in tr:
            if "angle" in tr["light"]:
                shared_meta["light_angle"] = tr["light"]["angle"]
            if "radius" in tr["light"]:
                shared_meta["light_radius"] = tr["light"]["radius"]
        walk_one(node_index, parent)

    def walk_one(node_index, parent):
        node = js["nodes"][node_index]
        m = parent @ _node_matrix(node)
        if "mesh" in node:
            name, groups = meshes[node["mesh"]]
            for v, idx, mat in groups:
                scene.add_instance(v, idx, mat, transform=m)
                out.instance_names.append(name)
        if "camera" in node:
            c = js["cameras"][node["camera"]]
            if c["type"] == "perspective":
                p = c["perspective"]
                out.cameras.append(dict(transform=m, yfov_deg=math.degrees(p["yfov"]), aspect=p.get("aspectRatio", 1.0),
                                        near=p["znear"], far=p.get("zfar", 1e9)))
        kl = node.get("extensions", {}).get("KHR_lights_punctual")
        if kl is not None:
            l = lights[kl["light"]]
            color = np.asarray(l.get("color", [1, 1, 1]), dtype=np.float64) * l.get("intensity", 1.0)
            pos = m[:3, 3]
            if l["type"] == "directional":
                scene.add_directional_light(color, _global_direction(m), math.degrees(shared_meta.get("light_angle", 0.0)))
            elif l["type"] == "point":
                scene.add_point_light(color / (4 * math.pi), pos, radius=shared_meta.get("light_radius", 0.0))
            elif l["type"] == "spot":
                outer = math.degrees(l["spot"].get("outerConeAngle", math.pi / 4))
                inner = math.degrees(l["spot"].get("innerConeAngle", 0.0))
                if inner <= 0:
                    falloff = 1.0
                else:
                    ci = math.cos(math.radians(inner))
                    co = math.cos(math.radians(outer))
                    falloff = math.log(4 / 255.0) / math.log(max(1.0 - ci, 0.0) / (1.0 - co))
                scene.add_spotlight(color / (4 * math.pi), pos, _global_direction(m), outer, falloff,
                                    radius=shared_meta.get("light_radius", 0.0))
        for child in node.get("children", []):
            walk_shared(child, m)

    for root in js["scenes"][js.get("scene", 0)]["nodes"]:
        walk_shared(root, np.eye(4))
    return out
