"""The C-ABI library loads and exports every symbol include/tauray_b200.h declares; struct layouts of the
ctypes mirror equal the C compiler's; without a GPU the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_text():
    with open(os.path.join(ROOT, "include", "tauray_b200.h")) as f:
        return f.read()


def declared_functions():
    return sorted(set(re.findall(r"\b(tr_pt_[a-z_0-9]+)\s*\(", header_text())))


def test_library_exports_every_declared_symbol():
    from tauray_b200 import library
    lib = library.lib()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libtauray_b200.so does not export %s" % n
    assert set(names) == set(library.EXPORTS), (set(names) ^ set(library.EXPORTS))
    from tauray_b200 import abi
    assert lib.tr_pt_abi_version() == abi.ABI_VERSION == 2


def test_every_entry_point_cites_the_reference():
    text = header_text()
    assert text.count("src/") > 40 and "shader/" in text


def test_struct_layouts_match_the_c_compiler(tmp_path):
    from tauray_b200 import abi
    src = tmp_path / "sizes.c"
    names = ["tr_vertex", "tr_material", "tr_instance", "tr_directional_light", "tr_point_light", "tr_tri_light",
             "tr_camera_data", "tr_camera_pair", "tr_alias_table_entry", "tr_mesh_ref", "tr_texture", "tr_pt_scene_desc",
             "tr_pt_options", "tr_pt_distribution", "tr_pt_stats", "tr_pt_accumulation_state"]
    body = "".join('printf("%s %%zu\\n", sizeof(%s));' % (n, n) for n in names)
    body += 'printf("off_stats_render_ms %zu\\n", offsetof(tr_pt_stats, render_ms));'
    body += 'printf("off_opt_kernel_mode %zu\\n", offsetof(tr_pt_options, kernel_mode));'
    body += 'printf("off_desc_envproj %zu\\n", offsetof(tr_pt_scene_desc, environment_proj));'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "tauray_b200.h"\nint main(){%s return 0;}' % body)
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    want = {
        "tr_vertex": abi.VERTEX.itemsize, "tr_material": abi.MATERIAL.itemsize, "tr_instance": abi.INSTANCE.itemsize,
        "tr_directional_light": abi.DIRECTIONAL_LIGHT.itemsize, "tr_point_light": abi.POINT_LIGHT.itemsize,
        "tr_tri_light": abi.TRI_LIGHT.itemsize, "tr_camera_data": abi.CAMERA_DATA.itemsize,
        "tr_camera_pair": abi.CAMERA_PAIR.itemsize, "tr_mesh_ref": C.sizeof(abi.MeshRef), "tr_texture": C.sizeof(abi.Texture),
        "tr_pt_scene_desc": C.sizeof(abi.SceneDesc), "tr_pt_options": C.sizeof(abi.Options),
        "tr_pt_distribution": C.sizeof(abi.Distribution), "tr_pt_stats": C.sizeof(abi.Stats),
        "tr_pt_accumulation_state": C.sizeof(abi.AccumulationState),
        "off_stats_render_ms": abi.Stats.render_ms.offset, "off_opt_kernel_mode": abi.Options.kernel_mode.offset,
        "off_desc_envproj": abi.SceneDesc.environment_proj.offset,
    }
    for k, v in want.items():
        assert int(out[k]) == v, (k, out[k], v)
    # the reference's wire formats
    assert (int(out["tr_vertex"]), int(out["tr_material"]), int(out["tr_instance"])) == (48, 80, 288)
    assert (int(out["tr_directional_light"]), int(out["tr_point_light"]), int(out["tr_tri_light"])) == (32, 64, 64)
    assert (int(out["tr_camera_data"]), int(out["tr_camera_pair"]), int(out["tr_alias_table_entry"])) == (320, 640, 16)


def test_default_options_are_the_cli_defaults():
    from tauray_b200 import abi, library
    o = abi.Options()
    library.lib().tr_pt_default_options(C.byref(o))
    d = abi.default_options()
    for name, _ in abi.Options._fields_:
        if name == "reserved":
            continue
        assert getattr(o, name) == pytest.approx(getattr(d, name)), name
    # src/options.hh defaults
    assert (o.max_ray_depth, o.samples_per_pixel, o.samples_per_pass) == (8, 1, 1)
    assert o.min_ray_dist == pytest.approx(1e-4) and o.film_radius == pytest.approx(0.5)
    assert o.local_sampler == abi.SAMPLER_UNIFORM_RANDOM and o.film == abi.FILM_POINT
    assert o.mis_mode == abi.MIS_POWER_HEURISTIC and o.bounce_mode == abi.BOUNCE_MATERIAL
    assert o.tri_light_mode == abi.TRI_LIGHT_SOLID_ANGLE and o.russian_roulette_delta == 0


def test_invalid_options_are_rejected_before_touching_the_gpu():
    from tauray_b200 import abi, library
    lib = library.lib()
    ctx = C.c_void_p()
    bad = abi.default_options(max_ray_depth=0)
    assert lib.tr_pt_create(C.byref(ctx), 0, C.byref(bad)) == abi.ERR_INVALID_ARGUMENT
    unknown_sampler = abi.default_options(local_sampler=4)
    assert lib.tr_pt_create(C.byref(ctx), 0, C.byref(unknown_sampler)) == abi.ERR_INVALID_ARGUMENT
    assert b"sampler" in lib.tr_pt_last_error(None)
    assert lib.tr_pt_create(None, 0, C.byref(bad)) == abi.ERR_INVALID_ARGUMENT


def test_no_gpu_means_a_loud_failure_not_a_fallback():
    """Product path without a device: tr_pt_create must fail with TR_ERR_NO_DEVICE."""
    code = (
        "import sys, ctypes as C; sys.path.insert(0, %r)\n"
        "from tauray_b200 import abi, library\n"
        "ctx = C.c_void_p(); o = abi.default_options()\n"
        "rc = library.lib().tr_pt_create(C.byref(ctx), 0, C.byref(o))\n"
        "print(rc, library.lib().tr_pt_last_error(None).decode())\n" % ROOT)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout
    assert out.startswith("-2 "), out
    assert "no CPU fallback" in out


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tauray_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".hh")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in text and "liboracle" not in text and "oracle/" not in text.replace("oracle/README", ""), f
