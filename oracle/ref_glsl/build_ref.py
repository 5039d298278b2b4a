#!/usr/bin/env python3
"""build_ref.py -- TEST INFRASTRUCTURE ONLY: builds oracle/_ref/, the reference's OWN shader code as a CPU library.

The reference's path tracer is GLSL (shader/*.glsl, *.rgen, *.rahit, ... of /root/reference) that only a Vulkan
ray-tracing device can run. Its text, however, is plain scalar code over vec/mat types, so it compiles as C++ under a
small shim (glsl_shim.hpp: GLSL types, swizzles and built-ins) after a purely LEXICAL rewrite done here:

  * comments, #version / #extension lines dropped; `#include "x.glsl"` -> the rewritten copy;
  * parameter qualifiers: `inout T x`, `out T x` -> `T& x`; `in T x` -> `T x`;
  * floating-point literals get an `f` suffix (GLSL literals are 32-bit);
  * interface blocks `layout(...) buffer|uniform name { T field[]; } inst;` -> a struct with pointer members and a
    variable; opaque uniforms (sampler2D, image2DArray, accelerationStructureEXT), ray payloads and hit attributes ->
    plain (thread-local) variables; `layout(local_size_x = ...) in;` dropped;
  * the array constructor of sobol_lookup_table -> a brace list; every shader's `main` -> `<stage>_main`.

No statement, expression or constant of the shader code is changed, and nothing is copied into the repository: the
sources are read from /root/reference at build time and every output goes to oracle/_ref/ (git-ignored). The harness
around them (ref_harness.inc: buffers, texture unit, traceRayEXT emulated with the oracle's ray caster as the
"hardware", launch loop) is ours. One shared library per shader permutation (the reference compiles one pipeline per
option set, src/path_tracer_stage.cc:45-81, src/rt_common.cc:6-79, src/rt_stage.cc:27-47).

usage: python oracle/ref_glsl/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")

# one entry per translation unit: (unit name, entry shader files in order, prefix for their main())
UNITS = {
    "pt": ["path_tracer.rgen", "rt_common.rahit", "rt_common.rchit", "rt_common.rmiss", "rt_common_shadow.rahit",
           "rt_common_shadow.rchit", "rt_common_shadow.rmiss", "rt_common_point_light.rint", "rt_common_point_light.rchit"],
    "pre": ["pre_transform.comp"],
    "tri": ["extract_tri_lights.comp"],
}

# Shader permutations (name -> preprocessor definitions), mirroring what the stages define.
def variant_defines(v):
    d = {"MAX_BOUNCES": str(v["depth"]), "SAMPLES_PER_PASS": str(v.get("spp_per_pass", 1)), "PRE_TRANSFORMED_VERTICES": "",
         "USE_COLOR_TARGET": "", "USE_DIFFUSE_TARGET": "", "USE_REFLECTION_TARGET": "",
         "CAMERA_PROJECTION_TYPE": str(v.get("projection", 0)), "DISTRIBUTION_STRATEGY": str(v.get("distribution", 0))}
    sampler = v.get("sampler", "uniform")
    if sampler == "sobol_owen":
        d["USE_SOBOL_OWEN_SAMPLING"] = ""
    elif sampler in ("sobol_z2", "sobol_z3"):
        d["USE_SOBOL_Z_ORDER_SAMPLING"] = ""
        d["SOBOL_Z_ORDER_CURVE_DIMS"] = sampler[-1]
    d[{"point": "USE_POINT_FILTER", "box": "USE_BOX_FILTER", "bh": "USE_BLACKMAN_HARRIS_FILTER"}[v.get("film", "point")]] = ""
    mis = v.get("mis", "power")
    if mis == "power":
        d["MIS_POWER_HEURISTIC"] = ""
    elif mis == "balance":
        d["MIS_BALANCE_HEURISTIC"] = ""
    d[{"material": "BOUNCE_MATERIAL", "hemisphere": "BOUNCE_HEMISPHERE", "cosine": "BOUNCE_COSINE_HEMISPHERE"}[v.get("bounce", "material")]] = ""
    d[{"area": "TRI_LIGHT_SAMPLE_AREA", "solid_angle": "TRI_LIGHT_SAMPLE_SOLID_ANGLE", "hybrid": "TRI_LIGHT_SAMPLE_HYBRID"}[v.get("tri", "solid_angle")]] = ""
    for key, macro in (("w_point", "NEE_SAMPLE_POINT_LIGHTS"), ("w_dir", "NEE_SAMPLE_DIRECTIONAL_LIGHTS"),
                       ("w_env", "NEE_SAMPLE_ENVMAP"), ("w_tri", "NEE_SAMPLE_EMISSIVE_TRIANGLES")):
        w = v.get(key, 1.0)
        if w > 0:
            d[macro] = "%.6ff" % w          # std::to_string(float) prints six decimals (src/rt_common.cc:69-78)
    for key, macro in (("rr", "USE_RUSSIAN_ROULETTE"), ("dof", "USE_DEPTH_OF_FIELD"), ("hide_lights", "HIDE_LIGHTS"),
                       ("regularization", "PATH_SPACE_REGULARIZATION"), ("white_albedo", "USE_WHITE_ALBEDO_ON_FIRST_BOUNCE"),
                       ("transparent_background", "USE_TRANSPARENT_BACKGROUND"), ("shadow_terminator", "USE_SHADOW_TERMINATOR_FIX")):
        if v.get(key):
            d[macro] = ""
    return d


VARIANTS = {
    # the bench / CLI-like permutation
    "v0": dict(depth=4, sampler="sobol_owen", tri="solid_angle", bounce="material", mis="power"),
    # CLI defaults of the reference's own golden render (uniform random, depth 8)
    "v1": dict(depth=8, sampler="uniform", tri="solid_angle", bounce="material", mis="power"),
    "v2": dict(depth=3, sampler="sobol_z3", tri="area", bounce="hemisphere", mis="balance", film="box", projection=1),
    "v3": dict(depth=9, sampler="sobol_z2", tri="hybrid", bounce="cosine", mis="none", film="bh", dof=1, rr=1, regularization=1),
    "v4": dict(depth=2, sampler="sobol_owen", tri="solid_angle", bounce="material", mis="power", w_point=0.0, w_dir=0.0, w_env=0.0, w_tri=0.0,
               hide_lights=1, white_albedo=1, transparent_background=1),
}

FLOAT_LITERAL = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def rewrite_block(m):
    name, body, inst, arr = m.group(1), m.group(2), m.group(3), m.group(4)
    body = re.sub(r"(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2;", body)
    decl = "static %s* %s;" % (name, inst) if arr else "static %s %s;" % (name, inst)
    return "struct %s\n{%s};\n%s" % (name, body, decl)


def translate(text, stage_prefix=None):
    text = strip_comments(text)
    text = re.sub(r"^[ \t]*#(version|extension)[^\n]*\n", "\n", text, flags=re.M)
    text = re.sub(r'#include\s+"([^"]+)"', lambda m: '#include "%s.h"' % os.path.basename(m.group(1)), text)
    text = re.sub(r"layout\s*\(\s*local_size_x[^)]*\)\s*in\s*;", "", text)
    # interface blocks and opaque uniforms
    text = re.sub(r"layout\s*\([^)]*\)\s*(?:readonly\s+|writeonly\s+)*(?:buffer|uniform)\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*(\[\s*\])?\s*;",
                  rewrite_block, text, flags=re.S)
    text = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"static \1* \2;", text)
    text = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+(\w+)\s+(\w+)\s*;", r"static \1 \2;", text)
    text = re.sub(r"layout\s*\([^)]*\)\s*rayPayload(?:In)?EXT\s+(\w+)\s+(\w+)\s*;", r"static thread_local \1 \2;", text)
    text = re.sub(r"hitAttributeEXT\s+\w+\s+\w+\s*;", "", text)           # declared once by the harness
    # parameter qualifiers
    text = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"\bin\s+(\w+)\s+(\w+)", r"\1 \2", text)
    # 32-bit float literals
    text = FLOAT_LITERAL.sub(lambda m: m.group(1) + "f", text)
    # array constructor -> brace list (sobol_lookup_table.glsl)
    text = re.sub(r"=\s*uvec4\s*\[\s*\]\s*\((.*?)\)\s*;", lambda m: "= {" + m.group(1) + "};", text, flags=re.S)
    if stage_prefix:
        text = re.sub(r"\bvoid\s+main\s*\(\s*\)", "void %s_main()" % stage_prefix, text)
    return text


def generate(reference, gen_dir):
    shader_dir = os.path.join(reference, "shader")
    os.makedirs(gen_dir, exist_ok=True)
    done = set()

    def emit(name, stage_prefix=None):
        if name in done:
            return
        done.add(name)
        with open(os.path.join(shader_dir, name)) as f:
            src = f.read()
        for inc in re.findall(r'#include\s+"([^"]+)"', strip_comments(src)):
            emit(os.path.basename(inc))
        with open(os.path.join(gen_dir, name + ".h"), "w") as f:
            f.write("// generated by oracle/ref_glsl/build_ref.py from the reference's shader/%s -- do not commit\n" % name)
            f.write(translate(src, stage_prefix))

    for unit, files in UNITS.items():
        for fn in files:
            emit(fn, re.sub(r"\W", "_", fn))
    return sorted(done)


def build(reference="/root/reference", force=False, verbose=False):
    """Returns {variant: path of the shared library}; {} when the reference is not available (GPU box)."""
    libs = {v: os.path.join(OUT, "libref_glsl_%s.so" % v) for v in VARIANTS}
    if not os.path.isdir(os.path.join(reference, "shader")):
        return {v: p for v, p in libs.items() if os.path.exists(p)}
    gen_dir = os.path.join(OUT, "gen")
    stamp_path = os.path.join(OUT, "stamp")
    h = hashlib.sha256()
    for fn in sorted(os.listdir(HERE)):
        with open(os.path.join(HERE, fn), "rb") as f:
            h.update(f.read())
    for fn in sorted(os.listdir(os.path.join(reference, "shader"))):
        if fn.endswith((".glsl", ".rgen", ".rahit", ".rchit", ".rmiss", ".rint", ".comp")):
            with open(os.path.join(reference, "shader", fn), "rb") as f:
                h.update(f.read())
    stamp = h.hexdigest()
    if not force and os.path.exists(stamp_path) and open(stamp_path).read() == stamp and all(os.path.exists(p) for p in libs.values()):
        return libs
    generate(reference, gen_dir)
    procs = []
    for v, spec in VARIANTS.items():
        defs = ["-D%s=%s" % (k, val) if val != "" else "-D" + k for k, val in variant_defines(spec).items()]
        cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-w", "-fpermissive",
               "-I", HERE, "-I", gen_dir, "-DREF_VARIANT_NAME=\"%s\"" % v] + defs + \
              [os.path.join(HERE, "ref_unit_%s.cc" % u) for u in ("pt", "pre", "tri", "fn")] + ["-o", libs[v], "-lpthread"]
        if verbose:
            print(" ".join(cmd))
        procs.append((v, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for v, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("variant %s failed:\n%s\n" % (v, out[-6000:]))
    if failed:
        raise RuntimeError("oracle/_ref build failed")
    with open(stamp_path, "w") as f:
        f.write(stamp)
    return libs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args()
    for v, p in build(a.reference, a.force, a.verbose).items():
        print(v, p)
