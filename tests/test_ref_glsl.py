"""The oracle's restatement against the reference's OWN shader text.

oracle/_ref (built from /root/reference by oracle/ref_glsl/build_ref.py; see its docstring) is the reference's path
tracer -- path_tracer.rgen with everything it includes, its hit / any-hit / miss / intersection shaders,
pre_transform.comp, extract_tri_lights.comp -- compiled for the CPU under a GLSL-as-C++ shim. Two levels:

  * function level: 44 functions of math.glsl / sampling.glsl / ggx.glsl / light.glsl / camera.glsl / color.glsl /
    material.glsl / alias_table.glsl on random inputs, for five shader permutations; integer results (PCG, Sobol, Owen
    scrambling, Morton, RGB9E5, sampler streams) bit for bit, floating point within 1e-5 relative (tests/ref_compare.py
    states the allowance for inputs on a branch or a cancellation);
  * image level: whole frames of the reference pipeline (raygen -> traceRayEXT -> any-hit -> evaluate_ray -> accumulate)
    against the oracle's frames, ray counts included, with the oracle's BVH as the ray caster for triangle candidates.

Where /root/reference (and so oracle/_ref) is absent the same comparisons run against fixtures written from the live
library by tools/make_ref_fixtures.py (tests/golden/ref_glsl_*.npz)."""
import os

import numpy as np
import pytest

from tauray_b200 import abi, distribution, testscenes
from oracle.oracle_py import Oracle
from oracle.ref_glsl import ref_py

import ref_cases as rc
from ref_compare import compare_function, compare_images

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LIVE = ref_py.available()
live = pytest.mark.skipif(not LIVE, reason="oracle/_ref is built from /root/reference, which is not available here")

IMAGE_CASES = {
    # name: (variant, scene factory, cameras factory, size, extra options)
    "cornell_sobol_owen": ("v0", lambda: testscenes.cornell(), lambda a: testscenes.cornell_cameras(a), (64, 48), dict(samples_per_pixel=2)),
    "cornell_cli_defaults_depth8": ("v1", lambda: testscenes.cornell(), lambda a: testscenes.cornell_cameras(a), (48, 36), dict(samples_per_pixel=1)),
    "courtyard_envmap": ("v1", lambda: testscenes.courtyard(emissive=True, directional=True), lambda a: testscenes.courtyard_cameras(a), (48, 36), dict(samples_per_pixel=2)),
    "cornell_z3_area_hemisphere_box_ortho": ("v2", lambda: testscenes.cornell(), None, (48, 36), dict(samples_per_pixel=2)),
    "cornell_z2_hybrid_cosine_bh_dof_rr_reg": ("v3", lambda: testscenes.cornell(), lambda a: testscenes.cornell_cameras(a), (48, 36), dict(samples_per_pixel=2)),
    "cornell_no_nee_hidden_lights": ("v4", lambda: testscenes.cornell(), lambda a: testscenes.cornell_cameras(a), (48, 36), dict(samples_per_pixel=2)),
}


# (fraction of pixels within 1e-4, relative difference of the image means). The courtyard has a glass block of
# roughness 0.05 (alpha = 0.0025): its GGX term cancels catastrophically in binary32, so depth-8 paths through it
# diverge between any two correct implementations (27 of 1728 pixels here, a few of them fireflies)
IMAGE_LIMITS = {"courtyard_envmap": (0.975, 2e-2)}
DEFAULT_LIMITS = (0.985, 1e-3)


def image_case(name):
    variant, scene_f, cams_f, size, extra = IMAGE_CASES[name]
    scene = scene_f()
    if cams_f is None:
        from tauray_b200.scene import camera_pairs, look_at_transform, orthographic_camera
        cams = camera_pairs([orthographic_camera(look_at_transform((0.3, 0.2, 5.5), (0, -0.2, 0)), -2.2, 2.2, -1.65, 1.65, 0.1, 100.0)])
    else:
        cams = cams_f(size[0] / size[1])
    opt = abi.default_options(**dict(rc.VARIANTS[variant], **extra))
    return variant, scene, cams, size, opt


def oracle_frames(scene, cams, size, opt, frames):
    o = Oracle(opt)
    o.set_scene(scene)
    o.set_cameras(cams)
    o.set_distribution(distribution.make_distribution(size, abi.DISTRIBUTION_DUPLICATE))
    for _ in range(frames):
        o.render()
    return o, (o.read_color() if frames else None), o.stats()


# ---- function level ----------------------------------------------------------------------------------------
def tangent_space_records(rng, n, ref_fn):
    """function 19 needs an orthonormal basis: taken from create_tangent_space (function 9) of the same side"""
    normals = rc.cat(rc.direction(rng, n))
    tbn = ref_fn(9, normals)
    return rc.cat(rc.direction(rng, n), tbn)


@live
@pytest.mark.parametrize("variant", sorted(rc.VARIANTS))
def test_functions_against_the_reference_shader_text(variant):
    o = Oracle(abi.default_options(**rc.VARIANTS[variant]))
    n = 20000
    failures = []
    for fn, (name, kinds, gen) in sorted(rc.FUNCTIONS.items()):
        rng = np.random.RandomState(1000 + fn)
        rec = gen(rng, n, variant) if gen else tangent_space_records(rng, n, lambda f, r: ref_py.function(variant, f, r))
        wi, wo = ref_py.function_words(variant, fn)
        assert rec.shape[1] == wi and len(kinds) == wo, (fn, rec.shape, wi, wo)
        msgs = compare_function(fn, kinds, ref_py.function(variant, fn, rec), o.function(fn, rec, wo))
        failures += ["%s (%d): %s" % (name, fn, m) for m in msgs]
    assert not failures, "\n".join(failures)


@pytest.mark.parametrize("variant", ["v0", "v2", "v3"])
def test_functions_against_committed_reference_outputs(variant):
    """The same, from tests/golden/ref_glsl_functions_<variant>.npz (inputs and the reference shader's outputs)."""
    fx = np.load(os.path.join(GOLDEN, "ref_glsl_functions_%s.npz" % variant))
    o = Oracle(abi.default_options(**rc.VARIANTS[variant]))
    failures = []
    for fn, (name, kinds, gen) in sorted(rc.FUNCTIONS.items()):
        rec, want = fx["in_%d" % fn], fx["out_%d" % fn]
        msgs = compare_function(fn, kinds, want, o.function(fn, rec, want.shape[1]))
        failures += ["%s (%d): %s" % (name, fn, m) for m in msgs]
    assert not failures, "\n".join(failures)


# ---- scene preparation shaders --------------------------------------------------------------------------------
@live
def test_pre_transform_and_tri_light_extraction_shaders():
    """pre_transform.comp + extract_tri_lights.comp of the reference against the oracle's tri-light table (positions,
    RGB9E5 emission, half uvs, ids): byte-identical."""
    for scene in (testscenes.cornell(), testscenes.courtyard(emissive=True)):
        o = Oracle(abi.default_options())
        o.set_scene(scene)
        rp = ref_py.RefPathTracer("v0", scene, testscenes.cornell_cameras(1.0), o)
        assert rp.tri_light_count > 0
        assert rp.tri_lights[:rp.tri_light_count].tobytes() == o.tri_lights().tobytes()


# ---- image level ----------------------------------------------------------------------------------------------
@live
@pytest.mark.parametrize("name", sorted(IMAGE_CASES))
def test_frames_against_the_reference_pipeline(name):
    variant, scene, cams, size, opt = image_case(name)
    frames = 2
    o, want, so = oracle_frames(scene, cams, size, opt, frames)
    rp = ref_py.RefPathTracer(variant, scene, cams, o)
    col, dif, refl, cnt = rp.render(size, opt, frames=frames)
    close, mean_rel = compare_images(col, want, name)
    min_close, max_mean = IMAGE_LIMITS.get(name, DEFAULT_LIMITS)
    assert close >= min_close, "%s: only %.2f %% of the pixels within 1e-4" % (name, 100 * close)
    assert mean_rel <= max_mean, "%s: image means differ by %.2e" % (name, mean_rel)
    assert abs(cnt["closest_rays"] - so["closest_rays"]) <= 0.002 * so["closest_rays"]
    assert abs(cnt["shadow_rays"] - so["shadow_rays"]) <= 0.002 * so["shadow_rays"] + 2


@pytest.mark.parametrize("name", sorted(IMAGE_CASES))
def test_frames_against_committed_reference_images(name):
    variant, scene, cams, size, opt = image_case(name)
    fx = np.load(os.path.join(GOLDEN, "ref_glsl_images.npz"))
    o, want, so = oracle_frames(scene, cams, size, opt, 2)
    close, mean_rel = compare_images(fx[name], want, name)
    min_close, max_mean = IMAGE_LIMITS.get(name, DEFAULT_LIMITS)
    assert close >= min_close and mean_rel <= max_mean, (name, close, mean_rel)


@live
def test_directly_visible_emitter_shows_twice_its_emission_in_the_reference_code():
    """The one integrator golden upstream (validate_path-tracer.exr) shows 1x the emission on a directly visible
    emissive non-metal; the shader text at this commit adds L0 to `reflection` through primary_lobes = (0,0,0,1)
    (path_tracer.glsl:383,428) AND returns it as first_hit_material.emission (:434), summed at path_tracer.rgen:112.
    Executing that text settles it: 2x, as the oracle (and the CUDA path) compute."""
    from tauray_b200.scene import Scene, camera_pairs, look_at_transform, make_material, perspective_camera
    s = Scene()
    s.add_instance(*testscenes.quad((-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0), 1), make_material(albedo=(0.3, 0.3, 0.3, 1), emission=(1.0, 0.5, 0.25)))
    cams = camera_pairs([perspective_camera(look_at_transform((0, 0, 3), (0, 0, 0)), 30.0, 1.0, 0.1, 100.0)])
    opt = abi.default_options(**rc.VARIANTS["v1"])
    size = (8, 8)
    o, want, _ = oracle_frames(s, cams, size, opt, 1)
    col, _, _, _ = ref_py.RefPathTracer("v1", s, cams, o).render(size, opt)
    centre = col[0, 4, 4, :3]
    assert np.allclose(centre, (2.0, 1.0, 0.5), rtol=1e-6), centre
    assert np.allclose(want[0, 4, 4, :3], centre, rtol=1e-6)


@live
def test_test_glb_through_the_reference_pipeline(glb_scene):
    """The reference's own test asset (textures, alpha-tested logo through rt_common.rahit, transmission, spot light,
    sphere-light intersection shader) at the CLI defaults of its golden render: reference pipeline vs oracle."""
    from tauray_b200.scene import camera_pairs
    size = (96, 96)
    cams = camera_pairs([glb_scene.camera_data(0, aspect=1.0)])
    opt = abi.default_options(**dict(rc.VARIANTS["v1"], samples_per_pixel=2))
    o, want, so = oracle_frames(glb_scene.scene, cams, size, opt, 1)
    col, _, _, cnt = ref_py.RefPathTracer("v1", glb_scene.scene, cams, o).render(size, opt, frames=1)
    close, mean_rel = compare_images(col, want, "test.glb")
    assert close >= 0.98 and mean_rel <= 5e-3, (close, mean_rel)
    assert abs(cnt["closest_rays"] - so["closest_rays"]) <= 0.003 * so["closest_rays"]
