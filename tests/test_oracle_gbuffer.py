"""Aux G-buffer layers of write_all_outputs (shader/path_tracer.glsl:547-574) in the oracle: cross-checked against
the feature (AOV) tracer -- which is pinned by the reference's own goldens -- and against the colour layer."""
import numpy as np

from tauray_b200 import abi, testscenes
from oracle.oracle_py import Oracle
from helpers import full_dist

SIZE = (48, 32)
ALL = [abi.GBUFFER_DIFFUSE, abi.GBUFFER_REFLECTION, abi.GBUFFER_ALBEDO, abi.GBUFFER_MATERIAL, abi.GBUFFER_NORMAL,
       abi.GBUFFER_POS, abi.GBUFFER_SCREEN_MOTION, abi.GBUFFER_INSTANCE_ID]


def _setup(opt, scene=None):
    scene = scene or testscenes.cornell(alpha_card=False, sphere_light=False)
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    o = Oracle(opt)
    o.set_scene(scene)
    o.set_cameras(cams)
    o.set_distribution(full_dist(SIZE))
    return o, scene, cams


def _octahedral_unpack(q):
    p = q.astype(np.float64) / 32767.0
    n = np.stack([p[..., 0], p[..., 1], 1 - np.abs(p[..., 0]) - np.abs(p[..., 1])], axis=-1)
    neg = n[..., 2] < 0
    x = (1 - np.abs(n[..., 1])) * np.where(n[..., 0] >= 0, 1, -1)
    y = (1 - np.abs(n[..., 0])) * np.where(n[..., 1] >= 0, 1, -1)
    n[..., 0] = np.where(neg, x, n[..., 0])
    n[..., 1] = np.where(neg, y, n[..., 1])
    return n / np.linalg.norm(n, axis=-1, keepdims=True)


def test_first_sample_layers_agree_with_the_feature_tracer():
    opt = abi.default_options(max_ray_depth=3, film=abi.FILM_POINT)
    o, scene, cams = _setup(opt)
    o.set_gbuffer_layers(ALL)
    o.render()
    inst = o.read_gbuffer_layer(abi.GBUFFER_INSTANCE_ID)[0, :, 0]
    pos = o.read_gbuffer_layer(abi.GBUFFER_POS)[0]
    nrm = _octahedral_unpack(o.read_gbuffer_layer(abi.GBUFFER_NORMAL)[0])
    alb = o.read_gbuffer_layer(abi.GBUFFER_ALBEDO)[0].astype(np.float32)
    mat = o.read_gbuffer_layer(abi.GBUFFER_MATERIAL)[0]
    feat = {}
    for f in (abi.FEATURE_INSTANCE_ID, abi.FEATURE_WORLD_POS, abi.FEATURE_WORLD_NORMAL, abi.FEATURE_ALBEDO):
        o.render_feature(f, (-1.0, -1.0, -1.0, 0.0))
        feat[f] = o.read_color()[0].reshape(-1, 4).copy()
    hit = feat[abi.FEATURE_INSTANCE_ID][:, 3] > 0
    assert hit.mean() > 0.9
    assert (inst[hit] == feat[abi.FEATURE_INSTANCE_ID][hit, 0].astype(np.int32)).all()
    assert np.allclose(pos[hit, :3], feat[abi.FEATURE_WORLD_POS][hit, :3], atol=2e-5)
    # the path tracer flips normals of back faces towards the viewer; the feature tracer does too (rt.glsl:93-99)
    assert (np.abs(nrm[hit] - feat[abi.FEATURE_WORLD_NORMAL][hit, :3]).max(axis=-1) < 2e-3).mean() > 0.999
    assert np.allclose(alb[hit, :3], feat[abi.FEATURE_ALBEDO][hit, :3], rtol=2e-3, atol=1e-4)
    # material layer: metallic / roughness of the scene's materials, ior/4
    assert set(np.unique(mat[hit, 0])) <= {0, 255}
    assert mat[hit, 2].min() > 0


def test_light_layers_recompose_the_colour_layer():
    """color = emission + modulate_color(albedo, diffuse, reflection) (path_tracer.rgen:112, material.glsl:57-65):
    checked for dielectric, untextured first hits, where modulate_color = albedo * diffuse + reflection."""
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=4, film=abi.FILM_POINT, local_sampler=abi.SAMPLER_SOBOL_OWEN,
                              sampling_weight_emissive_triangles=1.0)
    scene = testscenes.cornell(alpha_card=False, sphere_light=False, spotlight=False, transmissive=False)
    o, scene, cams = _setup(opt, scene)
    o.set_gbuffer_layers(ALL)
    o.render()
    col = o.read_color()[0].reshape(-1, 4)
    dif = o.read_gbuffer_layer(abi.GBUFFER_DIFFUSE)[0].astype(np.float64)
    ref = o.read_gbuffer_layer(abi.GBUFFER_REFLECTION)[0].astype(np.float64)
    alb = o.read_gbuffer_layer(abi.GBUFFER_ALBEDO)[0].astype(np.float64)
    mat = o.read_gbuffer_layer(abi.GBUFFER_MATERIAL)[0]
    inst = o.read_gbuffer_layer(abi.GBUFFER_INSTANCE_ID)[0, :, 0]
    # first four instances: the walls (white, red, blue) -- dielectric, non-emissive
    sel = (inst >= 0) & (inst <= 2) & (mat[:, 0] == 0)
    assert sel.sum() > 200
    want = alb[sel, :3] * dif[sel, :3] + ref[sel, :3]
    got = col[sel, :3].astype(np.float64)
    err = np.abs(want - got) / np.maximum(np.abs(got), 1e-2)
    assert np.median(err) < 3e-3 and np.percentile(err, 95) < 2e-2, (float(np.median(err)), float(np.percentile(err, 95)))
    # alpha of the light layers: reciprocal distance to the second path vertex, only set when bounce 1 shades a surface
    assert (dif[sel, 3] >= 0).all() and (dif[sel, 3] > 0).mean() > 0.5
    assert np.array_equal(dif[:, 3], ref[:, 3])


def test_first_sample_layers_are_written_once_and_light_layers_accumulate():
    opt = abi.default_options(max_ray_depth=3, film=abi.FILM_BOX)   # jittered camera rays: later samples hit elsewhere
    o, scene, cams = _setup(opt)
    o.set_gbuffer_layers(ALL)
    o.render()
    pos1 = o.read_gbuffer_layer(abi.GBUFFER_POS).copy()
    dif1 = o.read_gbuffer_layer(abi.GBUFFER_DIFFUSE).copy()
    o.render()
    assert np.array_equal(pos1, o.read_gbuffer_layer(abi.GBUFFER_POS))
    assert not np.array_equal(dif1, o.read_gbuffer_layer(abi.GBUFFER_DIFFUSE))
    o.reset_accumulation(True)
    o.render()
    assert np.array_equal(pos1, o.read_gbuffer_layer(abi.GBUFFER_POS))
    assert np.array_equal(dif1, o.read_gbuffer_layer(abi.GBUFFER_DIFFUSE))


def test_screen_motion_of_a_static_scene_is_the_pixel_itself():
    opt = abi.default_options(max_ray_depth=2, film=abi.FILM_POINT)
    o, scene, cams = _setup(opt)
    o.set_gbuffer_layers([abi.GBUFFER_SCREEN_MOTION, abi.GBUFFER_INSTANCE_ID])
    o.render()
    mo = o.read_gbuffer_layer(abi.GBUFFER_SCREEN_MOTION)[0].reshape(SIZE[1], SIZE[0], 4)
    inst = o.read_gbuffer_layer(abi.GBUFFER_INSTANCE_ID)[0, :, 0].reshape(SIZE[1], SIZE[0])
    ys, xs = np.mgrid[0:SIZE[1], 0:SIZE[0]]
    u = (xs + 0.5) / SIZE[0]
    v = 1.0 - (ys + 0.5) / SIZE[1]
    hit = inst >= 0
    assert np.allclose(mo[..., 0][hit], u[hit], atol=1e-4)
    assert np.allclose(mo[..., 1][hit], v[hit], atol=1e-4)
    assert (mo[..., 2][hit] > 0).all()
