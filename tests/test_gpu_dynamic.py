"""Animated scenes: tr_pt_update_instances (the per-frame part of scene_stage::update,
src/scene_stage.cc:1572-1675) must give exactly what a fresh tr_pt_set_scene with the new instance
records gives, and that must be the oracle's image of the moved scene."""
import numpy as np
import pytest

from tauray_b200 import abi, testscenes
from tauray_b200.library import PathTracer, TaurayB200Error
from oracle.oracle_py import Oracle

from helpers import bits_equal, full_dist

pytestmark = pytest.mark.gpu

SIZE = (72, 48)


def _moved(scene, step):
    """The animation: blocks slide and turn, the emissive panel (a tri-light instance) drops a little."""
    prev = {k: scene.transforms[k].copy() for k in range(len(scene.transforms))}
    for k in (6, 7):
        t = scene.transforms[k].copy()
        a = 0.35 * step * (1 if k == 6 else -1)
        rot = np.eye(4, dtype=np.float32)
        rot[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        t = (t @ rot).astype(np.float32)
        t[:3, 3] += np.float32(step) * np.array([0.15, 0.0, 0.1 * (k - 6)], dtype=np.float32)
        scene.transforms[k] = t
    t = scene.transforms[5].copy()
    t[1, 3] -= np.float32(0.2 * step)
    scene.transforms[5] = t
    scene.prev_transforms = prev


def _render(t, cams, dist, frames=1):
    t.set_cameras(cams)
    t.set_distribution(dist)
    for _ in range(frames):
        t.render()
    if hasattr(t, "sync"):
        t.sync()
    return t.read_color()


@pytest.mark.parametrize("mode", [abi.KERNEL_WAVEFRONT, abi.KERNEL_MEGAKERNEL])
def test_update_instances_equals_a_fresh_scene_and_the_oracle(mode):
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    opt.kernel_mode = mode
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    dist = full_dist(SIZE)
    scene = testscenes.cornell()
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_gbuffer_layers([abi.GBUFFER_SCREEN_MOTION])
        first = _render(pt, cams, dist)
        for step in (1, 2):
            _moved(scene, step)
            pt.update_instances(scene.instance_table())
            pt.reset_accumulation(False)
            got = _render(pt, cams, dist)
            motion = pt.read_gbuffer_layer(abi.GBUFFER_SCREEN_MOTION)
            lights = pt.tri_lights()
            with PathTracer(opt) as fresh:
                fresh.set_scene(scene)
                fresh.set_gbuffer_layers([abi.GBUFFER_SCREEN_MOTION])
                # same frame counter as `pt` (one frame rendered per step before this one)
                for _ in range(step):
                    fresh.set_cameras(cams); fresh.set_distribution(dist); fresh.render()
                fresh.reset_accumulation(False)
                want = _render(fresh, cams, dist)
                assert bits_equal(fresh.read_gbuffer_layer(abi.GBUFFER_SCREEN_MOTION), motion)
                assert fresh.tri_lights().tobytes() == lights.tobytes()
            assert bits_equal(got, want), "step %d: update_instances differs from a fresh set_scene" % step
            assert not bits_equal(got, first)
    # the oracle's image of the final pose, from a clean frame counter
    ref = Oracle(opt)
    ref.set_scene(scene)
    want = _render(ref, cams, dist)
    with PathTracer(opt) as pt:
        base = testscenes.cornell()
        pt.set_scene(base)
        _moved(base, 1)
        _moved(base, 2)
        pt.update_instances(base.instance_table())
        got = _render(pt, cams, dist)
    assert bits_equal(got, want)


def test_update_instances_rejects_a_different_layout():
    opt = abi.default_options(max_ray_depth=2)
    scene = testscenes.cornell()
    with PathTracer(opt) as pt:
        with pytest.raises(TaurayB200Error):
            pt.update_instances(scene.instance_table())          # no scene yet
        pt.set_scene(scene)
        with pytest.raises(TaurayB200Error):
            pt.update_instances(scene.instance_table()[:-1])     # instance count differs
        table = scene.instance_table()
        table[5]["light_base_id"] = -1                           # the emissive panel stops being a tri-light instance
        with pytest.raises(TaurayB200Error):
            pt.update_instances(table)
        pt.update_instances(scene.instance_table())              # unchanged records are fine


# ---- TR_UPDATE_REFIT: the scene's tree keeps its topology, only the boxes follow the motion ---------------------------
@pytest.mark.parametrize("mode", [abi.KERNEL_WAVEFRONT, abi.KERNEL_MEGAKERNEL])
def test_refit_updates_equal_a_fresh_scene(mode):
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    opt.kernel_mode = mode
    cams = testscenes.cornell_cameras(SIZE[0] / SIZE[1])
    dist = full_dist(SIZE)
    scene = testscenes.cornell()
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_update_mode(abi.UPDATE_REFIT)
        nodes_before = pt.stats()["bvh_node_count"]
        first = _render(pt, cams, dist)
        for step in (1, 2, 3):
            _moved(scene, step)
            pt.update_instances(scene.instance_table())
            assert pt.stats()["bvh_node_count"] == nodes_before          # same tree
            pt.reset_accumulation(False)
            got = _render(pt, cams, dist)
            lights = pt.tri_lights()
            with PathTracer(opt) as fresh:
                fresh.set_scene(scene)
                for _ in range(step):
                    fresh.set_cameras(cams); fresh.set_distribution(dist); fresh.render()
                fresh.reset_accumulation(False)
                want = _render(fresh, cams, dist)
                assert fresh.tri_lights().tobytes() == lights.tobytes()
            assert bits_equal(got, want), "step %d: the refitted tree renders another image than a fresh scene" % step
            assert not bits_equal(got, first)
        with pytest.raises(TaurayB200Error):
            pt.set_update_mode(7)


def test_an_identity_refit_reproduces_the_built_tree():
    """Refitting to the unchanged instance records must give back the boxes the build wrote: the same image AND exactly the
    same number of node visits and triangle tests (a looser box would still render the right image). Counted with the
    megakernel organisation -- one ray per thread, a fixed traversal order -- because the visit counts of the persistent
    wavefront kernel depend a little on which rays happen to share a warp (postponed triangle groups)."""
    from tauray_b200 import synthetic
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=1, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    opt.kernel_mode = abi.KERNEL_MEGAKERNEL
    size = (160, 90)
    scene = synthetic.atrium(20000)
    cams = synthetic.atrium_cameras(size[0] / size[1])
    dist = full_dist(size)
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_instrumentation(abi.INSTRUMENT_VISITS)
        want = _render(pt, cams, dist)
        st0 = pt.stats()
        pt.set_update_mode(abi.UPDATE_REFIT)
        pt.update_instances(scene.instance_table())
        pt.reset_accumulation(True)
        got = _render(pt, cams, dist)
        st1 = pt.stats()
    assert bits_equal(got, want)
    for k in ("closest_rays", "shadow_rays", "closest_node_visits", "closest_triangle_tests", "shadow_node_visits", "shadow_triangle_tests"):
        assert st0[k] == st1[k] and st0[k] > 0, (k, st0[k], st1[k])
