"""The C++ host mirror (tauray_b200/host/tauray_b200.hh: path_tracer_stage, rt_renderer, load_balancer with the
reference's names) drives the same C ABI: its image must equal the Python-driven one bit for bit, for one device
and for two stages stitched through peer images (TR_FAKE_DEVICES=2 == the reference's --fake-devices)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tauray_b200 import abi, testscenes
from tauray_b200.library import PathTracer

from helpers import bits_equal, full_dist, run

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tauray_b200", "host", "host_demo")


def dump_scene(path, scene, cams):
    inst = scene.instance_table()
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(scene.meshes)))
        for i, (v, idx) in enumerate(scene.meshes):
            f.write(inst[i].tobytes())
            f.write(struct.pack("<III", v.shape[0], idx.size, scene.opaque[i]))
            f.write(v.tobytes())
            f.write(idx.tobytes())
        f.write(struct.pack("<I", scene.directional_lights.shape[0]))
        f.write(scene.directional_lights.tobytes())
        f.write(struct.pack("<I", scene.point_lights.shape[0]))
        f.write(scene.point_lights.tobytes())
        f.write(struct.pack("<I", cams.shape[0]))
        f.write(cams.tobytes())


def test_host_mirror_builds_without_a_gpu():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tauray_b200", "host")], check=True)
    assert os.path.exists(DEMO)


@pytest.mark.gpu
@pytest.mark.parametrize("fake_devices", [1, 2])
def test_cpp_host_mirror_matches_python(tmp_path, fake_devices):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tauray_b200", "host")], check=True)
    size = (160, 90)
    scene = testscenes.cornell(alpha_card=False)
    cams = testscenes.cornell_cameras(size[0] / size[1])
    dump_scene(tmp_path / "scene.bin", scene, cams)
    env = dict(os.environ, TR_FAKE_DEVICES=str(fake_devices))
    r = subprocess.run([DEMO, str(tmp_path / "scene.bin"), str(tmp_path / "out.raw"), str(size[0]), str(size[1]), "2", "4", "1"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "out.raw", dtype=np.float32).reshape(1, size[1], size[0], 4)
    opt = abi.default_options(max_ray_depth=4, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    with PathTracer(opt) as pt:
        want = run(pt, scene, cams, full_dist(size))
    assert bits_equal(got, want)


@pytest.mark.gpu
def test_cpp_rt_renderer_rebalances_like_the_reference(tmp_path):
    """rt_renderer::set_device_workloads under accumulation through the C++ mirror (2 stages on this GPU): device 0
    is the display device and keeps accumulating in place, device 1 restarts and its strip is blended with 1/3."""
    import _rebalance_common as rc
    from tauray_b200 import distribution
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tauray_b200", "host")], check=True)
    size = (160, 90)
    scene = testscenes.cornell(alpha_card=False)
    cams = testscenes.cornell_cameras(size[0] / size[1])
    dump_scene(tmp_path / "scene.bin", scene, cams)
    env = dict(os.environ, TR_FAKE_DEVICES="2")
    r = subprocess.run([DEMO, str(tmp_path / "scene.bin"), str(tmp_path / "out.raw"), str(size[0]), str(size[1]), "2", "4", "1", "rebalance"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "out.raw", dtype=np.float32).reshape(1, size[1], size[0], 4)
    opt = abi.default_options(max_ray_depth=4, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_cameras(cams)
        _, a_pri, a_oth, _, _ = rc.expected_images(pt, full_dist(size), 2)
    dists = distribution.device_distributions(size, abi.DISTRIBUTION_SHUFFLED_STRIPS, [0.3, 0.7], primary_device=0)
    mask = rc.primary_mask(size, dists, 0)
    assert 0 < mask.sum() < mask.size
    assert rc.mismatches(got, rc.compose(mask, a_pri, a_oth)) <= 2


@pytest.mark.gpu
def test_cpp_direct_stage_matches_python(tmp_path):
    """tr_b200::direct_stage (the mirror of src/direct_stage.hh) == the Python-driven direct estimator."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tauray_b200", "host")], check=True)
    size = (96, 64)
    scene = testscenes.cornell(alpha_card=False)
    cams = testscenes.cornell_cameras(size[0] / size[1])
    dump_scene(tmp_path / "scene.bin", scene, cams)
    r = subprocess.run([DEMO, str(tmp_path / "scene.bin"), str(tmp_path / "out.raw"), str(size[0]), str(size[1]), "4", "4", "1", "direct"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "out.raw", dtype=np.float32).reshape(1, size[1], size[0], 4)
    opt = abi.default_options(estimator=abi.ESTIMATOR_DIRECT, max_ray_depth=4, samples_per_pixel=4, samples_per_pass=2,
                              local_sampler=abi.SAMPLER_SOBOL_OWEN, film=abi.FILM_BOX, film_radius=0.5)
    with PathTracer(opt) as pt:
        want = run(pt, scene, cams, full_dist(size))
    assert bits_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("fake_devices,demo_mode", [(1, "animate"), (2, "animate"), (1, "refit")])
def test_cpp_update_instances_matches_python(tmp_path, fake_devices, demo_mode):
    """rt_renderer::update_instances (scene_stage::update of an animated scene) through the C++ mirror; `refit` =
    rt_renderer::set_update_mode(TR_UPDATE_REFIT), which must render the same picture."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tauray_b200", "host")], check=True)
    size = (96, 64)
    scene = testscenes.cornell(alpha_card=False)
    cams = testscenes.cornell_cameras(size[0] / size[1])
    dump_scene(tmp_path / "scene.bin", scene, cams)
    env = dict(os.environ, TR_FAKE_DEVICES=str(fake_devices))
    r = subprocess.run([DEMO, str(tmp_path / "scene.bin"), str(tmp_path / "out.raw"), str(size[0]), str(size[1]), "2", "3", "1", demo_mode],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "out.raw", dtype=np.float32).reshape(1, size[1], size[0], 4)
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    table = scene.instance_table()
    table[6]["model_prev"] = table[6]["model"]
    flat = table[6]["model"].reshape(-1)            # column-major wire layout: floats 12..14 = translation
    flat[12] += np.float32(0.2)
    flat[14] += np.float32(0.1)
    with PathTracer(opt) as pt:
        pt.set_scene(scene); pt.set_cameras(cams); pt.set_distribution(full_dist(size))
        pt.render()
        pt.update_instances(table)
        pt.reset_accumulation(True)
        pt.render(); pt.sync()
        want = pt.read_color()
    assert bits_equal(got, want)


@pytest.mark.gpu
def test_cpp_rt_renderer_scanline_across_three_stages(tmp_path):
    """--distribution-strategy=scanline through the C++ mirror: three stages (TR_FAKE_DEVICES=3) render every third
    row; the display stage renders in place, the others store their rows into its image over peer memory."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tauray_b200", "host")], check=True)
    size = (96, 50)
    scene = testscenes.cornell(alpha_card=False)
    cams = testscenes.cornell_cameras(size[0] / size[1])
    dump_scene(tmp_path / "scene.bin", scene, cams)
    env = dict(os.environ, TR_FAKE_DEVICES="3")
    r = subprocess.run([DEMO, str(tmp_path / "scene.bin"), str(tmp_path / "out.raw"), str(size[0]), str(size[1]), "2", "3", "1", "scanline"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(tmp_path / "out.raw", dtype=np.float32).reshape(1, size[1], size[0], 4)
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    with PathTracer(opt) as pt:
        want = run(pt, scene, cams, full_dist(size))
    assert bits_equal(got, want)


@pytest.mark.gpu
def test_cpp_errors_surface_as_runtime_error(tmp_path):
    r = subprocess.run([DEMO, str(tmp_path / "missing.bin"), str(tmp_path / "o.raw"), "8", "8", "1", "1", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr
