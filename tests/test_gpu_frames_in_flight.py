"""tr_pt_read_color_async: the staging copy of frame k runs on a copy stream beside frame k + 1 (two device targets, two
pinned host images -- the reference's MAX_FRAMES_IN_FLIGHT = 2 headless loop, src/headless.cc:125-171). Every host image
must equal what the blocking path gives for the same frame, with different content per frame."""
import numpy as np
import pytest

from tauray_b200 import abi, testscenes
from tauray_b200.scene import camera_pairs, look_at_transform, perspective_camera
from tauray_b200.library import PathTracer, TaurayB200Error

from helpers import bits_equal, full_dist

pytestmark = pytest.mark.gpu

SIZE = (160, 96)


def _cams(k):
    # a different view per frame
    cam = perspective_camera(look_at_transform((0.15 * k, 0, 5.5), (0, -0.2, 0)), 45.0, SIZE[0] / SIZE[1], 0.1, 100.0)
    return camera_pairs([cam])


def test_async_staging_copies_equal_the_blocking_read():
    import torch
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    scene = testscenes.cornell()
    frames = 6
    want = []
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_cameras(_cams(0))
        pt.set_distribution(full_dist(SIZE))
        for k in range(frames):
            pt.set_cameras(_cams(k))
            pt.reset_accumulation()
            pt.render()
            pt.sync()
            want.append(pt.read_color().copy())
    assert not bits_equal(want[0], want[1])
    render_stream = torch.cuda.Stream()
    copy_stream = torch.cuda.Stream()
    shape = want[0].shape
    targets = [torch.zeros(shape, dtype=torch.float32, device="cuda") for _ in range(2)]
    hosts = [torch.zeros(shape, dtype=torch.float32).pin_memory() for _ in range(frames)]
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_cameras(_cams(0))
        pt.set_distribution(full_dist(SIZE))
        with pytest.raises(TaurayB200Error):
            pt.read_color_async(hosts[0].data_ptr(), hosts[0].numel() * 4, copy_stream.cuda_stream)   # nothing rendered yet
        # the camera tables go up in stream order too (tr_pt_set_cameras_async): one pinned table per frame, untouched
        # until its copy has run; the host never waits inside this loop
        pinned = []
        for k in range(frames):
            t = torch.from_numpy(np.frombuffer(_cams(k).tobytes(), dtype=np.uint8).copy()).pin_memory()
            pinned.append(t)
        for k in range(frames):
            pt.set_cameras_async(pinned[k].numpy().view(abi.CAMERA_PAIR), render_stream.cuda_stream)
            pt.set_color_buffer(targets[k & 1].data_ptr(), targets[k & 1].numel() * 4)
            pt.reset_accumulation()
            pt.render(render_stream.cuda_stream)
            pt.read_color_async(hosts[k].data_ptr(), hosts[k].numel() * 4, copy_stream.cuda_stream)
        with pytest.raises(TaurayB200Error):
            pt.read_color_async(hosts[0].data_ptr(), 16, copy_stream.cuda_stream)                     # size mismatch
        copy_stream.synchronize()
        for k in range(frames):
            assert bits_equal(hosts[k].numpy(), want[k]), "frame %d" % k
        pt.set_color_buffer(None, 0)


def test_a_render_into_the_same_target_waits_for_the_pending_copy():
    import torch
    opt = abi.default_options(max_ray_depth=2, samples_per_pixel=1, local_sampler=abi.SAMPLER_UNIFORM_RANDOM)
    scene = testscenes.cornell()
    render_stream = torch.cuda.Stream()
    copy_stream = torch.cuda.Stream()
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_distribution(full_dist(SIZE))
        want = []
        for k in range(3):
            pt.set_cameras(_cams(k))
            pt.reset_accumulation()
            pt.render()
            pt.sync()
            want.append(pt.read_color().copy())
        hosts = [torch.zeros(want[0].shape, dtype=torch.float32).pin_memory() for _ in range(3)]
        # ONE target: frame k + 1 must not overwrite it before the copy of frame k has read it
        for k in range(3):
            pt.set_cameras(_cams(k))
            pt.reset_accumulation()
            pt.render(render_stream.cuda_stream)
            pt.read_color_async(hosts[k].data_ptr(), hosts[k].numel() * 4, copy_stream.cuda_stream)
        copy_stream.synchronize()
        for k in range(3):
            assert bits_equal(hosts[k].numpy(), want[k]), "frame %d" % k
