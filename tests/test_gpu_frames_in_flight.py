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


# ---- frames in flight on the device: tr_pt_render_begin / tr_pt_render_end -------------------------------------------
def _reference_frames(opt, scene, frames, accumulate):
    out = []
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_cameras(_cams(0))
        pt.set_distribution(full_dist(SIZE))
        for k in range(frames):
            pt.set_cameras(_cams(0 if accumulate else k))
            if not accumulate:
                pt.reset_accumulation()
            pt.render()
            pt.sync()
            out.append(pt.read_color().copy())
    return out


@pytest.mark.parametrize("accumulate", [False, True])
def test_begin_end_pipeline_equals_frame_by_frame(accumulate):
    """begin(0); { begin(k + 1); end(k) }: two frames' traversal work in the queue at any time. Different cameras per frame
    (stream-ordered uploads, each frame keeps its own copy of the table) or one accumulating running mean -- every frame
    must be the picture the frame-by-frame path gives, and the ray counters must be the frame's own."""
    import torch
    opt = abi.default_options(max_ray_depth=4, samples_per_pixel=4, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    scene = testscenes.cornell()
    frames = 6
    want = _reference_frames(opt, scene, frames, accumulate)
    stream = torch.cuda.Stream()
    copy_stream = torch.cuda.Stream()
    pinned = [torch.from_numpy(np.frombuffer(_cams(0 if accumulate else k).tobytes(), dtype=np.uint8).copy()).pin_memory() for k in range(frames)]
    hosts = [torch.zeros(want[0].shape, dtype=torch.float32).pin_memory() for _ in range(frames)]
    targets = [torch.zeros(want[0].shape, dtype=torch.float32, device="cuda") for _ in range(2)]
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_cameras(_cams(0))
        pt.set_distribution(full_dist(SIZE))
        with pytest.raises(TaurayB200Error):
            pt.render_end(stream.cuda_stream)                    # nothing begun

        def begin(k):
            pt.set_cameras_async(pinned[k].numpy().view(abi.CAMERA_PAIR), stream.cuda_stream)
            if not accumulate:
                pt.set_color_buffer(targets[k & 1].data_ptr(), targets[k & 1].numel() * 4)
                pt.reset_accumulation()
            pt.render_begin(stream.cuda_stream)

        begin(0)
        rays = []
        for k in range(frames):
            if k + 1 < frames:
                begin(k + 1)
            pt.render_end(stream.cuda_stream)
            pt.read_color_async(hosts[k].data_ptr(), hosts[k].numel() * 4, copy_stream.cuda_stream)
            if k == 2:
                with pytest.raises(TaurayB200Error):
                    pt.render()                                  # the frame-by-frame call while a frame is begun
                rays.append(pt.stats()["closest_rays"])
        with pytest.raises(TaurayB200Error):
            pt.render_end(stream.cuda_stream)
        copy_stream.synchronize()
        for k in range(frames):
            assert bits_equal(hosts[k].numpy(), want[k]), "frame %d" % k
        assert rays[0] > 0
        # and the frame-by-frame call works again afterwards
        pt.set_color_buffer(None, 0)
        pt.set_cameras(_cams(1))
        pt.reset_accumulation()
        pt.render()
        pt.sync()
        assert bits_equal(pt.read_color(), _reference_frames(opt, scene, 2, False)[1])


def test_begin_end_falls_back_for_frames_it_cannot_split():
    """Aux G-buffer layers (or several wavefront batches, the megakernel, ...): tr_pt_render_end renders the whole frame."""
    opt = abi.default_options(max_ray_depth=3, samples_per_pixel=2, local_sampler=abi.SAMPLER_SOBOL_OWEN)
    scene = testscenes.cornell()
    with PathTracer(opt) as pt:
        pt.set_scene(scene)
        pt.set_cameras(_cams(0))
        pt.set_distribution(full_dist(SIZE))
        pt.set_gbuffer_layers([abi.GBUFFER_ALBEDO])
        pt.render()
        pt.sync()
        want = pt.read_color().copy()
        pt.reset_accumulation()
        pt.render_begin()
        pt.render_begin()                                        # such frames queue up and are rendered one by one
        with pytest.raises(TaurayB200Error):
            pt.render_begin()                                    # ... but never more than two
        pt.render_end()
        pt.sync()
        assert bits_equal(pt.read_color(), want)
        # the second frame took its place in the accumulation when it was begun: samples 2..3 on top of the first
        pt.render_end()
        pt.sync()
        with PathTracer(opt) as ref:
            ref.set_scene(scene); ref.set_cameras(_cams(0)); ref.set_distribution(full_dist(SIZE))
            ref.set_gbuffer_layers([abi.GBUFFER_ALBEDO])
            ref.render(); ref.render(); ref.sync()
            assert bits_equal(pt.read_color(), ref.read_color())
        with pytest.raises(TaurayB200Error):
            pt.render_end()
