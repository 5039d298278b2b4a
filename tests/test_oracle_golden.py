"""Pin the oracle against every golden image the reference's own tests hold for this path
(test/CMakeLists.txt:29-37, test/validate_render.py:27-43): six feature renders of test.glb at 512x512
(tolerance 1 in ImageMagick 'compare -metric mse' units) and the path-tracer render (tolerance 10000)."""
import os

import cv2
import numpy as np
import pytest

from tauray_b200 import abi, distribution
from tauray_b200.scene import camera_pairs
from oracle.oracle_py import Oracle

FEATURES = {
    "albedo": abi.FEATURE_ALBEDO, "world-normal": abi.FEATURE_WORLD_NORMAL, "view-normal": abi.FEATURE_VIEW_NORMAL,
    "world-pos": abi.FEATURE_WORLD_POS, "view-pos": abi.FEATURE_VIEW_POS, "distance": abi.FEATURE_DISTANCE,
}


def read_exr(path):
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert img is not None, path
    return img[..., ::-1].astype(np.float32)


def imagemagick_mse(a, b):
    """`compare -metric mse` prints normalised-MSE x QuantumRange (65535) first; that is the number
    test/validate_render.py compares with the tolerance. Values are clamped to [0, 1] like a Q16 image."""
    a = np.clip(a, 0, 1).astype(np.float64)
    b = np.clip(b, 0, 1).astype(np.float64)
    return float(((a - b) ** 2).mean() * 65535.0)


@pytest.fixture(scope="module")
def glb_oracle(glb_scene, oracle_lib):
    o = Oracle(abi.default_options())           # CLI defaults: depth 8, 1 spp, uniform-random
    o.set_scene(glb_scene.scene)
    o.set_cameras(camera_pairs([glb_scene.camera_data(0, aspect=1.0)]))
    o.set_distribution(distribution.make_distribution((512, 512), abi.DISTRIBUTION_DUPLICATE))
    return o


def test_test_glb_matches_the_survey(glb_scene):
    s = glb_scene.scene
    assert s.triangle_count == 81364
    assert glb_scene.instance_names == ["Cube"] * 4 + ["Teapot", "Suzanne", "Torus", "Plane"]
    assert s.opaque == [1, 1, 1, 1, 1, 0, 1, 0]          # Suzanne transmissive, Plane alpha-blended
    assert s.directional_lights.shape[0] == 1 and s.point_lights.shape[0] == 1
    assert s.textures[0].shape == (2048, 2048, 4)
    inst = s.instance_table()
    assert inst["light_base_id"][6] == 0 and (inst["light_base_id"][:6] == -1).all()


@pytest.mark.parametrize("name", sorted(FEATURES))
def test_feature_goldens(glb_oracle, golden_dir, name):
    ref = read_exr(os.path.join(golden_dir, "references", "validate_%s.exr" % name))
    glb_oracle.render_feature(FEATURES[name], (0, 0, 0, 0))
    img = glb_oracle.read_color()[0, :, :, :3]
    # the reference's criterion (tolerance 1) ...
    assert imagemagick_mse(img, ref) < 1.0
    # ... and much tighter: the goldens are half floats, so agree to half rounding everywhere but
    # on a handful of silhouette pixels (bilinear alpha of the logo plane differs in the last texture bits)
    d = np.abs(img - ref)
    tol = np.abs(ref) * 2.0 ** -10 + 1e-3
    bad = (d > tol).any(axis=-1)
    assert bad.sum() <= 4, (name, int(bad.sum()), float(d.max()))


def test_feature_known_values(glb_oracle):
    """distance at the image centre = camera z 6.828 + back wall 2."""
    glb_oracle.render_feature(abi.FEATURE_DISTANCE, (0, 0, 0, 0))
    d = glb_oracle.read_color()[0, 256, 256, 0]
    assert abs(d - 8.828) < 2e-3
    glb_oracle.render_feature(abi.FEATURE_WORLD_POS, (0, 0, 0, 0))
    p = glb_oracle.read_color()[0, 256, 256, :3]
    assert np.allclose(p, (0.0071, -0.0071, -2.0), atol=2e-4)


def filmic(c):
    """tonemap_filmic.comp:15-23 followed by the gamma stage: net = the Hejl curve."""
    c = np.maximum(0.0, np.clip(c, 0, 1000) - 0.004)
    return (c * (6.2 * c + 0.5)) / (c * (6.2 * c + 1.7) + 0.06)


def test_path_tracer_golden_statistically(glb_scene, golden_dir, oracle_lib):
    """validate_path-tracer.exr pins sampler/BSDF/NEE/accumulate only statistically (tolerance 10000)."""
    ref = read_exr(os.path.join(golden_dir, "references", "validate_path-tracer.exr"))
    o = Oracle(abi.default_options(samples_per_pixel=8))
    o.set_scene(glb_scene.scene)
    o.set_cameras(camera_pairs([glb_scene.camera_data(0, aspect=1.0)]))
    o.set_distribution(distribution.make_distribution((512, 512), abi.DISTRIBUTION_DUPLICATE))
    o.render()
    img = o.read_color()[0, :, :, :3]
    assert not np.isnan(img).any()
    tm = filmic(img)
    assert imagemagick_mse(tm, ref) < 10000.0          # the reference's own criterion
    assert imagemagick_mse(tm, ref) < 4000.0           # ours (8 spp noise dominates this number)
    # Compare in LINEAR space (averaging noisy samples commutes with blurring there, unlike after the tonemap):
    # invert the monotonic Hejl curve of the golden by table lookup.
    xs = np.linspace(0.0, 8.0, 80001)
    ref_lin = np.interp(np.clip(ref, 0, filmic(xs[-1])), filmic(xs), xs)
    ok = (ref < 0.97).all(axis=-1, keepdims=True)       # skip saturated pixels (lights, speculars)
    blur = lambda a: cv2.GaussianBlur(a.astype(np.float32), (0, 0), 8)  # noqa: E731
    a, b = blur(np.minimum(img, 8.0) * ok), blur(ref_lin * ok)
    assert np.abs(a - b).mean() / b.mean() < 0.12
    assert abs(a.mean() - b.mean()) / b.mean() < 0.08
    # KNOWN deviation from the golden, kept because the restatement follows the shader text: the directly seen
    # emissive torus is 2x the golden (path_tracer.glsl:383,428,434 + path_tracer.rgen:112 add L0 both as
    # first_hit emission and through primary_lobes = (0,0,0,1)); the golden predates that code (it is also far
    # less noisy than 1 spp). See DESIGN.md "bug compatibility".
    assert abs(img[150, 110, 0] - 2.0) < 0.05
