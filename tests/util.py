"""Shared helpers of the test-suite: scene construction, parameter sets and comparison rules."""
import importlib
import os

import numpy as np

abi = importlib.import_module("cosig-raytracing_b200.abi")
scene_mod = importlib.import_module("cosig-raytracing_b200.scene")
synth = importlib.import_module("cosig-raytracing_b200.synth")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REFERENCE_SCENES = "/root/reference/Assets/Resources/Scenes"  # build container only; never read by -m gpu tests


def params(width=320, height=240, depth=3, aa=1, **kw):
    p = abi.default_params()
    p.has_resolution, p.width, p.height = 1, width, height
    p.max_depth = depth
    p.aa_samples = aa
    for k, v in kw.items():
        if isinstance(v, (tuple, list)):
            getattr(p, k)[:] = list(v)
        else:
            setattr(p, k, v)
    return p


def oracle_scene(oracle, obj):
    """OracleScene of an ObjectData; returns (scene, holder) — keep `holder` alive while building."""
    packed = scene_mod.pack_scene(obj)
    return oracle.OracleScene.from_desc(packed.desc), packed


def rgb_agreement(a, b):
    """(fraction of pixels whose RGB channels all differ by <= 1, fraction identical, max abs diff)."""
    d = np.abs(a[..., :3].astype(np.int32) - b[..., :3].astype(np.int32)).max(axis=-1)
    return float((d <= 1).mean()), float((d == 0).mean()), int(d.max())


def assert_rgb_parity(a, b, what=""):
    """north_star bar: RGB within 1/255 per channel on >= 99.9 % of pixels."""
    within, same, worst = rgb_agreement(a, b)
    assert a.shape == b.shape, what
    assert within >= 0.999, f"{what}: only {within * 100:.4f}% of pixels within 1/255 (identical {same * 100:.4f}%, worst {worst})"
    assert (a[..., 3] == 255).all(), what
    return within, same, worst


def tiny_scene(n_tris=1):
    """A few triangles facing the sample camera, one light."""
    s = scene_mod.ObjectData()
    synth._sample_camera_and_light(s)
    s.Image = scene_mod.ImageSettings(64, 48, (0.1, 0.2, 0.3))
    s.Materials = [scene_mod.MaterialDescription((0.9, 0.2, 0.2), 0.1, 0.7, 0.3, 0.0, 1.0)]
    rng = np.random.RandomState(7)
    v = (rng.rand(n_tris, 3, 3).astype(np.float32) - 0.5) * 30.0
    v[:, :, 2] *= 0.2
    s.TriangleMeshes.append(scene_mod.TrianglesMesh(0, materials=np.zeros(n_tris, np.int32), vertices=v))
    return s
