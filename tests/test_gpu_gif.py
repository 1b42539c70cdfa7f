"""GIF sweep, GPU part: the device palette kernel and the fused render -> quantise -> readback -> LZW sweep
(through the C ABI) against the oracle's restatement of GifGenerator.cs.  Bar: byte-identical indices and files.
"""
import importlib

import numpy as np
import pytest

from util import abi, scene_mod, synth

pytestmark = pytest.mark.gpu

rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
gif = importlib.import_module("cosig-raytracing_b200.gif_generator")


@pytest.fixture(scope="module")
def rt(pkg):
    r = rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_REFERENCE)
    yield r
    r.close()


BASE = dict(ResolutionOverride=(160, 120), MaxDepth=3, CameraPositionOverride=(0.0, 0.0, 0.0), CameraRotationOverride=(-60.0, 0.0, 0.0))


@pytest.mark.parametrize("h,w", [(48, 64), (33, 37), (1, 1), (270, 482), (1080, 1920)])
def test_palette_kernel_matches_convert_to_indexed(rt, oracle, h, w):
    rng = np.random.RandomState(h * 7 + w)
    frame = rng.randint(0, 256, size=(h, w, 4)).astype(np.uint8)
    out = np.zeros((h, w), np.uint8)
    rt._check(abi.load().rtb_gif_index_frame(rt._ctx, frame.ctypes.data, w, h, out.ctypes.data))
    assert (out == oracle.gif_convert_to_indexed(frame)).all()


def test_palette_kernel_every_channel_value(rt, oracle):
    # all 256 values of each channel against the FP32 formula (int)(byte / 255f * 5.99f) of the restatement
    frame = np.zeros((3, 256, 4), np.uint8)
    for c in range(3):
        frame[c, :, c] = np.arange(256)
    out = np.zeros((3, 256), np.uint8)
    rt._check(abi.load().rtb_gif_index_frame(rt._ctx, frame.ctypes.data, 256, 3, out.ctypes.data))
    assert (out == oracle.gif_convert_to_indexed(frame)).all()
    assert set(np.unique(out[2])) == {0, 36, 72, 108, 144, 180}  # row 0 (red ramp) comes out last


def test_render_begin_indexed_equals_indexed_render(rt, oracle):
    obj = synth.sample_scene("test_scene_1")
    st = scene_mod.RenderSettings(**BASE)
    rgba = np.zeros((120, 160, 4), np.uint8)
    rt.RenderInto(obj, st, rgba)
    idx = np.zeros((120, 160), np.uint8)
    rt.RenderEnd(rt.RenderBeginIndexed(obj, st, idx))
    assert (idx == oracle.gif_convert_to_indexed(rgba)).all()
    assert len(np.unique(idx)) > 8
    small = np.zeros((10,), np.uint8)
    with pytest.raises(rt_mod.RtbError) as e:
        rt.RenderBeginIndexed(obj, st, small)
    assert e.value.code == abi.RTB_E_SIZE


@pytest.mark.parametrize("name", ["test_scene_1", "eval_scene"])
def test_rotation_gif_byte_identical(rt, oracle, tmp_path, name):
    """GenerateRotationFrames + SaveGifAsync: the fused library sweep, the two-step mirror (frames -> SaveGif with the device
    palette kernel) and the oracle's SaveGif over the same frames must write the same file."""
    from PIL import Image
    obj = synth.sample_scene(name)
    st = scene_mod.RenderSettings(**BASE)
    g = gif.GifGenerator(rt, obj)
    frames = g.GenerateRotationFrames(st)
    assert len(frames) == 36
    fused, two_step, ref = str(tmp_path / "fused.gif"), str(tmp_path / "two_step.gif"), str(tmp_path / "oracle.gif")
    g.RenderRotationGif(st, fused)
    g.SaveGif(frames, two_step)
    oracle.gif_save(ref, np.stack([f.pixels for f in frames]), 10)
    want = open(ref, "rb").read()
    assert open(two_step, "rb").read() == want
    assert open(fused, "rb").read() == want
    im = Image.open(fused)
    assert im.n_frames == 36 and im.size == (160, 120) and im.info.get("loop") == 0


def test_rotation_gif_honours_frame_count_step_and_delay(rt, oracle, tmp_path):
    obj = synth.sample_scene("test_scene_2")
    st = scene_mod.RenderSettings(**BASE)
    g = gif.GifGenerator(rt, obj)
    path = str(tmp_path / "five.gif")
    g.RenderRotationGif(st, path, frameDelay=4, totalFrames=5, stepDeg=72.0, threads=2)
    frames = []
    for k in range(5):
        s = scene_mod.RenderSettings(**{**BASE, "CameraRotationOverride": (-60.0, 0.0, 72.0 * k)})
        px = np.zeros((120, 160, 4), np.uint8)
        rt.RenderInto(obj, s, px)
        frames.append(px)
    ref = str(tmp_path / "five_oracle.gif")
    oracle.gif_save(ref, np.stack(frames), 4)
    assert open(path, "rb").read() == open(ref, "rb").read()


def test_cpp_gif_generator_mirror(oracle, tmp_path):
    """include/rtb_raytracer.hpp: rtb::GifGenerator (sweep + SaveGif, and the fused call) writes the oracle's file."""
    import subprocess
    from test_host_cpu import build_cpp_test
    exe = build_cpp_test()
    obj = synth.sample_scene("test_scene_1")
    scene_file, out = tmp_path / "scene.txt", tmp_path / "out.gif"
    scene_file.write_bytes(synth.scene_to_text(obj).encode())
    r = subprocess.run([exe, "gif", str(scene_file), str(out), "96", "72", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    packed = scene_mod.pack_scene(obj)
    osc = oracle.OracleScene.from_desc(packed.desc)
    frames = []
    for k in range(36):
        st = scene_mod.RenderSettings(ResolutionOverride=(96, 72), MaxDepth=2, CameraPositionOverride=(0.0, 0.0, 0.0), CameraRotationOverride=(-60.0, 0.0, 10.0 * k))
        frames.append(osc.render(st.to_params())["rgba8"])
    ref = str(tmp_path / "oracle.gif")
    oracle.gif_save(ref, np.stack(frames), 10)
    assert open(out, "rb").read() == open(ref, "rb").read()


def test_fused_sweep_reproduces_the_committed_golden_gif(rt, tmp_path):
    """No oracle at run time: rtb_gif_render_rotation (4 frames, 90-degree steps, 48x36, depth 2) must write exactly
    tests/golden/rotation_test_scene_1_48x36.gif."""
    import os
    want = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rotation_test_scene_1_48x36.gif"), "rb").read()
    obj = synth.sample_scene("test_scene_1")
    st = scene_mod.RenderSettings(ResolutionOverride=(48, 36), MaxDepth=2, CameraPositionOverride=(0.0, 0.0, 0.0), CameraRotationOverride=(-60.0, 0.0, 0.0))
    path = str(tmp_path / "golden_check.gif")
    gif.GifGenerator(rt, obj).RenderRotationGif(st, path, frameDelay=10, totalFrames=4, stepDeg=90.0)
    assert open(path, "rb").read() == want
