"""Consumes golden frames dumped by the REAL reference (csharp/Editor/DumpGoldens.cs run in the Unity editor on the stock
RayTracer + BVHRayTracing.compute) when a maintainer has put them under tests/golden/unity/ — the one thing that can pin
oracle/oracle.cpp against pixels the reference itself produced (until then the oracle is "parity unpinned").

Absent the dump (it cannot be produced in the build image: no Unity), the real comparisons skip; the self-test below keeps the
consuming code exercised by feeding it a dump synthesised from the oracle in the dump's exact format.
"""
import json
import os

import numpy as np
import pytest

from util import GOLDEN, abi, oracle_scene, params, synth

UNITY = os.path.join(GOLDEN, "unity")


def load_dump(folder):
    """[(entry dict, rgba8 [h, w, 4] row 0 = bottom)] + manifest of a DumpGoldens.cs output folder."""
    man = json.load(open(os.path.join(folder, "manifest.json")))
    w, h = man["width"], man["height"]
    out = []
    for e in man["frames"]:
        raw = np.fromfile(os.path.join(folder, e["file"]), np.uint8)
        assert raw.size == w * h * 4, f"{e['file']}: {raw.size} bytes, expected {w * h * 4}"
        out.append((e, raw.reshape(h, w, 4)))
    return man, out


def srgb_encode(rgba):
    c = rgba[..., :3].astype(np.float64) / 255.0
    enc = np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055)
    out = rgba.copy()
    out[..., :3] = np.floor(np.clip(enc, 0, 1) * 255.0 + 0.5).astype(np.uint8)
    return out


def compare_dump(oracle, folder, render=None):
    """Every frame of the dump against the oracle (or `render(obj, p)` — the GPU path).  Returns a list of per-frame reports;
    raises on a frame outside the bar.  Diagnoses the one open question of SURVEY App. A.9 (sRGB encoding of the ARGB32 target)."""
    man, frames = load_dump(folder)
    reports = []
    for e, got in frames:
        obj = synth.sample_scene(e["scene"])
        p = params(man["width"], man["height"], man["max_depth"], e["aa_samples"], debug_mode=e["debug_mode"])
        if render is None:
            osc, holder = oracle_scene(oracle, obj)
            ours = osc.render(p)["rgba8"]
        else:
            ours = render(obj, p)
        d = np.abs(got[..., :3].astype(np.int32) - ours[..., :3].astype(np.int32)).max(axis=-1)
        within, same = float((d <= 1).mean()), float((d == 0).mean())
        rep = dict(frame=e["file"], within_1=within, identical=same, worst=int(d.max()))
        if within < 0.999:
            ds = np.abs(got[..., :3].astype(np.int32) - srgb_encode(ours)[..., :3].astype(np.int32)).max(axis=-1)
            if float((ds <= 1).mean()) >= 0.999:
                raise AssertionError(f"{e['file']}: the Unity frame matches the sRGB-ENCODED oracle frame — SURVEY App. A.9 resolves the other way on "
                                     f"{man.get('graphics_api')}: render with srgb_encode = 1")
            flipped = np.abs(got[::-1, :, :3].astype(np.int32) - ours[..., :3].astype(np.int32)).max(axis=-1)
            hint = " (it matches upside down: row order)" if float((flipped <= 1).mean()) >= 0.999 else ""
            raise AssertionError(f"{e['file']}: only {within * 100:.3f} % of pixels within 1/255 of the reference's frame, worst {rep['worst']}{hint}")
        reports.append(rep)
    return man, reports


@pytest.mark.skipif(not os.path.exists(os.path.join(UNITY, "manifest.json")), reason="no Unity dump under tests/golden/unity (csharp/Editor/DumpGoldens.cs)")
def test_oracle_matches_the_reference_frames(oracle):
    man, reports = compare_dump(oracle, UNITY)
    assert len(reports) >= 9
    print(json.dumps(dict(unity=man.get("unity_version"), api=man.get("graphics_api"), frames=reports), indent=1))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(UNITY, "manifest.json")), reason="no Unity dump under tests/golden/unity (csharp/Editor/DumpGoldens.cs)")
def test_gpu_matches_the_reference_frames(oracle):
    import importlib
    rt_mod = importlib.import_module("cosig-raytracing_b200.raytracer")
    with rt_mod.RayTracer(bvh_mode=abi.RTB_BVH_REFERENCE) as rt:
        compare_dump(oracle, UNITY, render=lambda obj, p: rt.RenderAsync(obj, p).pixels)


def test_dump_consumer_on_a_synthesised_dump(oracle, tmp_path):
    """The consumer itself: a dump in DumpGoldens.cs' format made from the oracle passes; the same dump sRGB-encoded, flipped or
    perturbed is rejected with the matching diagnosis."""
    w, h = 96, 72
    frames = []
    for scene in ("test_scene_1", "eval_scene"):
        obj = synth.sample_scene(scene)
        osc, holder = oracle_scene(oracle, obj)
        for aa, dbg, tag in ((1, 0, "aa1"), (4, 0, "aa4"), (1, 1, "depth")):
            img = osc.render(params(w, h, 3, aa, debug_mode=dbg))["rgba8"]
            img.tofile(tmp_path / f"{scene}_{tag}.rgba")
            frames.append(dict(scene=scene, file=f"{scene}_{tag}.rgba", aa_samples=aa, debug_mode=dbg))
    man = dict(unity_version="synthetic", graphics_api="none", color_space="Linear", width=w, height=h, max_depth=3, frames=frames)
    (tmp_path / "manifest.json").write_text(json.dumps(man))
    _, reports = compare_dump(oracle, str(tmp_path))
    assert len(reports) == 6 and all(r["identical"] == 1.0 for r in reports)
    first = np.fromfile(tmp_path / frames[0]["file"], np.uint8).reshape(h, w, 4)
    srgb_encode(first).tofile(tmp_path / frames[0]["file"])
    with pytest.raises(AssertionError, match="sRGB-ENCODED"):
        compare_dump(oracle, str(tmp_path))
    first[::-1].copy().tofile(tmp_path / frames[0]["file"])
    with pytest.raises(AssertionError, match="upside down"):
        compare_dump(oracle, str(tmp_path))
    noisy = first.copy()
    noisy[::3, ::3, :3] ^= 0x10
    noisy.tofile(tmp_path / frames[0]["file"])
    with pytest.raises(AssertionError, match="within 1/255"):
        compare_dump(oracle, str(tmp_path))
