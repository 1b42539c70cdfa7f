"""Regenerates tests/golden/rotation_test_scene_1_48x36.gif: the reference's rotation GIF (GifGenerator.cs:40-184) of
test_scene_1 reduced to 4 frames (Z = 0, 90, 180, 270 degrees), 48x36, depth 2, camera position override (0,0,0), rotation
override (-60, 0, Z), frame delay 10 — rendered and encoded by the CPU oracle (oracle.cpp + gif_oracle.cpp).

    python tests/golden/make_golden_gif.py

Needs only the committed scene copy (tests/golden/scenes).  Like the other goldens it pins the ORACLE against regressions and
gives the GPU path a fixture that needs no oracle at run time; it is not an output of the reference itself.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

scene_mod = importlib.import_module("cosig-raytracing_b200.scene")
synth = importlib.import_module("cosig-raytracing_b200.synth")
HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "rotation_test_scene_1_48x36.gif")
ANGLES = (0.0, 90.0, 180.0, 270.0)


def settings(angle):
    return scene_mod.RenderSettings(ResolutionOverride=(48, 36), MaxDepth=2, CameraPositionOverride=(0.0, 0.0, 0.0),
                                    CameraRotationOverride=(-60.0, 0.0, angle))


def oracle_frames():
    O.build()
    packed = scene_mod.pack_scene(synth.sample_scene("test_scene_1"))
    osc = O.OracleScene.from_desc(packed.desc)
    return np.stack([osc.render(settings(a).to_params())["rgba8"] for a in ANGLES])


def main():
    O.gif_save(PATH, oracle_frames(), 10)
    print(PATH, os.path.getsize(PATH), "bytes")


if __name__ == "__main__":
    main()
