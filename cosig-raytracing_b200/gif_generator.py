"""Host-side mirror of the render half of the reference's GifGenerator (Assets/Services/GifGenerator.cs:40-72): the
36-frame rotation sweep that calls RenderAsync once per angle.  Here the 36 frames go through the pipelined host API
(RayTracer.RenderBegin / RenderEnd: frames in flight on both lanes, readback overlapped), which is what makes a batch caller
faster than 36 blocking renders.  GIF encoding (GifGenerator.cs:82-501: palette, LZW) is a CPU post-process outside the render
path and is not restated; frames are returned as Texture2D like the reference's List<Texture2D>.
"""
from __future__ import annotations

import copy
from typing import Callable, List, Optional

import numpy as np

from .raytracer import RayTracer, Texture2D
from .scene import ObjectData, RenderSettings


class GifGenerator:
    def __init__(self, rayTracer: RayTracer, scene: ObjectData):  # GifGenerator.cs:27-31
        self.rayTracer = rayTracer
        self.scene = scene

    def GenerateRotationFrames(self, baseSettings: RenderSettings, progress: Optional[Callable[[float, str], None]] = None,
                               token=None, in_flight: int = 4) -> List[Texture2D]:
        """360 degrees around Z in 10-degree steps (GifGenerator.cs:47-69): frame k renders with
        CameraRotationOverride = (base.x, base.y, 10 k).  `token.is_set()` aborts between frames."""
        total = 36
        base = baseSettings.CameraRotationOverride or (0.0, 0.0, 0.0)
        w, h = self.rayTracer.resolve(self.scene, baseSettings)
        frames: List[Texture2D] = []
        pending = []  # (ticket, pixels)
        for index, angle in enumerate(range(0, 360, 10)):
            if token is not None and token.is_set():
                break
            if progress:
                progress(index / total, f"Rendering frame {index + 1}/{total} (Z={angle}°)")
            settings = copy.copy(baseSettings)
            settings.CameraRotationOverride = (base[0], base[1], float(angle))
            pixels = np.empty((h, w, 4), np.uint8)
            pending.append((self.rayTracer.RenderBegin(self.scene, settings, pixels), pixels))
            if len(pending) >= in_flight:
                ticket, done = pending.pop(0)
                self.rayTracer.RenderEnd(ticket)
                frames.append(Texture2D(done))
        for ticket, done in pending:
            self.rayTracer.RenderEnd(ticket)
            frames.append(Texture2D(done))
        return frames
