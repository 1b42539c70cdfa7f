"""Host-side mirror of the reference's GifGenerator (Assets/Services/GifGenerator.cs) over the C ABI.

* GenerateRotationFrames (:40-72): the 36-frame rotation sweep, one RenderAsync per angle in the reference; here the frames go
  through the pipelined host API (RayTracer.RenderBegin / RenderEnd: frames in flight on the lanes, readback on a copy stream).
* SaveGif / SaveGifAsync (:82-184): palette mapping (ConvertToIndexed :346-369) runs as a CUDA kernel, LZW (:411-501) on host
  threads inside librtb200.so; the file is byte-identical to what the reference's algorithm writes (tests/test_gif_*.py).
* RenderRotationGif: both halves fused inside the library (rtb_gif_render_rotation): frames never leave the GPU as RGBA — they
  are palette-mapped on the device, read back as 1 byte per pixel and compressed while the following frames render.

Nothing here computes pixels or codes; it marshals arguments.
"""
from __future__ import annotations

import copy
import ctypes as C
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import abi
from .raytracer import RayTracer, RtbError, Texture2D
from .scene import ObjectData, RenderSettings


class GifGenerator:
    TOTAL_FRAMES = 36   # 360 degrees at 10-degree increments, GifGenerator.cs:47
    STEP_DEG = 10.0

    def __init__(self, rayTracer: RayTracer, scene: ObjectData):  # GifGenerator.cs:27-31
        self.rayTracer = rayTracer
        self.scene = scene

    def GenerateRotationFrames(self, baseSettings: RenderSettings, progress: Optional[Callable[[float, str], None]] = None,
                               token=None, in_flight: int = 4) -> List[Texture2D]:
        """360 degrees around Z in 10-degree steps (GifGenerator.cs:47-69): frame k renders with
        CameraRotationOverride = (base.x, base.y, 10 k).  `token.is_set()` aborts between frames."""
        total = self.TOTAL_FRAMES
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

    def SaveGif(self, frames: Sequence[Texture2D], filePath: str, frameDelay: int = 10, threads: int = 0) -> None:
        """GifGenerator.cs:160-184.  `frames == null || Count == 0` returns without writing a file (:162)."""
        if not frames:
            return
        pixels = [np.ascontiguousarray(f.pixels if isinstance(f, Texture2D) else f, np.uint8) for f in frames]
        h, w = pixels[0].shape[:2]
        for p in pixels:
            if p.shape != (h, w, 4):
                raise ValueError("all frames must be [height, width, 4] uint8 of the first frame's size")
        ptrs = (C.c_void_p * len(pixels))(*[p.ctypes.data for p in pixels])
        lib = abi.load()
        rc = lib.rtb_gif_save(self.rayTracer._ctx, filePath.encode(), w, h, ptrs, len(pixels), int(frameDelay), int(threads))
        self.rayTracer._check(rc)

    def SaveGifAsync(self, frames: Sequence[Texture2D], filePath: str, progress: Optional[Callable[[float, str], None]] = None,
                     frameDelay: int = 10) -> None:
        """GifGenerator.cs:82-155: same bytes as SaveGif; the reference's progress milestones are reported around the call."""
        if not frames:
            return
        if progress:
            progress(0.0, "Starting GIF encoding...")
            progress(0.2, "Compressing frames...")
        self.SaveGif(frames, filePath, frameDelay)
        if progress:
            progress(1.0, "GIF saved!")

    def RenderRotationGif(self, baseSettings: RenderSettings, filePath: str, frameDelay: int = 10, totalFrames: int = TOTAL_FRAMES,
                          stepDeg: float = STEP_DEG, threads: int = 0) -> None:
        """GenerateRotationFrames + SaveGifAsync in one library call (rtb_gif_render_rotation)."""
        rt = self.rayTracer
        if not rt._ensure_scene(self.scene):
            raise RtbError(abi.RTB_E_NOSCENE, "no scene")
        p = rt._params(baseSettings)
        rt._check(abi.load().rtb_gif_render_rotation(rt._ctx, C.byref(p), int(totalFrames), float(stepDeg), filePath.encode(), int(frameDelay),
                                                     int(threads)))


def color_table() -> np.ndarray:
    """GenerateColorTable, GifGenerator.cs:219-247: 256 x RGB."""
    t = np.zeros(768, np.uint8)
    abi.load().rtb_gif_color_table(t.ctypes.data)
    return t.reshape(256, 3)


def lzw_compress(indexed: np.ndarray) -> bytes:
    """LzwCompress, GifGenerator.cs:411-501 (host code inside librtb200.so)."""
    data = np.ascontiguousarray(indexed, np.uint8).reshape(-1)
    lib = abi.load()
    cap = lib.rtb_gif_lzw_bound(data.size)
    out = np.zeros(cap, np.uint8)
    n = lib.rtb_gif_lzw(data.ctypes.data, data.size, out.ctypes.data, cap)
    if n < 0:
        raise RtbError(int(n), "rtb_gif_lzw failed")
    return out[:n].tobytes()


def save_indexed(filePath: str, frames: Sequence[np.ndarray], frameDelay: int = 10, threads: int = 0) -> None:
    """SaveGifAsync for frames that already are palette indices ([height, width] uint8, top row first)."""
    arrs = [np.ascontiguousarray(f, np.uint8) for f in frames]
    h, w = arrs[0].shape
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    rc = abi.load().rtb_gif_save_indexed(filePath.encode(), w, h, ptrs, len(arrs), int(frameDelay), int(threads))
    if rc != abi.RTB_OK:
        raise RtbError(rc, f"rtb_gif_save_indexed failed for {filePath}")
