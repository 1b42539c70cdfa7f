"""Host-side mirror of the reference's render API, over the C ABI of librtb200.so.

`RayTracer` has the public surface of Assets/Services/RayTracer.cs:17 (same method names and argument meaning):
InvalidateBVHCache :38, ReleaseBuffers :47, ClearRenderTarget :65, RenderToTexture :82, RenderAsync :212, SaveTexture :504.
`SceneService.LoadScene` mirrors Assets/Services/SceneService.cs:26 through the library's native parser.

Error behaviour follows the reference where it has one: rendering before a scene is known returns None (the reference
returns null when its shader is missing, RayTracer.cs:84-88; a cancelled render also returns null, :283); everything the
reference would throw on surfaces as RtbError.  There is no CPU fallback — without the CUDA library or a GPU every
render call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import abi
from .scene import ObjectData, PackedScene, RenderSettings, pack_scene, unpack_scene


class RtbError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"rtb error {code}: {message}")
        self.code = code


class Texture2D:
    """What RenderAsync returns: RGBA32 pixels, row 0 = bottom (Unity's Texture2D convention)."""

    def __init__(self, pixels: np.ndarray):
        self.pixels = pixels  # uint8 [height, width, 4]

    @property
    def width(self) -> int:
        return self.pixels.shape[1]

    @property
    def height(self) -> int:
        return self.pixels.shape[0]

    def top_down_rgb(self) -> np.ndarray:
        return self.pixels[::-1, :, :3]


class DeviceTexture:
    """What RenderToTexture returns: the frame left in device memory (no readback)."""

    def __init__(self, ptr: int, width: int, height: int):
        self.ptr, self.width, self.height = ptr, width, height


class RayTracer:
    def __init__(self, devices=None, bvh_mode: int = abi.RTB_BVH_REFERENCE, primitive_mode: int = abi.RTB_PRIM_TESSELLATED):
        self._lib = abi.load()
        self._ctx = C.c_void_p()
        ids = None
        n = 0
        if devices is not None:
            n = len(devices)
            ids = (C.c_int32 * n)(*devices)
        rc = self._lib.rtb_create(C.byref(self._ctx), ids, n)
        if rc != abi.RTB_OK:
            raise RtbError(rc, (self._lib.rtb_last_error(None) or b"").decode())
        self.bvh_mode = bvh_mode
        self.primitive_mode = primitive_mode
        self._cached_scene = None      # RayTracer.cs:118: the BVH cache key is the scene object's identity
        self._packed: Optional[PackedScene] = None
        self._needs_rebuild = True
        self._pinned = None
        self._pinned_bytes = 0

    # ---- plumbing --------------------------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != abi.RTB_OK:
            raise RtbError(rc, (self._lib.rtb_last_error(self._ctx) or b"").decode())

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            if self._pinned:
                self._lib.rtb_free_pinned(self._pinned)
                self._pinned = None
            self._lib.rtb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ensure_scene(self, scene) -> bool:
        if scene is None:
            return False
        if self._needs_rebuild or self._cached_scene is not scene:  # RayTracer.cs:118-123, 273-278
            desc_holder = scene if isinstance(scene, PackedScene) else pack_scene(scene)
            self._check(self._lib.rtb_upload_scene(self._ctx, desc_holder.ptr(), self.primitive_mode, self.bvh_mode))
            self._packed = desc_holder
            self._cached_scene = scene
            self._needs_rebuild = False
        return True

    def _pinned_buffer(self, nbytes: int):
        if self._pinned_bytes < nbytes:
            if self._pinned:
                self._lib.rtb_free_pinned(self._pinned)
            self._pinned = self._lib.rtb_alloc_pinned(nbytes)
            if not self._pinned:
                raise RtbError(abi.RTB_E_CUDA, "cudaHostAlloc failed")
            self._pinned_bytes = nbytes
        return self._pinned

    @staticmethod
    def _params(settings) -> abi.RenderParams:
        return settings.to_params() if isinstance(settings, RenderSettings) else settings

    # ---- the reference's public surface ----------------------------------------------------------------------------
    def InvalidateBVHCache(self):  # RayTracer.cs:38-42
        self._needs_rebuild = True
        self._check(self._lib.rtb_invalidate(self._ctx))

    def ReleaseBuffers(self):  # RayTracer.cs:47-59
        self._cached_scene = None
        self._needs_rebuild = True
        self._check(self._lib.rtb_invalidate(self._ctx))
        self._check(self._lib.rtb_clear_target(self._ctx))

    def ClearRenderTarget(self):  # RayTracer.cs:65-72
        self._check(self._lib.rtb_clear_target(self._ctx))

    def resolve(self, scene, settings):
        """Resolved (width, height) for (scene, settings), RayTracer.cs:221-222."""
        holder = scene if isinstance(scene, PackedScene) else (self._packed if scene is self._cached_scene and self._packed else pack_scene(scene))
        wh = (C.c_int32 * 2)()
        p = self._params(settings)
        rc = self._lib.rtb_resolve_frame(holder.ptr(), C.byref(p), None, wh)
        if rc != abi.RTB_OK:
            raise RtbError(rc, (self._lib.rtb_last_error(None) or b"").decode())
        return wh[0], wh[1]

    def RenderAsync(self, scene, settings, progress=None, token=None) -> Optional[Texture2D]:  # RayTracer.cs:212-380
        """Blocking here (the reference's Task only yields to Unity's frame loop).  `token`, if given, is an object with
        an `is_set()` method or a ctypes c_int32 polled between wavefront depths."""
        if not self._ensure_scene(scene):
            return None
        if progress:
            progress(0.1)
        p = self._params(settings)
        w, h = self.resolve(self._cached_scene, p)
        flag = None
        if token is not None:
            if hasattr(token, "is_set"):
                if token.is_set():
                    return None
            else:
                flag = token
                self._check(self._lib.rtb_set_cancel_flag(self._ctx, C.addressof(flag)))
        nbytes = w * h * 4
        buf = self._pinned_buffer(nbytes)
        ow, oh = C.c_int32(), C.c_int32()
        rc = self._lib.rtb_render(self._ctx, C.byref(p), buf, nbytes, C.byref(ow), C.byref(oh))
        if flag is not None:
            self._lib.rtb_set_cancel_flag(self._ctx, None)
        if rc == abi.RTB_E_CANCELLED:
            return None
        self._check(rc)
        if progress:
            progress(1.0)
        pixels = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), shape=(h, w, 4)).copy()
        return Texture2D(pixels)

    def RenderInto(self, scene, settings, out: np.ndarray) -> None:
        """RenderAsync into a caller-owned uint8 [h, w, 4] array (no intermediate copy when it is pinned)."""
        if not self._ensure_scene(scene):
            raise RtbError(abi.RTB_E_NOSCENE, "no scene")
        p = self._params(settings)
        self._check(self._lib.rtb_render(self._ctx, C.byref(p), out.ctypes.data, out.nbytes, None, None))

    def RenderBegin(self, scene, settings, out: np.ndarray) -> int:
        """Pipelined RenderInto: enqueue the frame and its readback into `out` (ideally pinned), return a ticket for RenderEnd."""
        if not self._ensure_scene(scene):
            raise RtbError(abi.RTB_E_NOSCENE, "no scene")
        p = self._params(settings)
        ticket = C.c_int32()
        self._check(self._lib.rtb_render_begin(self._ctx, C.byref(p), out.ctypes.data, out.nbytes, C.byref(ticket)))
        return ticket.value

    def RenderEnd(self, ticket: int) -> None:
        self._check(self._lib.rtb_render_end(self._ctx, ticket))

    def RenderBeginIndexed(self, scene, settings, out: np.ndarray) -> int:
        """RenderBegin whose frame is mapped to GIF palette indices on the device (GifGenerator.ConvertToIndexed): `out` receives
        width*height bytes, top row first."""
        if not self._ensure_scene(scene):
            raise RtbError(abi.RTB_E_NOSCENE, "no scene")
        p = self._params(settings)
        ticket = C.c_int32()
        self._check(self._lib.rtb_render_begin_indexed(self._ctx, C.byref(p), out.ctypes.data, out.nbytes, C.byref(ticket)))
        return ticket.value

    def RenderToTexture(self, scene, settings, dst_ptr: Optional[int] = None, dst_bytes: int = 0, sync: bool = True) -> Optional[DeviceTexture]:
        """RayTracer.cs:82-202: render and leave the frame on the device.  With dst_ptr the frame (or this rank's bands) is
        written there — a torch tensor's data_ptr(), or a peer mapping from frame_import for the NVLink gather."""
        if not self._ensure_scene(scene):
            return None
        p = self._params(settings)
        w, h = self.resolve(self._cached_scene, p)
        if dst_ptr is None:
            ptr = C.c_void_p()
            self._check(self._lib.rtb_frame_export(self._ctx, w * h * 4, C.byref(ptr), None))  # the context's own frame, not exported
            dst_ptr, dst_bytes = ptr.value, w * h * 4
        self._check(self._lib.rtb_render_device(self._ctx, C.byref(p), dst_ptr, dst_bytes, 1 if sync else 0))
        return DeviceTexture(dst_ptr, w, h)

    @staticmethod
    def SaveTexture(tex: Texture2D, path: str):  # RayTracer.cs:504-509 (EncodeToPNG)
        from PIL import Image
        Image.fromarray(np.ascontiguousarray(tex.pixels[::-1])).save(path)

    # ---- additions used by tests and the bench ----------------------------------------------------------------------
    def primary_hits(self, scene, settings):
        """(prim_id, t, material) maps of the pixel-centre primary rays, each [h, w], row 0 = bottom."""
        if not self._ensure_scene(scene):
            raise RtbError(abi.RTB_E_NOSCENE, "no scene")
        p = self._params(settings)
        w, h = self.resolve(self._cached_scene, p)
        prim = np.zeros((h, w), np.int32)
        t = np.zeros((h, w), np.float32)
        mat = np.zeros((h, w), np.int32)
        self._check(self._lib.rtb_render_aux(self._ctx, C.byref(p), prim.ctypes.data, t.ctypes.data, mat.ctypes.data))
        return prim, t, mat

    def triangles(self):
        n = C.c_int64()
        self._check(self._lib.rtb_get_triangles(self._ctx, None, None, 0, C.byref(n)))
        vn = np.zeros((n.value, 18), np.float32)
        mat = np.zeros(n.value, np.int32)
        if n.value:
            self._check(self._lib.rtb_get_triangles(self._ctx, vn.ctypes.data, mat.ctypes.data, n.value, C.byref(n)))
        return vn, mat

    def bvh(self):
        nn = C.c_int64()
        self._check(self._lib.rtb_get_bvh(self._ctx, None, 0, C.byref(nn), None, 0))
        words = self._lib.rtb_get_bvh_node_words(self._ctx)
        nodes = np.zeros((nn.value, words), np.float32)
        st = self.stats()
        perm = np.zeros(st.n_triangles, np.int32)
        self._check(self._lib.rtb_get_bvh(self._ctx, nodes.ctypes.data, nodes.nbytes, C.byref(nn), perm.ctypes.data, perm.size))
        return nodes, perm

    def stats(self) -> abi.Stats:
        s = abi.Stats()
        self._check(self._lib.rtb_get_stats(self._ctx, C.byref(s)))
        return s

    def set_profiling(self, on: bool):
        self._check(self._lib.rtb_set_profiling(self._ctx, 1 if on else 0))

    def stream(self, index: int = 0) -> int:
        """cudaStream_t of device `index` (wrap with torch.cuda.ExternalStream to record events on it)."""
        return self._lib.rtb_get_stream(self._ctx, index)

    def flush(self):
        """Make stream(0) wait for every frame in flight on the context's second stream (see rtb_flush)."""
        self._check(self._lib.rtb_flush(self._ctx))

    def synchronize(self):
        self._check(self._lib.rtb_synchronize(self._ctx))

    def frame_export(self, nbytes: int):
        ptr = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        self._check(self._lib.rtb_frame_export(self._ctx, nbytes, C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def frame_import(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        h = (C.c_uint8 * 64)(*handle)
        self._check(self._lib.rtb_frame_import(self._ctx, h, C.byref(ptr)))
        return ptr.value

    # ---- process-per-GPU frame ring (rtb_group_*) ------------------------------------------------------------------------------
    def group_create(self, rank: int, world: int, frame_bytes: int, n_buffers: int = 4, handle: Optional[bytes] = None) -> bytes:
        """Rank 0 returns the 64-byte handle the other ranks must pass in (send it over torch.distributed, MPI, a pipe ...)."""
        h = (C.c_uint8 * 64)(*(handle or bytes(64)))
        self._check(self._lib.rtb_group_create(self._ctx, rank, world, frame_bytes, n_buffers, h))
        return bytes(h)

    def GroupRenderBegin(self, scene, settings, out: Optional[np.ndarray] = None) -> int:
        """Enqueue this rank's bands of the next frame; on rank 0 `out` (pinned uint8 [h, w, 4]) receives the whole frame."""
        if not self._ensure_scene(scene):
            raise RtbError(abi.RTB_E_NOSCENE, "no scene")
        p = self._params(settings)
        ticket = C.c_int32()
        ptr, n = (out.ctypes.data, out.nbytes) if out is not None else (None, 0)
        self._check(self._lib.rtb_group_render_begin(self._ctx, C.byref(p), ptr, n, C.byref(ticket)))
        return ticket.value

    def GroupRenderEnd(self, ticket: int) -> None:
        self._check(self._lib.rtb_group_render_end(self._ctx, ticket))

    def group_create_host(self, rank: int, world: int, frame_bytes: int, n_buffers: int = 4, name: Optional[str] = None) -> str:
        """Host-ring form: frames land in POSIX shared memory every rank has page-locked, each rank's bands over its own PCIe link.
        Rank 0 returns the name of the shared object (send it to the other ranks, which pass it in); use GroupRenderBegin(..., out=None)
        and group_frame(ticket, height, width) afterwards."""
        if name is None:
            if rank != 0:
                raise ValueError("ranks other than 0 need the name rank 0 returned")
            import os
            name = f"/rtb200-{os.getpid()}-{id(self) & 0xffffff:x}"
        self._check(self._lib.rtb_group_create_host(self._ctx, rank, world, frame_bytes, n_buffers, name.encode()))
        return name

    def group_frame(self, ticket: int, height: int, width: int) -> np.ndarray:
        """Rank 0 of a host-ring group: the frame of `ticket` as a [height, width, 4] uint8 VIEW of the ring (valid until n_buffers more
        frames have been begun; copy it to keep it)."""
        ptr = C.c_void_p()
        self._check(self._lib.rtb_group_frame(self._ctx, ticket, C.byref(ptr)))
        buf = (C.c_uint8 * (height * width * 4)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=np.uint8).reshape(height, width, 4)

    def group_destroy(self):
        self._check(self._lib.rtb_group_destroy(self._ctx))

    def external_import(self, handle_type: int, handle: int, nbytes: int, dedicated: bool = False) -> int:
        """Maps a graphics-API allocation (POSIX fd / NT handle) into the device's address space; returns a device pointer usable
        as RenderToTexture's dst_ptr (the zero-copy realtime path)."""
        ptr = C.c_void_p()
        self._check(self._lib.rtb_external_import(self._ctx, handle_type, C.c_void_p(handle), nbytes, 1 if dedicated else 0, C.byref(ptr)))
        return ptr.value

    def external_release(self, ptr: int):
        self._check(self._lib.rtb_external_release(self._ctx, C.c_void_p(ptr)))

    def frame_read(self, out: np.ndarray):
        self._check(self._lib.rtb_frame_read(self._ctx, out.ctypes.data, out.nbytes))


class SceneService:
    """Assets/Services/SceneService.cs: LoadScene(path) -> ObjectData, parsed by the library's native parser."""

    @staticmethod
    def LoadScene(path: str) -> ObjectData:
        lib = abi.load()
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib.rtb_scene_load(path.encode(), C.byref(h), err, 512)
        if rc == abi.RTB_E_IO:
            return ObjectData()  # the reference logs an error and returns an empty scene, SceneService.cs:28-33
        if rc != abi.RTB_OK:
            raise RtbError(rc, err.value.decode())
        try:
            return unpack_scene(lib.rtb_scene_get(h).contents)
        finally:
            lib.rtb_scene_free(h)

    @staticmethod
    def ParseScene(text: bytes) -> ObjectData:
        lib = abi.load()
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib.rtb_scene_parse(text, len(text), C.byref(h), err, 512)
        if rc != abi.RTB_OK:
            raise RtbError(rc, err.value.decode())
        try:
            return unpack_scene(lib.rtb_scene_get(h).contents)
        finally:
            lib.rtb_scene_free(h)
