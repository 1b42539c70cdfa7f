"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_abi = importlib.import_module("cosig-raytracing_b200.abi")


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("rays_primary", "rays_continuation", "rays_shadow", "nodes_visited", "tris_tested",
                                         "closest_hits", "primary_hits")] + [("max_stack", C.c_int32), ("threads", C.c_int32),
                                                                             ("seconds", C.c_double), ("nodes_visited_shadow", C.c_int64),
                                                                             ("tris_tested_shadow", C.c_int64)]

    @property
    def rays(self):
        return self.rays_primary + self.rays_continuation + self.rays_shadow


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, "oracle.cpp"), os.path.join(_HERE, "gif_oracle.cpp"), os.path.join(_HERE, "..", "include", "rtb.h")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(p) for p in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        VP = C.c_void_p
        L.orc_build.argtypes = [C.POINTER(_abi.SceneDesc), C.POINTER(VP)]
        L.orc_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(VP), C.c_char_p, C.c_size_t]
        L.orc_load.argtypes = [C.c_char_p, C.POINTER(VP), C.c_char_p, C.c_size_t]
        L.orc_free.argtypes = [VP]
        L.orc_desc.argtypes = [VP]
        L.orc_desc.restype = C.POINTER(_abi.SceneDesc)
        for n in ("orc_n_triangles", "orc_n_nodes"):
            getattr(L, n).argtypes = [VP]
            getattr(L, n).restype = C.c_int64
        L.orc_max_leaf.argtypes = [VP]
        L.orc_set_primitive_mode.argtypes = [VP, C.c_int32]
        L.orc_n_primitives.argtypes = [VP]
        L.orc_n_primitives.restype = C.c_int64
        L.orc_get_triangles.argtypes = [VP, VP, VP, VP]
        L.orc_get_bvh.argtypes = [VP, VP, VP]
        L.orc_resolve.argtypes = [VP, C.POINTER(_abi.RenderParams), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.orc_get_frame.argtypes = [VP, C.POINTER(_abi.RenderParams), VP]
        L.orc_render.argtypes = [VP, C.POINTER(_abi.RenderParams), C.c_int32, C.c_int32, C.c_int32, C.c_int32, VP, VP, VP, VP, VP,
                                 C.POINTER(Counters)]
        L.orc_primary_ray.argtypes = [VP, C.POINTER(_abi.RenderParams), C.c_int32, C.c_int32, VP, VP]
        L.orc_brute_closest.argtypes = [VP, VP, VP, VP, VP, C.c_int32]
        L.orc_random_unit_vector.argtypes = [VP, VP]
        L.orc_sample_ray.argtypes = [VP, C.POINTER(_abi.RenderParams), C.c_int32, C.c_int32, C.c_int32, VP, VP]
        L.orc_set_leaf_accel.argtypes = [C.c_int32]
        L.orc_set_exact_closest.argtypes = [VP, C.c_int32]
        L.orc_n_accelerated_leaves.argtypes = [VP]
        L.orc_gif_color_table.argtypes = [VP]
        L.orc_gif_convert_to_indexed.argtypes = [VP, C.c_int32, C.c_int32, VP]
        L.orc_gif_lzw.argtypes = [VP, C.c_int64, VP, C.c_int64]
        L.orc_gif_lzw.restype = C.c_int64
        L.orc_gif_save.argtypes = [C.c_char_p, C.c_int32, C.c_int32, VP, C.c_int32, C.c_int32]
        _lib = L
    return _lib


class OracleScene:
    def __init__(self, handle):
        self.h = handle

    @staticmethod
    def from_desc(desc) -> "OracleScene":
        h = C.c_void_p()
        rc = lib().orc_build(C.byref(desc), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"orc_build failed: {rc}")
        return OracleScene(h)

    @staticmethod
    def from_text(text: bytes) -> "OracleScene":
        h = C.c_void_p()
        err = C.create_string_buffer(256)
        rc = lib().orc_parse(text, len(text), C.byref(h), err, 256)
        if rc != 0:
            raise ValueError(f"orc_parse failed: {rc} {err.value.decode()}")
        return OracleScene(h)

    @staticmethod
    def from_file(path: str) -> "OracleScene":
        with open(path, "rb") as f:
            return OracleScene.from_text(f.read())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_free(self.h)
            self.h = None

    def set_exact_closest(self, on: bool):
        """Checker-only: the traversal returns the exact closest hit (what a scan of all triangles finds) instead of the reference's
        answer, which now and then loses the nearer of two nearly coincident surfaces to its FP32 cull (oracle.cpp: exact_closest)."""
        if lib().orc_set_exact_closest(self.h, 1 if on else 0) != 0:
            raise RuntimeError("orc_set_exact_closest failed")

    def set_primitive_mode(self, mode: int):
        """0 = tessellated (reference behaviour), 1 = analytic spheres / boxes (HittableObjects.cs semantics)."""
        if lib().orc_set_primitive_mode(self.h, mode) != 0:
            raise ValueError("bad primitive mode")
        return self

    @property
    def n_primitives(self) -> int:
        return lib().orc_n_primitives(self.h)

    @property
    def desc(self):
        return lib().orc_desc(self.h).contents

    @property
    def n_triangles(self) -> int:
        return lib().orc_n_triangles(self.h)

    @property
    def n_nodes(self) -> int:
        return lib().orc_n_nodes(self.h)

    @property
    def max_leaf(self) -> int:
        return lib().orc_max_leaf(self.h)

    @property
    def n_accelerated_leaves(self) -> int:
        return lib().orc_n_accelerated_leaves(self.h)

    def triangles(self):
        n = self.n_triangles
        vn = np.zeros((n, 18), np.float32)
        mat = np.zeros(n, np.int32)
        cen = np.zeros((n, 3), np.float32)
        lib().orc_get_triangles(self.h, vn.ctypes.data, mat.ctypes.data, cen.ctypes.data)
        return vn, mat, cen

    def bvh(self):
        nodes = np.zeros((self.n_nodes, 8), np.float32)
        orig = np.zeros(self.n_triangles, np.int32)
        lib().orc_get_bvh(self.h, nodes.ctypes.data, orig.ctypes.data)
        return nodes, orig

    def resolve(self, params):
        w, h = C.c_int32(), C.c_int32()
        lib().orc_resolve(self.h, C.byref(params), C.byref(w), C.byref(h))
        return w.value, h.value

    def frame(self, params):
        out = np.zeros(25, np.float32)
        lib().orc_get_frame(self.h, C.byref(params), out.ctypes.data)
        return out

    def render(self, params, rows=(0, -1, 1), threads=0, want_rgbf=False, want_aux=False):
        """Returns dict(rgba8[h,w,4], rgbf, prim, t, mat, counters); rows=(begin,end,step), row 0 = bottom."""
        w, h = self.resolve(params)
        rgba = np.zeros((h, w, 4), np.uint8)
        rgbf = np.zeros((h, w, 3), np.float32) if want_rgbf else None
        prim = np.full((h, w), -2, np.int32) if want_aux else None
        t = np.zeros((h, w), np.float32) if want_aux else None
        mat = np.full((h, w), -2, np.int32) if want_aux else None
        cnt = Counters()
        p = lambda a: a.ctypes.data if a is not None else None
        rc = lib().orc_render(self.h, C.byref(params), rows[0], rows[1], rows[2], threads, p(rgba), p(rgbf), p(prim), p(t), p(mat),
                              C.byref(cnt))
        if rc != 0:
            raise RuntimeError(f"orc_render failed: {rc}")
        return dict(rgba8=rgba, rgbf=rgbf, prim=prim, t=t, mat=mat, counters=cnt, width=w, height=h)

    def primary_ray(self, params, px, py):
        o = np.zeros(3, np.float32)
        d = np.zeros(3, np.float32)
        lib().orc_primary_ray(self.h, C.byref(params), px, py, o.ctypes.data, d.ctypes.data)
        return o, d

    def sample_ray(self, params, px, py, sample):
        o = np.zeros(3, np.float32)
        d = np.zeros(3, np.float32)
        lib().orc_sample_ray(self.h, C.byref(params), px, py, sample, o.ctypes.data, d.ctypes.data)
        return o, d

    def brute_closest(self, o, d, cap=16):
        o = np.ascontiguousarray(o, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        t = C.c_float()
        ids = np.zeros(cap, np.int32)
        n = lib().orc_brute_closest(self.h, o.ctypes.data, d.ctypes.data, C.byref(t), ids.ctypes.data, cap)
        return t.value, ids[:min(n, cap)].copy(), n


def set_leaf_accel(min_count: int) -> int:
    """Checker-only: leaves of the reference-shape BVH with more than `min_count` triangles are searched through a private
    sub-tree instead of linearly (same results, see oracle.cpp: LeafAccel); 0 = never.  Applies to scenes built afterwards;
    returns the previous setting."""
    return lib().orc_set_leaf_accel(min_count)


def random_unit_vector(seed) -> np.ndarray:
    s = np.ascontiguousarray(seed, np.float32)
    out = np.zeros(3, np.float32)
    lib().orc_random_unit_vector(s.ctypes.data, out.ctypes.data)
    return out


# ---- GIF writer restatement (oracle/gif_oracle.cpp: GifGenerator.cs) ------------------------------------------------------
def gif_color_table() -> np.ndarray:
    t = np.zeros(768, np.uint8)
    lib().orc_gif_color_table(t.ctypes.data)
    return t


def gif_convert_to_indexed(rgba8: np.ndarray) -> np.ndarray:
    """rgba8: [h, w, 4] uint8, row 0 = bottom.  Returns [h, w] palette indices, top row first."""
    rgba8 = np.ascontiguousarray(rgba8, np.uint8)
    h, w = rgba8.shape[:2]
    out = np.zeros((h, w), np.uint8)
    lib().orc_gif_convert_to_indexed(rgba8.ctypes.data, w, h, out.ctypes.data)
    return out


def gif_lzw(indexed: np.ndarray) -> bytes:
    data = np.ascontiguousarray(indexed, np.uint8).reshape(-1)
    cap = (data.size + 2) * 2 + 16
    out = np.zeros(cap, np.uint8)
    n = lib().orc_gif_lzw(data.ctypes.data, data.size, out.ctypes.data, cap)
    assert 0 <= n <= cap
    return out[:n].tobytes()


def gif_save(path: str, frames: np.ndarray, frame_delay: int = 10) -> None:
    """frames: [n, h, w, 4] uint8 RGBA, row 0 = bottom (Texture2D order)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape[:3]
    rc = lib().orc_gif_save(path.encode(), w, h, frames.ctypes.data, n, frame_delay)
    if rc != 0:
        raise OSError(f"orc_gif_save failed: {rc}")
