"""ctypes mirror of include/rtb.h (the C ABI of librtb200.so).

Field order and types must match the header exactly; tests/test_abi.py checks sizes against the values the library
reports through rtb_abi_sizes().  Nothing here computes anything.
"""
from __future__ import annotations

import ctypes as C
import os

RTB_OK = 0
RTB_E_ARG = -1
RTB_E_CUDA = -2
RTB_E_NOSCENE = -3
RTB_E_SIZE = -4
RTB_E_CANCELLED = -5
RTB_E_IO = -6
RTB_E_PARSE = -7

RTB_XF_T, RTB_XF_RX, RTB_XF_RY, RTB_XF_RZ, RTB_XF_S = 0, 1, 2, 3, 4
RTB_PRIM_TESSELLATED, RTB_PRIM_ANALYTIC = 0, 1
RTB_BVH_REFERENCE, RTB_BVH_LBVH = 0, 1
RTB_OUT_FRAME, RTB_OUT_COMPACT = 0, 1
RTB_EXT_OPAQUE_FD, RTB_EXT_OPAQUE_WIN32, RTB_EXT_D3D12_HEAP, RTB_EXT_D3D12_RESOURCE = 1, 2, 4, 5


class XformElem(C.Structure):
    _fields_ = [("type", C.c_int32), ("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("angle_deg", C.c_float)]


class Material(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("r", "g", "b", "ka", "kd", "ks", "kr", "ior")]


class Triangle(C.Structure):
    _fields_ = [("material", C.c_int32), ("v0", C.c_float * 3), ("v1", C.c_float * 3), ("v2", C.c_float * 3)]


class Mesh(C.Structure):
    _fields_ = [("xform", C.c_int32), ("reserved", C.c_int32), ("first_tri", C.c_int64), ("n_tris", C.c_int64)]


class Prim(C.Structure):
    _fields_ = [("xform", C.c_int32), ("material", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [
        ("has_image", C.c_int32), ("image_w", C.c_int32), ("image_h", C.c_int32), ("bg", C.c_float * 3),
        ("has_camera", C.c_int32), ("cam_xform", C.c_int32), ("cam_distance", C.c_float), ("cam_vfov_deg", C.c_float),
        ("n_xforms", C.c_int32), ("xform_offsets", C.POINTER(C.c_int32)), ("xform_elems", C.POINTER(XformElem)),
        ("n_lights", C.c_int32), ("light_xforms", C.POINTER(C.c_int32)), ("light_rgb", C.POINTER(C.c_float)),
        ("n_materials", C.c_int32), ("materials", C.POINTER(Material)),
        ("n_meshes", C.c_int32), ("meshes", C.POINTER(Mesh)),
        ("n_triangles", C.c_int64), ("triangles", C.POINTER(Triangle)),
        ("n_spheres", C.c_int32), ("spheres", C.POINTER(Prim)),
        ("n_boxes", C.c_int32), ("boxes", C.POINTER(Prim)),
    ]


class RenderParams(C.Structure):
    _fields_ = [
        ("has_resolution", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("has_bg", C.c_int32), ("bg", C.c_float * 3),
        ("light_intensity", C.c_float),
        ("has_cam_pos", C.c_int32), ("cam_pos", C.c_float * 3),
        ("has_cam_rot", C.c_int32), ("cam_rot_euler_deg", C.c_float * 3),
        ("has_fov", C.c_int32), ("fov_deg", C.c_float),
        ("max_depth", C.c_int32),
        ("enable_ambient", C.c_int32), ("enable_diffuse", C.c_int32), ("enable_specular", C.c_int32),
        ("enable_refraction", C.c_int32),
        ("is_orthographic", C.c_int32),
        ("aa_samples", C.c_int32),
        ("soft_shadows", C.c_int32), ("light_size", C.c_float),
        ("glossy", C.c_int32), ("roughness", C.c_float),
        ("motion_blur", C.c_int32), ("shutter_speed", C.c_float),
        ("debug_mode", C.c_int32), ("srgb_encode", C.c_int32),
        ("band_rank", C.c_int32), ("band_world", C.c_int32), ("band_rows", C.c_int32),
        ("out_layout", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("rays_primary", C.c_int64), ("rays_continuation", C.c_int64), ("rays_shadow", C.c_int64),
        ("paths_hit_primary", C.c_int64),
        ("n_triangles", C.c_int64), ("n_nodes", C.c_int64),
        ("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("chunks", C.c_int32),
        ("kernel_launches", C.c_int32), ("n_devices", C.c_int32),
        ("ms_upload", C.c_float), ("ms_build", C.c_float),
        ("ms_render_device", C.c_float),
        ("ms_traverse", C.c_float), ("ms_shade", C.c_float), ("ms_resolve", C.c_float),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("reserved", C.c_int64 * 4),
        ("rays_traversed", C.c_int64), ("packet_node_fetches", C.c_int64), ("packet_tri_fetches", C.c_int64), ("bytes_per_slot", C.c_int64),
    ]


def default_params() -> RenderParams:
    """RenderSettings as the reference UI builds them (SceneBuilder.cs:335-343,401,439-445)."""
    p = RenderParams()
    p.light_intensity = 1.0
    p.max_depth = 2
    p.enable_ambient = p.enable_diffuse = p.enable_specular = p.enable_refraction = 1
    p.aa_samples = 1
    return p


# Exported symbols of librtb200.so with (restype, argtypes); tests check that every one of them resolves.
_VP = C.c_void_p
SYMBOLS = {
    "rtb_create": (C.c_int, [C.POINTER(_VP), C.POINTER(C.c_int32), C.c_int32]),
    "rtb_destroy": (None, [_VP]),
    "rtb_params_default": (None, [C.POINTER(RenderParams)]),
    "rtb_upload_scene": (C.c_int, [_VP, C.POINTER(SceneDesc), C.c_int32, C.c_int32]),
    "rtb_invalidate": (C.c_int, [_VP]),
    "rtb_clear_target": (C.c_int, [_VP]),
    "rtb_render": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "rtb_render_begin": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, C.c_size_t, C.POINTER(C.c_int32)]),
    "rtb_render_end": (C.c_int, [_VP, C.c_int32]),
    "rtb_render_device": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, C.c_size_t, C.c_int32]),
    "rtb_render_aux": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _VP, _VP]),
    "rtb_get_triangles": (C.c_int, [_VP, _VP, _VP, C.c_int64, C.POINTER(C.c_int64)]),
    "rtb_get_stats": (C.c_int, [_VP, C.POINTER(Stats)]),
    "rtb_set_profiling": (C.c_int, [_VP, C.c_int32]),
    "rtb_set_cancel_flag": (C.c_int, [_VP, _VP]),
    "rtb_synchronize": (C.c_int, [_VP]),
    "rtb_get_stream": (_VP, [_VP, C.c_int32]),
    "rtb_flush": (C.c_int, [_VP]),
    "rtb_last_error": (C.c_char_p, [_VP]),
    "rtb_alloc_pinned": (_VP, [C.c_size_t]),
    "rtb_free_pinned": (None, [_VP]),
    "rtb_frame_export": (C.c_int, [_VP, C.c_size_t, C.POINTER(_VP), C.POINTER(C.c_uint8)]),
    "rtb_frame_import": (C.c_int, [_VP, C.POINTER(C.c_uint8), C.POINTER(_VP)]),
    "rtb_scene_load": (C.c_int, [C.c_char_p, C.POINTER(_VP), C.c_char_p, C.c_size_t]),
    "rtb_scene_parse": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(_VP), C.c_char_p, C.c_size_t]),
    "rtb_scene_get": (C.POINTER(SceneDesc), [_VP]),
    "rtb_scene_free": (None, [_VP]),
    "rtb_api_version": (C.c_int, []),
    "rtb_abi_sizes": (None, [C.POINTER(C.c_int32), C.c_int32]),
    "rtb_resolve_frame": (C.c_int, [C.POINTER(SceneDesc), C.POINTER(RenderParams), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "rtb_get_bvh": (C.c_int, [_VP, _VP, C.c_int64, C.POINTER(C.c_int64), _VP, C.c_int64]),
    "rtb_get_bvh_node_words": (C.c_int, [_VP]),
    "rtb_group_create": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_size_t, C.c_int32, _VP]),
    "rtb_group_render_begin": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, C.c_size_t, C.POINTER(C.c_int32)]),
    "rtb_group_render_end": (C.c_int, [_VP, C.c_int32]),
    "rtb_group_destroy": (C.c_int, [_VP]),
    "rtb_group_create_host": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_size_t, C.c_int32, C.c_char_p]),
    "rtb_group_frame": (C.c_int, [_VP, C.c_int32, C.POINTER(C.c_void_p)]),
    "rtb_external_import": (C.c_int, [_VP, C.c_int32, _VP, C.c_size_t, C.c_int32, C.POINTER(_VP)]),
    "rtb_external_release": (C.c_int, [_VP, _VP]),
    "rtb_frame_read": (C.c_int, [_VP, _VP, C.c_size_t]),
    "rtb_build_reference_bvh": (C.c_int, [_VP, C.c_int32, _VP, C.c_int64, C.POINTER(C.c_int64), _VP]),
    # GIF sweep (GifGenerator.cs)
    "rtb_gif_color_table": (None, [_VP]),
    "rtb_gif_lzw_bound": (C.c_int64, [C.c_int64]),
    "rtb_gif_lzw": (C.c_int64, [_VP, C.c_int64, _VP, C.c_int64]),
    "rtb_gif_index_frame": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP]),
    "rtb_gif_index_device": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP]),
    "rtb_render_begin_indexed": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, C.c_size_t, C.POINTER(C.c_int32)]),
    "rtb_gif_save_indexed": (C.c_int, [C.c_char_p, C.c_int32, C.c_int32, C.POINTER(_VP), C.c_int32, C.c_int32, C.c_int32]),
    "rtb_gif_save": (C.c_int, [_VP, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(_VP), C.c_int32, C.c_int32, C.c_int32]),
    "rtb_gif_render_rotation": (C.c_int, [_VP, C.POINTER(RenderParams), C.c_int32, C.c_float, C.c_char_p, C.c_int32, C.c_int32]),
}

LIB_NAME = "librtb200.so"


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


_lib = None


def load():
    """Loads the CUDA library.  There is no CPU fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            f"{LIB_NAME} is not built ({path}); run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C cosig-raytracing_b200/csrc`.  There is no CPU fallback.")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
