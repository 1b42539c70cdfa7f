"""Independent pins for the oracle (CPU only): float64 numpy restatements written from the published definitions, not from
oracle.cpp — Möller–Trumbore, the slab test, and the Unity transform conventions of SURVEY App. D (Translate / Scale /
AngleAxis composition order, Quaternion.Euler = Z then X then Y, TRS inverse).  They cannot prove the oracle equals a real
Unity run (parity stays unpinned) but they do catch a mis-transcribed formula shared by oracle and kernels."""
import ctypes as C

import numpy as np
import pytest

from util import abi, oracle_scene, params, scene_mod, synth


def _one_triangle_scene(v):
    s = scene_mod.ObjectData()
    s.Materials = [scene_mod.MaterialDescription((1, 1, 1), 0.1, 0.7, 0, 0, 1)]
    s.TriangleMeshes.append(scene_mod.TrianglesMesh(0, materials=np.zeros(1, np.int32), vertices=np.asarray(v, np.float32).reshape(1, 3, 3)))
    return s


def _mt64(o, d, v0, v1, v2):
    """Möller–Trumbore in float64 (Möller & Trumbore 1997), two-sided, returns t or None."""
    e1, e2 = v1 - v0, v2 - v0
    p = np.cross(d, e2)
    det = e1 @ p
    if abs(det) < 1e-4:
        return None
    tv = o - v0
    u = (tv @ p) / det
    q = np.cross(tv, e1)
    v = (d @ q) / det
    if u < 0 or u > 1 or v < 0 or u + v > 1:
        return None
    t = (e2 @ q) / det
    return t if t > 1e-4 else None


def test_moller_trumbore_against_float64(oracle):
    rng = np.random.RandomState(11)
    checked = hits = 0
    for trial in range(60):
        tri = (rng.rand(3, 3) - 0.5) * 8.0
        osc, holder = oracle_scene(oracle, _one_triangle_scene(tri))
        tri32 = tri.astype(np.float32).astype(np.float64)
        for k in range(60):
            o = ((rng.rand(3) - 0.5) * 20.0).astype(np.float32)
            target = tri32[0] + rng.rand() * (tri32[1] - tri32[0]) * 1.3 + rng.rand() * (tri32[2] - tri32[0]) * 1.3
            d = (target - o)
            d = (d / np.linalg.norm(d)).astype(np.float32)
            t32, ids, n = osc.brute_closest(o, d)
            ref = _mt64(o.astype(np.float64), d.astype(np.float64), *tri32)
            # barycentric / determinant margins: skip rays within rounding distance of an edge or of the det threshold
            e1, e2 = tri32[1] - tri32[0], tri32[2] - tri32[0]
            p = np.cross(d.astype(np.float64), e2)
            det = e1 @ p
            if abs(abs(det) - 1e-4) < 1e-3 * max(1.0, abs(det)):
                continue
            tv = o.astype(np.float64) - tri32[0]
            u = (tv @ p) / det
            v = (d.astype(np.float64) @ np.cross(tv, e1)) / det
            if min(abs(u), abs(u - 1), abs(v), abs(u + v - 1)) < 1e-3:
                continue
            checked += 1
            if ref is None:
                assert n == 0, (trial, k)
            else:
                hits += 1
                assert n == 1 and abs(t32 - ref) <= 2e-4 * max(1.0, ref), (trial, k, t32, ref)
    assert checked > 2000 and hits > 500


def _float64_matrix(elems):
    """Composite transform in float64 from the textbook forms: M = E0 * E1 * ... (column vectors)."""
    M = np.eye(4)
    for kind, x, y, z, a in elems:
        E = np.eye(4)
        r = np.deg2rad(a)
        c, s = np.cos(r), np.sin(r)
        if kind == abi.RTB_XF_T:
            E[:3, 3] = (x, y, z)
        elif kind == abi.RTB_XF_S:
            E[0, 0], E[1, 1], E[2, 2] = x, y, z
        elif kind == abi.RTB_XF_RX:
            E[1:3, 1:3] = [[c, -s], [s, c]]
        elif kind == abi.RTB_XF_RY:
            E[0, 0], E[0, 2], E[2, 0], E[2, 2] = c, s, -s, c
        elif kind == abi.RTB_XF_RZ:
            E[0:2, 0:2] = [[c, -s], [s, c]]
        M = M @ E
    return M


def test_composite_transform_and_camera_inverse_against_float64(pkg, oracle):
    lib = abi.load()
    rng = np.random.RandomState(5)
    T = scene_mod.TransformElement
    for trial in range(25):
        elems = []
        for _ in range(rng.randint(1, 6)):
            kind = int(rng.choice([abi.RTB_XF_T, abi.RTB_XF_S, abi.RTB_XF_RX, abi.RTB_XF_RY, abi.RTB_XF_RZ]))
            x, y, z = (rng.rand(3) * 4 + 0.5) if kind == abi.RTB_XF_S else (rng.rand(3) - 0.5) * 40
            elems.append((kind, float(x), float(y), float(z), float(rng.rand() * 360 - 180)))
        s = scene_mod.ObjectData()
        s.Transformations = [scene_mod.CompositeTransformation([T(k, (x, y, z), a) for k, x, y, z, a in elems])]
        s.Camera = scene_mod.CameraSettings(0, 30.0, 40.0)
        s.Lights = [scene_mod.LightSource(0, (1, 1, 1))]
        holder = scene_mod.pack_scene(s)
        out = np.zeros(25, np.float32)
        wh = (C.c_int32 * 2)()
        assert lib.rtb_resolve_frame(holder.ptr(), C.byref(params(64, 64, 2)), out.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
        M = _float64_matrix(elems)
        want = np.linalg.inv(M)
        got = out[:16].reshape(4, 4).astype(np.float64)
        scale = max(1.0, np.abs(want).max())
        assert np.abs(got - want).max() <= 2e-4 * scale, (trial, got, want)
        assert np.abs(out[19:22] - M[:3, 3]).max() <= 1e-3  # light position = translation column of its matrix
        # the oracle resolves the same uniforms bit for bit (its own code path)
        osc, h2 = oracle_scene(oracle, s)
        assert osc.frame(params(64, 64, 2)).tobytes() == out.tobytes()


def test_camera_override_euler_order_is_z_then_x_then_y(pkg):
    """Quaternion.Euler(x, y, z) rotates about Z, then X, then Y (Unity scripting reference); TRS = T * R (unit scale)."""
    lib = abi.load()
    holder = scene_mod.pack_scene(synth.sample_scene("test_scene_1"))
    rng = np.random.RandomState(9)
    for _ in range(20):
        rx, ry, rz = (rng.rand(3) * 360 - 180)
        pos = (rng.rand(3) - 0.5) * 100
        p = params(32, 32, 1, has_cam_pos=1, cam_pos=tuple(float(v) for v in pos), has_cam_rot=1, cam_rot_euler_deg=(float(rx), float(ry), float(rz)))
        out = np.zeros(25, np.float32)
        assert lib.rtb_resolve_frame(holder.ptr(), C.byref(p), out.ctypes.data_as(C.POINTER(C.c_float)), None) == abi.RTB_OK
        R = _float64_matrix([(abi.RTB_XF_RY, 0, 0, 0, ry), (abi.RTB_XF_RX, 0, 0, 0, rx), (abi.RTB_XF_RZ, 0, 0, 0, rz)])  # Y * X * Z
        TRS = np.eye(4)
        TRS[:3, :3] = R[:3, :3]
        TRS[:3, 3] = pos
        want = np.linalg.inv(TRS)
        assert np.abs(out[:16].reshape(4, 4) - want).max() <= 3e-4 * max(1.0, np.abs(want).max())


def test_slab_semantics_against_float64(oracle):
    """Rays aimed at / past a single axis-aligned box (as 12 triangles): the oracle's BVH query must agree with a float64 slab +
    triangle test about hit / miss for rays that are not within rounding distance of the silhouette."""
    s = scene_mod.ObjectData()
    s.Transformations = [scene_mod.CompositeTransformation([scene_mod.TransformElement.Translation((1.0, -2.0, 3.0)), scene_mod.TransformElement.Scale((4.0, 2.0, 6.0))])]
    s.Materials = [scene_mod.MaterialDescription((1, 1, 1), 0.1, 0.7, 0, 0, 1)]
    s.Boxes = [scene_mod.BoxDescription(0, 0)]
    osc, holder = oracle_scene(oracle, s)
    lo, hi = np.array([-1.0, -3.0, 0.0]), np.array([3.0, -1.0, 6.0])
    rng = np.random.RandomState(3)
    checked = 0
    for _ in range(3000):
        o = (rng.rand(3) - 0.5) * 30
        if np.all(o > lo - 0.05) and np.all(o < hi + 0.05):
            continue
        target = lo + rng.rand(3) * (hi - lo) * 1.6 - 0.3 * (hi - lo)
        d = target - o
        d /= np.linalg.norm(d)
        o32, d32 = o.astype(np.float32), d.astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            t0, t1 = (lo - o32) / d32.astype(np.float64), (hi - o32) / d32.astype(np.float64)
        tn, tf = np.minimum(t0, t1).max(), np.maximum(t0, t1).min()
        if abs(tn - tf) < 1e-3 or abs(tf) < 1e-3:
            continue
        t32, ids, n = osc.brute_closest(o32, d32)
        checked += 1
        hit = tn <= tf and tf > 0
        assert (n > 0) == hit, (o, d, tn, tf, n)
        if hit:
            assert abs(t32 - (tn if tn > 1e-4 else tf)) <= 1e-3 * max(1.0, tf)
    assert checked > 1500
