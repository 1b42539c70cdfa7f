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


# ---- a whole frame: float64, vectorised, brute force (no BVH) --------------------------------------------------------------------
def _normalize64(v):
    return v / np.sqrt((v * v).sum(-1, keepdims=True))


def _closest64(o, d, v0, e1, e2):
    """Closest Möller–Trumbore hit of every ray against every triangle (BVHRayTracing.compute:153-190: |det| >= 1e-4, inclusive
    u / v bounds, t > 1e-4).  Returns (t, index, u, v) with index -1 for a miss."""
    p = np.cross(d[:, None, :], e2[None, :, :])
    det = (e1[None] * p).sum(-1)
    ok = np.abs(det) >= 1e-4
    inv = 1.0 / np.where(ok, det, 1.0)
    tv = o[:, None, :] - v0[None]
    u = (tv * p).sum(-1) * inv
    q = np.cross(tv, e1[None])
    v = (d[:, None, :] * q).sum(-1) * inv
    t = (e2[None] * q).sum(-1) * inv
    ok &= (u >= 0) & (u <= 1) & (v >= 0) & (u + v <= 1) & (t > 1e-4)
    t = np.where(ok, t, np.inf)
    idx = t.argmin(1)
    r = np.arange(len(o))
    best = t[r, idx]
    return best, np.where(np.isfinite(best), idx, -1), u[r, idx], v[r, idx]


def _analytic64(kind, M, o, d):
    """SphereInstance.Hit / BoxInstance.Hit (HittableObjects.cs:45-77, 149-174) in float64: ray to object space with the direction
    re-normalised, unit sphere (quadratic, :82-107) or unit cube (slabs with face tracking, :180-223), world t = |pWS - origin|,
    normal = normalize(worldToObject^T nOS).  Returns (tWS or inf, pWS, nWS)."""
    W = np.linalg.inv(M)
    oo = o @ W[:3, :3].T + W[:3, 3]
    dd = _normalize64(d @ W[:3, :3].T)
    n = len(o)
    if kind == 1:
        a = (dd * dd).sum(-1)
        b = 2.0 * (oo * dd).sum(-1)
        c = (oo * oo).sum(-1) - 1.0
        disc = b * b - 4.0 * a * c
        ok = disc >= 0
        sq = np.sqrt(np.where(ok, disc, 0.0))
        t0, t1 = (-b - sq) / (2.0 * a), (-b + sq) / (2.0 * a)
        tos = np.where(t0 > 1e-3, t0, t1)
        ok &= tos > 1e-3
        pos_os = oo + tos[:, None] * dd
        n_os = _normalize64(np.where(ok[:, None], pos_os, 1.0))
    else:
        tmin, tmax = np.full(n, -1e20), np.full(n, 1e20)
        nmin, nmax = np.zeros((n, 3)), np.zeros((n, 3))
        ok = np.ones(n, bool)
        for ax in range(3):
            with np.errstate(divide="ignore", invalid="ignore"):
                inv = np.where(np.abs(dd[:, ax]) > 1e-8, 1.0 / dd[:, ax], np.inf)
                t1, t2 = (-0.5 - oo[:, ax]) * inv, (0.5 - oo[:, ax]) * inv
            e1_, e2_ = np.zeros(3), np.zeros(3)
            e1_[ax], e2_[ax] = -1.0, 1.0
            swap = t1 > t2
            ta, tb = np.where(swap, t2, t1), np.where(swap, t1, t2)
            na = np.where(swap[:, None], e2_, e1_)
            nb = np.where(swap[:, None], e1_, e2_)
            up = ta > tmin
            tmin, nmin = np.where(up, ta, tmin), np.where(up[:, None], na, nmin)
            dn = tb < tmax
            tmax, nmax = np.where(dn, tb, tmax), np.where(dn[:, None], nb, nmax)
            ok &= ~(tmin > tmax) & ~(tmax < 1e-3)
        tos = np.where(tmin >= 1e-3, tmin, tmax)
        ok &= tos >= 1e-3
        n_os = np.where((tos == tmin)[:, None], nmin, nmax)
        pos_os = oo + tos[:, None] * dd
    pos_ws = pos_os @ M[:3, :3].T + M[:3, 3]
    tws = np.sqrt(((pos_ws - o) ** 2).sum(-1))
    ok &= tws > 1e-4
    n_ws = _normalize64(np.where(ok[:, None], n_os @ W[:3, :3], 1.0))  # W^T n
    return np.where(ok, tws, np.inf), pos_ws, n_ws


def _scene64(o, d, v0, e1, e2, n0, n1, n2, mat_idx, prims):
    """Closest hit over triangles, then analytic primitives in emission order (strict "<": the first of equal t wins).
    Returns (t or inf, position, shading normal, material index)."""
    n = len(o)
    if len(v0):
        t, idx, bu, bv = _closest64(o, d, v0, e1, e2)
        safe = np.maximum(idx, 0)
        pos = o + np.where(np.isfinite(t), t, 0.0)[:, None] * d
        nrm = _normalize64(np.where((idx >= 0)[:, None], (1 - bu - bv)[:, None] * n0[safe] + bu[:, None] * n1[safe] + bv[:, None] * n2[safe], 1.0))
        mi = mat_idx[safe].copy()
    else:
        t, pos, nrm, mi = np.full(n, np.inf), np.zeros((n, 3)), np.ones((n, 3)), np.zeros(n, np.int64)
    for kind, M, material in prims:
        ta, pa, na = _analytic64(kind, M, o, d)
        better = ta < t
        t = np.where(better, ta, t)
        pos, nrm = np.where(better[:, None], pa, pos), np.where(better[:, None], na, nrm)
        mi = np.where(better, material, mi)
    return t, pos, nrm, mi


def _render64(tri18, mat_idx, materials, u25, w, h, max_depth, prims=()):
    """The per-pixel loop of CSMain (BVHRayTracing.compute:283-340 ray generation, :356-478 depth loop) in float64 numpy, written
    from the shader text: one sample per pixel, perspective camera, all lighting toggles on, no distribution effects.
    `prims`: analytic spheres / boxes (kind, objectToWorld 4x4, material) tested after the triangles, as the oracle's analytic mode
    does with the semantics of HittableObjects.cs."""
    M = u25[:16].reshape(4, 4).astype(np.float64)
    cam_d, tan_half = float(u25[16]), float(u25[17])
    light, bg = u25[19:22].astype(np.float64), u25[22:25].astype(np.float64)
    t64 = tri18.astype(np.float64)
    v0, e1, e2 = t64[:, 0:3], t64[:, 3:6] - t64[:, 0:3], t64[:, 6:9] - t64[:, 0:3]
    n0, n1, n2 = t64[:, 9:12], t64[:, 12:15], t64[:, 15:18]
    ys, xs = np.mgrid[0:h, 0:w]
    plane_h = 2.0 * cam_d * tan_half
    plane_w = plane_h * (w / h)
    uu = ((xs.ravel() + 0.5) / w - 0.5) * plane_w
    vv = ((ys.ravel() + 0.5) / h - 0.5) * plane_h
    oc = np.array([0.0, 0.0, cam_d])
    dc = _normalize64(np.stack([uu, vv, np.zeros_like(uu)], -1) - oc)
    o = np.broadcast_to(M[:3, :3] @ oc + M[:3, 3], dc.shape).copy()
    d = _normalize64(dc @ M[:3, :3].T)
    n = w * h
    color = np.zeros((n, 3))
    att = np.ones((n, 3))
    alive = np.arange(n)
    mats = np.array([[*m.color, m.ambient, m.diffuse, m.specular, m.refraction, m.ior] for m in materials], np.float64)
    for _ in range(max_depth):
        if len(alive) == 0:
            break
        t, pos, nrm, mi = _scene64(o, d, v0, e1, e2, n0, n1, n2, mat_idx, prims)
        miss = ~np.isfinite(t)
        color[alive[miss]] += att[miss] * bg
        keep = ~miss
        alive, o, d, att, t, pos, nrm, mi = alive[keep], o[keep], d[keep], att[keep], t[keep], pos[keep], nrm[keep], mi[keep]
        if len(alive) == 0:
            break
        m = mats[mi]
        col, ka, kd, ks, kr, ior = m[:, :3], m[:, 3], m[:, 4], m[:, 5], m[:, 6], m[:, 7]
        local = col * ka[:, None]
        ldir = _normalize64(light - pos)
        ndl = np.maximum(0.0, (nrm * ldir).sum(-1))
        dist = np.sqrt(((light - pos) ** 2).sum(-1))
        need = ndl > 0
        lit = np.zeros(len(alive), bool)
        if need.any():
            st, _, _, _ = _scene64((pos + nrm * 1e-2)[need], ldir[need], v0, e1, e2, n0, n1, n2, mat_idx, prims)
            lit[need] = ~np.isfinite(st) | (st > dist[need])
        half = _normalize64(ldir + _normalize64(-d))
        spec = np.maximum((nrm * half).sum(-1), 0.0) ** 32
        local = local + np.where(lit[:, None], col * (kd * ndl)[:, None] + np.where(ks > 0, ks * spec, 0.0)[:, None], 0.0)
        color[alive] += att * local
        reflect, refract = ks > 0, kr > 0
        go = reflect | refract
        I = _normalize64(d)
        N = nrm.copy()
        eta = 1.0 / ior
        flip = refract & ((I * N).sum(-1) > 0)
        N[flip] = -N[flip]
        eta = np.where(flip, ior, eta)
        cosi = (-I * N).sum(-1)
        k = 1.0 - eta * eta * (1.0 - cosi * cosi)
        through = refract & (k >= 0)
        tir = refract & (k < 0)
        mirror = reflect & ~refract
        nd = np.zeros_like(d)
        start = pos.copy()
        rdir = eta[:, None] * I + (eta * cosi - np.sqrt(np.maximum(k, 0.0)))[:, None] * N
        nd[through] = rdir[through]
        start[through] += rdir[through] * 1e-2
        rtir = I - 2.0 * (N * I).sum(-1, keepdims=True) * N
        nd[tir] = rtir[tir]
        start[tir] += N[tir] * 1e-2
        rmir = I - 2.0 * (nrm * I).sum(-1, keepdims=True) * nrm
        nd[mirror] = rmir[mirror]
        start[mirror] += nrm[mirror] * 1e-2
        att = att * np.where(through[:, None], col * kr[:, None], np.where((tir | mirror)[:, None], col * ks[:, None], 1.0))
        alive, o, d, att = alive[go], start[go], _normalize64(nd[go]), att[go]
    img = np.floor(np.clip(color, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8).reshape(h, w, 3)
    return img


@pytest.mark.parametrize("w,h,depth", [(64, 48, 3), (96, 72, 6)])
@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_whole_frame_against_float64_brute_force(pkg, oracle, name, w, h, depth):
    """The oracle's frame of each shipped scene (FP32, BVH traversal, its own bookkeeping) against the float64 brute-force
    renderer above, which shares no code with it.  Tolerance: RGB within 1/255 on >= 99 % of the pixels — the two differ where
    FP32 rounding moves a silhouette, a shadow edge or a tie between two triangles of equal t across a pixel centre.  (Measured
    when the test was written: all six frames identical, byte for byte.)"""
    obj = synth.sample_scene(name)
    osc, holder = oracle_scene(oracle, obj)
    p = params(w, h, depth)
    ref = osc.render(p)["rgba8"][..., :3]
    u25 = np.zeros(25, np.float32)
    wh = (C.c_int32 * 2)()
    assert abi.load().rtb_resolve_frame(holder.ptr(), C.byref(p), u25.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
    tri18, mat_idx, _ = osc.triangles()
    got = _render64(tri18, mat_idx, obj.Materials, u25, w, h, depth)
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32)).max(-1)
    within = float((diff <= 1).mean())
    assert within >= 0.99, f"{name}: only {within * 100:.2f}% of pixels within 1/255 (worst {int(diff.max())})"
    assert len(np.unique(ref.reshape(-1, 3), axis=0)) > 20  # a real picture, not a constant frame


# ---- anti-aliasing samples: Hash22 jitter and ray generation in FP32, written from the shader text --------------------------------
def _f(x):
    return np.float32(x)


def _frac32(x):
    return _f(x - np.floor(x))


def _hash22_f32(px, py):
    """Hash22, BVHRayTracing.compute:108-113, every operation individually rounded to FP32 (dot = (a.x b.x + a.y b.y) + a.z b.z)."""
    a, b, c = _frac32(_f(px) * _f(.1031)), _frac32(_f(py) * _f(.1030)), _frac32(_f(px) * _f(.0973))
    k = _f(33.33)
    d = _f(_f(_f(a * _f(b + k)) + _f(b * _f(c + k))) + _f(c * _f(a + k)))  # dot(p3, p3.yzx + 33.33)
    a, b, c = _f(a + d), _f(b + d), _f(c + d)
    return _frac32(_f(_f(a + b) * c)), _frac32(_f(_f(a + c) * b))             # frac((p3.xx + p3.yz) * p3.zy)


def _normalize_f32(v):
    d = _f(_f(_f(v[0] * v[0]) + _f(v[1] * v[1])) + _f(v[2] * v[2]))
    r = _f(_f(1.0) / np.sqrt(d, dtype=np.float32))
    return np.array([_f(v[0] * r), _f(v[1] * r), _f(v[2] * r)], np.float32)


def _sample_ray_f32(u25, w, h, n_samples, px, py, i, ortho=False):
    """compute:283-340 for AA sample i of pixel (px, py); perspective camera, or the orthographic branch (:318-327)."""
    M = u25[:16].reshape(4, 4)
    cam_d, tan_half = _f(u25[16]), _f(u25[17])
    grid_w = int(np.ceil(np.sqrt(np.float32(n_samples))))
    grid_h = int(np.ceil(np.float32(n_samples) / np.float32(grid_w)))
    aspect = _f(_f(w) / _f(h))
    plane_h = _f(_f(2.0) * _f(cam_d * tan_half))
    plane_w = _f(plane_h * aspect)
    ox, oy = _f(0.5), _f(0.5)
    if n_samples > 1:
        gy, gx = i // grid_w, i % grid_w
        jx, jy = _hash22_f32(_f(_f(px) + _f(_f(i) * _f(13.0))), _f(_f(py) + _f(_f(i) * _f(7.0))))
        ox = _f(_f(_f(gx) + jx) / _f(grid_w))
        oy = _f(_f(_f(gy) + jy) / _f(grid_h))
    u = _f(_f(_f(_f(_f(px) + ox) / _f(w)) - _f(0.5)) * plane_w)
    v = _f(_f(_f(_f(_f(py) + oy) / _f(h)) - _f(0.5)) * plane_h)
    if ortho:
        half_h = _f(u25[18])
        half_w = _f(half_h * aspect)
        ou = _f(_f(_f(_f(_f(_f(px) + ox) / _f(w)) - _f(0.5)) * _f(2.0)) * half_w)
        ov = _f(_f(_f(_f(_f(_f(py) + oy) / _f(h)) - _f(0.5)) * _f(2.0)) * half_h)
        oc = np.array([ou, ov, cam_d], np.float32)
        dc = np.array([0.0, 0.0, -1.0], np.float32)
    else:
        dc = _normalize_f32(np.array([_f(u - _f(0.0)), _f(v - _f(0.0)), _f(_f(0.0) - cam_d)], np.float32))
        oc = np.array([0.0, 0.0, cam_d], np.float32)
    o = np.array([_f(_f(_f(_f(M[r, 0] * oc[0]) + _f(M[r, 1] * oc[1])) + _f(M[r, 2] * oc[2])) + _f(M[r, 3] * _f(1.0))) for r in range(3)], np.float32)
    d = _normalize_f32(np.array([_f(_f(_f(M[r, 0] * dc[0]) + _f(M[r, 1] * dc[1])) + _f(M[r, 2] * dc[2])) for r in range(3)], np.float32))
    return o, d


@pytest.mark.parametrize("n_samples", [1, 4, 9, 16, 5])
def test_aa_sample_rays_bit_exact_against_fp32_numpy(pkg, oracle, n_samples):
    """Stratified grid + Hash22 jitter + camera transform of every AA sample (compute:283-340), restated in numpy FP32 from the
    shader text with the rounding conventions of SURVEY App. D: origin and direction bits must equal the oracle's."""
    obj = synth.sample_scene("test_scene_1")
    osc, holder = oracle_scene(oracle, obj)
    w, h = 200, 150
    p = params(w, h, 1, n_samples)
    u25 = np.zeros(25, np.float32)
    wh = (C.c_int32 * 2)()
    assert abi.load().rtb_resolve_frame(holder.ptr(), C.byref(p), u25.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
    rng = np.random.RandomState(n_samples)
    with np.errstate(over="ignore"):
        for _ in range(60):
            px, py, i = int(rng.randint(0, w)), int(rng.randint(0, h)), int(rng.randint(0, n_samples))
            o, d = osc.sample_ray(p, px, py, i)
            o2, d2 = _sample_ray_f32(u25, w, h, n_samples, px, py, i)
            assert o.tobytes() == o2.tobytes(), (px, py, i, o, o2)
            assert d.tobytes() == d2.tobytes(), (px, py, i, d, d2)


def test_random_unit_vector_hash_bits_and_sincos_accuracy(oracle):
    """RandomUnitVector (compute:116-131), the source of the soft-shadow / glossy / motion-blur jitter.  z = 2 h.z - 1 must equal a
    numpy FP32 restatement of Hash33 bit for bit; x, y use this build's fixed FP32 polynomial for cos / sin (shared by oracle and
    kernels so that both agree exactly) and must stay within 5e-7 of r cos(a), r sin(a) evaluated in float64."""
    rng = np.random.RandomState(3)
    worst = 0.0
    for _ in range(300):
        seed = np.array([rng.randint(0, 4000) + rng.randint(0, 16) * 9.0, rng.randint(0, 3000) + rng.randint(0, 16) * 4.0, rng.randint(0, 40)], np.float32)
        got = oracle.random_unit_vector(seed)
        p = [_frac32(_f(seed[0]) * _f(.1031)), _frac32(_f(seed[1]) * _f(.1030)), _frac32(_f(seed[2]) * _f(.0973))]
        k = _f(33.33)
        d = _f(_f(_f(p[0] * _f(p[1] + k)) + _f(p[1] * _f(p[0] + k))) + _f(p[2] * _f(p[2] + k)))  # dot(p, p.yxz + 33.33)
        p = [_f(p[0] + d), _f(p[1] + d), _f(p[2] + d)]
        hx = _frac32(_f(_f(p[0] + p[1]) * p[2]))  # frac((p.xxy + p.yxx) * p.zyx)
        hz = _frac32(_f(_f(p[1] + p[0]) * p[0]))
        z = _f(_f(hz * _f(2.0)) - _f(1.0))
        assert got[2].tobytes() == z.tobytes()
        a = _f(hx * _f(6.2831853))
        r = np.sqrt(_f(_f(1.0) - _f(z * z)), dtype=np.float32)
        worst = max(worst, abs(float(got[0]) - float(r) * np.cos(float(a))), abs(float(got[1]) - float(r) * np.sin(float(a))))
    assert worst <= 5e-7, worst


# ---- the kernel's debug views (compute:484-508), one of which the Unity dump uses to expose the primary t ----------------------------------
@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_debug_views_against_float64(pkg, oracle, name):
    """_DebugMode 1 (depth: t / 100, red on a miss), 2 (normal * 0.5 + 0.5, blue on a miss), 3 (green hit mask on grey): the pixel-centre
    perspective ray re-traced after the sample loop, restated in float64 from the shader text.  csharp/Editor/DumpGoldens.cs dumps view 1
    of the real reference, so this is the view a future pin of the primary hit distance goes through.  Also with the orthographic switch
    on: the shader's debug block ignores it (the re-trace is always perspective)."""
    obj = synth.sample_scene(name)
    osc, holder = oracle_scene(oracle, obj)
    w, h = 80, 60
    tri18, mat_idx, _ = osc.triangles()
    t64 = tri18.astype(np.float64)
    v0, e1, e2 = t64[:, 0:3], t64[:, 3:6] - t64[:, 0:3], t64[:, 6:9] - t64[:, 0:3]
    n0, n1, n2 = t64[:, 9:12], t64[:, 12:15], t64[:, 15:18]
    for ortho in (0, 1):
        p0 = params(w, h, 2, is_orthographic=ortho)
        u25 = np.zeros(25, np.float32)
        wh = (C.c_int32 * 2)()
        assert abi.load().rtb_resolve_frame(holder.ptr(), C.byref(p0), u25.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
        M = u25[:16].reshape(4, 4).astype(np.float64)
        cam_d, tan_half = float(u25[16]), float(u25[17])
        ys, xs = np.mgrid[0:h, 0:w]
        plane_h = 2.0 * cam_d * tan_half
        plane_w = plane_h * (w / h)
        uu = ((xs.ravel() + 0.5) / w - 0.5) * plane_w
        vv = ((ys.ravel() + 0.5) / h - 0.5) * plane_h
        oc = np.array([0.0, 0.0, cam_d])
        dc = _normalize64(np.stack([uu, vv, np.zeros_like(uu)], -1) - oc)
        o = np.broadcast_to(M[:3, :3] @ oc + M[:3, 3], dc.shape).copy()
        d = _normalize64(dc @ M[:3, :3].T)
        t, pos, nrm, mi = _scene64(o, d, v0, e1, e2, n0, n1, n2, mat_idx, ())
        hit = np.isfinite(t)
        want = {
            1: np.where(hit[:, None], np.repeat((np.where(hit, t, 0.0) / 100.0)[:, None], 3, 1), np.array([1.0, 0.0, 0.0])),
            2: np.where(hit[:, None], np.nan_to_num(nrm) * 0.5 + 0.5, np.array([0.0, 0.0, 1.0])),
            3: np.where(hit[:, None], np.array([0.0, 1.0, 0.0]), np.array([0.2, 0.2, 0.2])),
        }
        for mode in (1, 2, 3):
            ref = osc.render(params(w, h, 2, is_orthographic=ortho, debug_mode=mode))["rgba8"][..., :3]
            got = np.floor(np.clip(want[mode], 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8).reshape(h, w, 3)
            diff = np.abs(got.astype(np.int32) - ref.astype(np.int32)).max(-1)
            assert (diff <= 1).mean() >= 0.995, f"{name} debug view {mode} ortho {ortho}: {(diff <= 1).mean() * 100:.2f}% within 1/255, worst {int(diff.max())}"
        assert 0.05 < hit.mean() < 1.0


# ---- distribution effects: soft shadows, glossy reflections, motion blur — seeds and placement restated from the shader text -----------
def _ruv64(sx, sy, sz):
    """RandomUnitVector (compute:116-131): Hash33 in FP32 exactly as the test above restates it, then z, a, r and cos / sin in float64."""
    p = [_frac32(_f(sx) * _f(.1031)), _frac32(_f(sy) * _f(.1030)), _frac32(_f(sz) * _f(.0973))]
    k = _f(33.33)
    d = _f(_f(_f(p[0] * _f(p[1] + k)) + _f(p[1] * _f(p[0] + k))) + _f(p[2] * _f(p[2] + k)))
    p = [_f(p[0] + d), _f(p[1] + d), _f(p[2] + d)]
    hx = float(_frac32(_f(_f(p[0] + p[1]) * p[2])))
    hz = float(_frac32(_f(_f(p[1] + p[0]) * p[0])))
    z = hz * 2.0 - 1.0
    a = hx * 6.2831853
    r = np.sqrt(max(0.0, 1.0 - z * z))
    return np.array([r * np.cos(a), r * np.sin(a), z])


def _render64_fx(tri18, mat_idx, materials, u25, w, h, max_depth, n_samples, light_size=0.0, roughness=0.0, shutter=0.0, intensity=1.0,
                 ortho=False, ambient=True, diffuse=True, specular=True, refraction=True):
    """CSMain with its distribution effects (compute:283-478) in float64, one path at a time, written from the shader text: stratified
    AA samples (the FP32 restatement above, bit-exact), motion blur `origin += (RUV(x + i, y, i) - 0.5) * 0.2 * shutter` (:342-349), soft
    shadows `lightPos += RUV(x + 9 i, y + 4 i + depth, i) * lightSize` (:383-388), glossy `dir = normalize(dir + RUV(x + 55 i + depth,
    y + 22 i, 13 depth) * roughness)` (:459-470).  light_size / roughness / shutter = 0 switches the effect off."""
    light0, bg = u25[19:22].astype(np.float64), u25[22:25].astype(np.float64)
    t64 = tri18.astype(np.float64)
    v0, e1, e2 = t64[:, 0:3], t64[:, 3:6] - t64[:, 0:3], t64[:, 6:9] - t64[:, 0:3]
    n0, n1, n2 = t64[:, 9:12], t64[:, 12:15], t64[:, 15:18]
    mats = np.array([[*m.color, m.ambient, m.diffuse, m.specular, m.refraction, m.ior] for m in materials], np.float64)
    img = np.zeros((h, w, 3))

    def query(o, d):
        t, pos, nrm, mi = _scene64(o[None], d[None], v0, e1, e2, n0, n1, n2, mat_idx, ())
        return float(t[0]), pos[0], nrm[0], int(mi[0])

    with np.errstate(over="ignore"):
        for py in range(h):
            for px in range(w):
                acc = np.zeros(3)
                for i in range(n_samples):
                    o32, d32 = _sample_ray_f32(u25, w, h, n_samples, px, py, i, ortho)
                    o, d = o32.astype(np.float64), d32.astype(np.float64)
                    if shutter > 0.0:
                        o = o + (_ruv64(px + i, py, i) - 0.5) * 0.2 * shutter
                    col, att = np.zeros(3), np.ones(3)
                    for depth in range(max_depth):
                        t, pos, nrm, mi = query(o, d)
                        if not np.isfinite(t):
                            col += att * bg
                            break
                        c, ka, kd, ks, kr, ior = (mats[mi][:3], *mats[mi][3:]) if mi >= 0 else (np.ones(3), 0.1, 0.7, 0.0, 0.0, 1.0)
                        local = c * ka if ambient else np.zeros(3)
                        light = light0 + (_ruv64(px + i * 9.0, py + i * 4.0 + depth, i) * light_size if light_size > 0.0 else 0.0)
                        ldir = (light - pos) / np.linalg.norm(light - pos)
                        ndl = max(0.0, float(nrm @ ldir))
                        if diffuse and ndl > 0.0:
                            st, _, _, _ = query(pos + nrm * 1e-2, ldir)
                            if not np.isfinite(st) or st > np.linalg.norm(light - pos):
                                local = local + c * kd * ndl
                                if specular and ks > 0.0:
                                    hv = ldir + (-d) / np.linalg.norm(d)
                                    hv /= np.linalg.norm(hv)
                                    local = local + ks * max(float(nrm @ hv), 0.0) ** 32
                        col += att * local * intensity
                        reflect, refract = ks > 0.0, refraction and kr > 0.0
                        if not reflect and not refract:
                            break
                        I = d / np.linalg.norm(d)
                        start = pos.copy()
                        if refract:
                            N, eta = nrm.copy(), 1.0 / ior
                            if I @ N > 0:
                                N, eta = -N, ior
                            cosi = float(-I @ N)
                            k = 1.0 - eta * eta * (1.0 - cosi * cosi)
                            if k >= 0.0:
                                nd = eta * I + (eta * cosi - np.sqrt(k)) * N
                                att = att * c * kr
                                start = start + nd * 1e-2
                            else:
                                nd = I - 2.0 * float(N @ I) * N
                                att = att * c * ks
                                start = start + N * 1e-2
                        else:
                            nd = I - 2.0 * float(nrm @ I) * nrm
                            att = att * c * ks
                            start = start + nrm * 1e-2
                        if roughness > 0.0:
                            nd = nd + _ruv64(px + i * 55.0 + depth, py + i * 22.0, depth * 13) * roughness
                            nd = nd / np.linalg.norm(nd)
                        o, d = start, nd / np.linalg.norm(nd)
                    acc += col
                img[py, px] = acc / n_samples
    return np.floor(np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


@pytest.mark.parametrize("kw,fx", [
    (dict(soft_shadows=1, light_size=5.0), dict(light_size=5.0)),
    (dict(glossy=1, roughness=0.05), dict(roughness=0.05)),
    (dict(motion_blur=1, shutter_speed=1.0), dict(shutter=1.0)),
    (dict(soft_shadows=1, light_size=2.0, glossy=1, roughness=0.1, motion_blur=1, shutter_speed=0.5, light_intensity=1.3),
     dict(light_size=2.0, roughness=0.1, shutter=0.5, intensity=1.3)),
])
def test_distribution_effects_against_float64(pkg, oracle, kw, fx):
    """soft shadows, glossy reflections and motion blur of the oracle (FP32, shared polynomial sin / cos) against a
    float64 per-path restatement of the shader's loop with its own RandomUnitVector (FP32 hash bits, libm cos / sin): which seed goes
    where, at which depth and sample, and what is jittered before what.  4 samples per pixel, depth 4, the sample scene with its
    mirrors and its glass.  Tolerance: RGB within 1/255 on >= 98 % of the pixels, mean absolute difference below 0.2 of one 8-bit
    step — a jittered ray that lands on the other side of an edge in FP32 moves one sample of four.  (Measured when the test was written:
    all four frames identical, byte for byte.)"""
    obj = synth.sample_scene("test_scene_1")
    osc, holder = oracle_scene(oracle, obj)
    w, h, depth, aa = 36, 26, 4, 4
    p = params(w, h, depth, aa, **kw)
    ref = osc.render(p)["rgba8"][..., :3]
    u25 = np.zeros(25, np.float32)
    wh = (C.c_int32 * 2)()
    assert abi.load().rtb_resolve_frame(holder.ptr(), C.byref(p), u25.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
    tri18, mat_idx, _ = osc.triangles()
    got = _render64_fx(tri18, mat_idx, obj.Materials, u25, w, h, depth, aa, **fx)
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    within = float((diff.max(-1) <= 1).mean())
    assert within >= 0.98 and diff.mean() < 0.2, f"{kw}: {within * 100:.2f}% of pixels within 1/255, mean abs diff {diff.mean():.3f}, worst {int(diff.max())}"
    # and the effect is really on: the frame differs from the one without it
    plain = osc.render(params(w, h, depth, aa))["rgba8"][..., :3]
    assert (plain != ref).any(axis=-1).mean() > 0.02


@pytest.mark.parametrize("kw,fx", [
    (dict(is_orthographic=1), dict(ortho=True)),
    (dict(is_orthographic=1, aa_samples=4), dict(ortho=True)),
    (dict(enable_ambient=0), dict(ambient=False)),
    (dict(enable_diffuse=0), dict(diffuse=False)),
    (dict(enable_specular=0), dict(specular=False)),
    (dict(enable_refraction=0), dict(refraction=False)),
    (dict(light_intensity=1.7), dict(intensity=1.7)),
    (dict(has_bg=1, bg=(0.9, 0.1, 0.4), has_fov=1, fov_deg=55.0), dict()),
    (dict(has_cam_pos=1, cam_pos=(5.0, -60.0, 30.0), has_cam_rot=1, cam_rot_euler_deg=(-60.0, 10.0, 5.0)), dict()),
])
def test_render_settings_against_float64(pkg, oracle, kw, fx):
    """Every switch of RenderSettings that changes the picture — orthographic camera (compute:318-327), the four lighting toggles
    (:381, :392, :407, :423), light intensity (:418), background / fov / camera overrides (uniforms resolved by the library's host code) —
    through the float64 per-path restatement.  The specular toggle only gates the highlight (:407): a mirror still reflects with it off
    (:421).  Tolerance as for the whole-frame check: 99 % of the pixels within 1/255."""
    obj = synth.sample_scene("test_scene_1")
    osc, holder = oracle_scene(oracle, obj)
    w, h, depth = 44, 32, 4
    aa = kw.get("aa_samples", 1)
    p = params(w, h, depth, aa, **{k: v for k, v in kw.items() if k != "aa_samples"})
    ref = osc.render(p)["rgba8"][..., :3]
    u25 = np.zeros(25, np.float32)
    wh = (C.c_int32 * 2)()
    assert abi.load().rtb_resolve_frame(holder.ptr(), C.byref(p), u25.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
    tri18, mat_idx, _ = osc.triangles()
    got = _render64_fx(tri18, mat_idx, obj.Materials, u25, w, h, depth, aa, **fx)
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32)).max(-1)
    assert (diff <= 1).mean() >= 0.99, f"{kw}: {(diff <= 1).mean() * 100:.2f}% within 1/255, worst {int(diff.max())}"
    plain = osc.render(params(w, h, depth, aa))["rgba8"][..., :3]
    assert (plain != ref).any(axis=-1).mean() > 0.02, "the setting changes nothing: not a test"


@pytest.mark.parametrize("w,h,depth", [(64, 48, 3), (96, 72, 6)])
@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_analytic_mode_frame_against_float64_brute_force(pkg, oracle, name, w, h, depth):
    """The oracle's ANALYTIC primitive mode (spheres / boxes with the semantics of the reference's dead SphereInstance / BoxInstance,
    SURVEY A13) against the same float64 renderer with analytic primitives built from the scene description.  The product's analytic
    kernels are checked against this oracle mode on the GPU; this pins the oracle mode itself."""
    obj = synth.sample_scene(name)
    osc, holder = oracle_scene(oracle, obj)
    osc.set_primitive_mode(1)
    p = params(w, h, depth)
    ref = osc.render(p)["rgba8"][..., :3]
    u25 = np.zeros(25, np.float32)
    wh = (C.c_int32 * 2)()
    assert abi.load().rtb_resolve_frame(holder.ptr(), C.byref(p), u25.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
    tri18, mat_idx, _ = osc.triangles()  # analytic mode: the mesh triangles only

    def matrix(ti):
        t = obj.Transformations[ti] if 0 <= ti < len(obj.Transformations) else None  # out of range -> identity, SceneGeometryConverter.cs:85
        return _float64_matrix([(e.Type, *e.XYZ, e.AngleDeg) for e in t.Elements]) if t else np.eye(4)

    prims = [(2, matrix(b.transformationIndex), b.materialIndex) for b in obj.Boxes] + [(1, matrix(sp.transformationIndex), sp.materialIndex) for sp in obj.Spheres]
    assert len(prims) >= 2 and len(tri18) == sum(len(m.materials) for m in obj.TriangleMeshes)
    got = _render64(tri18, mat_idx, obj.Materials, u25, w, h, depth, prims)
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32)).max(-1)
    within = float((diff <= 1).mean())
    assert within >= 0.99, f"{name}: only {within * 100:.2f}% of pixels within 1/255 (worst {int(diff.max())})"


# ---- tessellation: AddCube / AddSphere / meshes in float64, written from SceneGeometryConverter.cs -----------------------------------
def _tessellate64(obj):
    """ExtractTriangles (SceneGeometryConverter.cs:18-51): meshes, then boxes (AddCube :120-155), then spheres (AddSphere :161-230,
    AddSmoothTri :245-264) in float64.  Returns [n, 18] (v0 v1 v2 n0 n1 n2) and materials."""
    def matrix(ti):
        ok = 0 <= ti < len(obj.Transformations)
        return _float64_matrix([(e.Type, *e.XYZ, e.AngleDeg) for e in obj.Transformations[ti].Elements]) if ok else np.eye(4)

    def point(M, v):
        return M[:3, :3] @ v + M[:3, 3]

    rows, mats = [], []

    def flat(a, b, c, m):
        n = np.cross(b - a, c - a)
        n = n / np.linalg.norm(n)
        rows.append(np.concatenate([a, b, c, n, n, n])); mats.append(m)

    for mesh in obj.TriangleMeshes:
        M = matrix(mesh.transformationIndex)
        for k in range(len(mesh.materials)):
            a, b, c = (point(M, mesh.vertices[k, j].astype(np.float64)) for j in range(3))
            flat(a, b, c, int(mesh.materials[k]))
    corners = np.array([[-.5, -.5, -.5], [.5, -.5, -.5], [.5, .5, -.5], [-.5, .5, -.5], [-.5, -.5, .5], [.5, -.5, .5], [.5, .5, .5], [-.5, .5, .5]])
    faces = [(0, 2, 1), (0, 3, 2), (5, 7, 6), (5, 4, 7), (3, 6, 2), (3, 7, 6), (4, 1, 5), (4, 0, 1), (4, 3, 7), (4, 0, 3), (1, 6, 2), (1, 5, 6)]
    for box in obj.Boxes:
        M = matrix(box.transformationIndex)
        v = [point(M, c) for c in corners]
        for i, j, k in faces:
            flat(v[i], v[j], v[k], box.materialIndex)
    nb_long, nb_lat = 24, 16
    sv = np.zeros(((nb_long + 1) * nb_lat + 2, 3))
    sv[0] = (0, 1, 0)
    for lat in range(nb_lat):
        a1 = np.pi * (lat + 1) / (nb_lat + 1)
        for lon in range(nb_long + 1):
            a2 = 2 * np.pi * (0 if lon == nb_long else lon) / nb_long
            sv[lon + lat * (nb_long + 1) + 1] = (np.sin(a1) * np.cos(a2), np.cos(a1), np.sin(a1) * np.sin(a2))
    sv[-1] = (0, -1, 0)
    for sph in obj.Spheres:
        M = matrix(sph.transformationIndex)
        NM = np.linalg.inv(M).T

        def smooth(a, b, c):
            ns = [NM[:3, :3] @ (p / np.linalg.norm(p)) for p in (a, b, c)]
            ns = [n / np.linalg.norm(n) for n in ns]
            rows.append(np.concatenate([point(M, a), point(M, b), point(M, c), *ns])); mats.append(sph.materialIndex)

        for lon in range(nb_long):
            smooth(sv[0], sv[lon + 2], sv[lon + 1])
        for lat in range(nb_lat - 1):
            for lon in range(nb_long):
                cur = lon + lat * (nb_long + 1) + 1
                nxt, below = cur + 1, cur + nb_long + 1
                smooth(sv[cur], sv[below], sv[nxt])
                smooth(sv[nxt], sv[below], sv[below + 1])
        last = len(sv) - 1
        for lon in range(nb_long):
            smooth(sv[last], sv[last - (nb_long + 1) + lon], sv[last - (nb_long + 1) + lon + 1])
    return np.array(rows), np.array(mats, np.int32)


@pytest.mark.parametrize("name", synth.SAMPLE_SCENES)
def test_tessellation_against_float64(pkg, oracle, name):
    """Emission order, winding, vertex positions and (flat / smooth, inverse-transpose) normals of every triangle the oracle
    flattens, against the float64 restatement above: same count and materials, positions within 2e-5, normals within 2e-5."""
    obj = synth.sample_scene(name)
    osc, _ = oracle_scene(oracle, obj)
    vn, mat, _ = osc.triangles()
    want, wmat = _tessellate64(obj)
    assert vn.shape == want.shape and (mat == wmat).all()
    assert np.abs(vn[:, :9] - want[:, :9]).max() <= 2e-5
    assert np.abs(vn[:, 9:] - want[:, 9:]).max() <= 2e-5
