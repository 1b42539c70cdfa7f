#!/usr/bin/env python
"""tools/bvh_study.py — OFFLINE PLANNING TOOL (CPU only): node records and triangle tests per ray of the library's ordered
traversal scheme under three builders (Morton LBVH as csrc/lbvh.cu builds it, binned SAH, the reference's spatial median), on
the C4 height field and the C3 sphere grid, for primary, mirror-reflection and shadow rays.

    python tools/bvh_study.py [--width 480 --height 270]

It answers one question for the next round: how much traversal work would a SAH-quality build on the GPU save?  Rays come from
the oracle's primary-hit maps (test infrastructure; this tool is not part of the product either).
"""
import argparse
import ctypes as C
import importlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--height", type=int, default=270)
    args = ap.parse_args()
    from oracle import oracle_py as O
    from util import abi, params, scene_mod, synth
    O.build()
    exe = os.path.join(ROOT, "tools", "bvh_study")
    subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "bvh_study.cpp")])
    tmp = tempfile.mkdtemp()
    w, h = args.width, args.height
    for name, obj in (("C4 height field (1 M triangles)", synth.heightfield_scene(1000, 500)), ("C3 sphere grid (196 620 triangles)", synth.sphere_grid_scene(16))):
        packed = scene_mod.pack_scene(obj)
        osc = O.OracleScene.from_desc(packed.desc)
        if "C3" in name:
            osc.set_primitive_mode(1)  # analytic primary hits: the reference-shape BVH degenerates on the tessellated grid
        tri18, _, _ = osc.triangles() if "C4" in name else O.OracleScene.from_desc(packed.desc).triangles()
        tri = np.ascontiguousarray(tri18[:, :9], np.float32)
        p = params(w, h, 1)
        u25 = np.zeros(25, np.float32)
        wh = (C.c_int32 * 2)()
        assert abi.load().rtb_resolve_frame(packed.ptr(), C.byref(p), u25.ctypes.data_as(C.POINTER(C.c_float)), wh) == abi.RTB_OK
        M = u25[:16].reshape(4, 4).astype(np.float64)
        cam_d, tan_half, light = float(u25[16]), float(u25[17]), u25[19:22].astype(np.float64)
        ys, xs = np.mgrid[0:h, 0:w]
        ph = 2.0 * cam_d * tan_half
        pw = ph * (w / h)
        dc = np.stack([((xs.ravel() + 0.5) / w - 0.5) * pw, ((ys.ravel() + 0.5) / h - 0.5) * ph, np.full(w * h, -cam_d)], -1)
        dc /= np.linalg.norm(dc, axis=-1, keepdims=True)
        o = np.broadcast_to(M[:3, :3] @ np.array([0.0, 0.0, cam_d]) + M[:3, 3], dc.shape)
        d = dc @ M[:3, :3].T
        d /= np.linalg.norm(d, axis=-1, keepdims=True)
        aux = osc.render(p, want_aux=True)
        t = aux["t"].ravel().astype(np.float64)
        hit = aux["prim"].ravel() >= 0
        pos = o[hit] + t[hit, None] * d[hit]
        # surface normal: numerical, from the hit geometry (height field: flat triangle normal; spheres: centre direction is not
        # available here, so the geometric normal of the nearest tessellation triangle would be needed — use the light-facing
        # finite-difference of neighbouring hits instead: good enough to generate representative secondary rays)
        if "C4" in name:
            prim = aux["prim"].ravel()[hit]
            v0, v1, v2 = tri[prim, 0:3].astype(np.float64), tri[prim, 3:6].astype(np.float64), tri[prim, 6:9].astype(np.float64)
            nrm = np.cross(v1 - v0, v2 - v0)
        else:
            centre = np.round((pos[:, :2] + 22.5) / 3.0) * 3.0 - 22.5
            on_sphere = pos[:, 2] > 1e-3
            nrm = np.where(on_sphere[:, None], pos - np.concatenate([centre, np.ones((len(pos), 1))], -1), np.array([0.0, 0.0, 1.0]))
        nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
        dh = d[hit]
        refl = dh - 2.0 * (nrm * dh).sum(-1, keepdims=True) * nrm
        to_l = light - pos
        dist = np.linalg.norm(to_l, axis=-1)
        ldir = to_l / dist[:, None]
        lit_side = (nrm * ldir).sum(-1) > 0
        sets = {
            "primary": np.concatenate([o, d, np.zeros((len(d), 1))], -1),
            "reflection": np.concatenate([pos + nrm * 1e-2, refl, np.zeros((len(pos), 1))], -1),
            "shadow": np.concatenate([(pos + nrm * 1e-2)[lit_side], ldir[lit_side], dist[lit_side, None]], -1),
        }
        tri_path = os.path.join(tmp, "tri.bin")
        with open(tri_path, "wb") as f:
            f.write(np.int64(len(tri)).tobytes())
            f.write(tri.tobytes())
        print(f"== {name}, {w}x{h} primary rays, {int(hit.sum())} hits", flush=True)
        for kind, rays in sets.items():
            ray_path = os.path.join(tmp, "rays.bin")
            with open(ray_path, "wb") as f:
                f.write(np.int64(len(rays)).tobytes())
                f.write(np.ascontiguousarray(rays, np.float32).tobytes())
            out = subprocess.run([exe, tri_path, ray_path], capture_output=True, text=True, check=True).stdout
            for line in out.strip().splitlines():
                print(f"  {kind:10s} {line}", flush=True)


if __name__ == "__main__":
    main()
