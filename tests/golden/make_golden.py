"""Regenerates tests/golden/* from the reference's shipped scene files with the CPU oracle.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
  ../../cosig-raytracing_b200/scenes/<name>.npz      packed copy of Assets/Resources/Scenes/<name>.txt (parsed by the oracle's SceneService restatement)
  <name>_c1.npz          oracle outputs at the C1 settings (320x240, depth 3, AA 1): primary prim_id / t / material maps and
                         the RGBA8 frame; plus the AA-4 frame
  summary.json           triangle / node counts, counters and SHA-256 of the flattened triangle arrays and BVH nodes

The reference ships no expected outputs, so these files pin the ORACLE against regressions; they are not
outputs of the reference itself ("parity unpinned").
"""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

abi = importlib.import_module("cosig-raytracing_b200.abi")
scene_mod = importlib.import_module("cosig-raytracing_b200.scene")
synth = importlib.import_module("cosig-raytracing_b200.synth")

REF = "/root/reference/Assets/Resources/Scenes"
HERE = os.path.dirname(os.path.abspath(__file__))


def c1_params(aa=1):
    p = abi.default_params()
    p.has_resolution, p.width, p.height = 1, 320, 240
    p.max_depth = 3
    p.aa_samples = aa
    return p


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    O.build()
    summary = {}
    for name in synth.SAMPLE_SCENES:
        sc = O.OracleScene.from_file(os.path.join(REF, name + ".txt"))
        obj = scene_mod.unpack_scene(sc.desc)
        synth.save_scene_npz(os.path.join(HERE, "..", "..", "cosig-raytracing_b200", "scenes", name + ".npz"), obj)
        # the packed copy must rebuild the very same oracle scene
        packed = scene_mod.pack_scene(synth.load_scene_npz(os.path.join(HERE, "..", "..", "cosig-raytracing_b200", "scenes", name + ".npz")))  # keep the buffers alive
        again = O.OracleScene.from_desc(packed.desc)
        vn, mat, cen = sc.triangles()
        vn2, mat2, cen2 = again.triangles()
        assert vn.tobytes() == vn2.tobytes() and mat.tobytes() == mat2.tobytes() and cen.tobytes() == cen2.tobytes()
        nodes, orig = sc.bvh()
        r = sc.render(c1_params(1), want_aux=True, want_rgbf=True)
        r4 = sc.render(c1_params(4))
        c = r["counters"]
        np.savez_compressed(os.path.join(HERE, name + "_c1.npz"), prim=r["prim"], t=r["t"], mat=r["mat"], rgba8=r["rgba8"],
                            rgba8_aa4=r4["rgba8"])
        summary[name] = dict(
            n_triangles=int(sc.n_triangles), n_nodes=int(sc.n_nodes), max_leaf=int(sc.max_leaf),
            sha_triangles=sha(vn), sha_materials=sha(mat), sha_centers=sha(cen), sha_nodes=sha(nodes), sha_perm=sha(orig),
            rays_primary=int(c.rays_primary), rays_continuation=int(c.rays_continuation), rays_shadow=int(c.rays_shadow),
            primary_hits=int(c.primary_hits), nodes_visited=int(c.nodes_visited), tris_tested=int(c.tris_tested), max_stack=int(c.max_stack),
            sha_rgbf=sha(r["rgbf"]))
        print(name, summary[name])
    with open(os.path.join(HERE, "summary.json"), "w") as f:
        json.dump(summary, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
