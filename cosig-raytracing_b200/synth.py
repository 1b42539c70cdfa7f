"""Deterministic synthetic scenes for the perf configurations C3-C5 and a loader for the packed copies
of the reference's three sample scenes (cosig-raytracing_b200/scenes/*.npz, made by tests/golden/make_golden.py).

No RNG state: every value is a closed-form function of indices, so every process (and every rank of a multi-GPU run)
builds bit-identical scenes.
"""
from __future__ import annotations

import os

import numpy as np

from . import abi
from .scene import (BoxDescription, CameraSettings, CompositeTransformation, ImageSettings, LightSource, MaterialDescription,
                    ObjectData, SphereDescription, TransformElement, TrianglesMesh)

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scenes")  # packed copies of the reference's three sample scenes, part of the package


def _sample_camera_and_light(scene: ObjectData):
    """Camera and light of the reference's sample scene (test_scene_1.txt): T[0] identity, T[1] camera, T[2] light."""
    scene.Transformations.append(CompositeTransformation([]))
    scene.Transformations.append(CompositeTransformation([TransformElement.Translation((0, 0, -74)), TransformElement.RotationX(-60.0),
                                                           TransformElement.RotationZ(45.0)]))
    scene.Transformations.append(CompositeTransformation([TransformElement.Translation((-7, 5, 66))]))
    scene.Camera = CameraSettings(1, 30.0, 30.0)
    scene.Lights.append(LightSource(2, (1.0, 1.0, 1.0)))
    scene.Image = ImageSettings(3840, 2160, (0.2, 0.2, 0.2))


def _pcg_hash(idx: np.ndarray, seed: int) -> np.ndarray:
    """PCG RXS-M-XS 32-bit output function of (idx ^ seed); uint32 in, uint32 out."""
    with np.errstate(over="ignore"):
        state = (idx.astype(np.uint32) ^ np.uint32(seed)) * np.uint32(747796405) + np.uint32(2891336453)
        word = ((state >> ((state >> np.uint32(28)) + np.uint32(4))) ^ state) * np.uint32(277803737)
        return (word >> np.uint32(22)) ^ word


def heightfield_scene(cells_x: int = 1000, cells_y: int = 500, seed: int = 0x5EED) -> ObjectData:
    """C4: 2*cells_x*cells_y triangles (1 000 000 by default) over [-32,32]^2,
    z = 3 sin(0.37 x) cos(0.29 y) + 0.25 (hash(vertex)/2^32 - 0.5); material per 25x25-cell block cycling
    diffuse / mirror / glass / diffuse; one mesh with the identity transform."""
    s = ObjectData()
    _sample_camera_and_light(s)
    s.Materials = [
        MaterialDescription((0.8, 0.6, 0.3), 0.1, 0.7, 0.0, 0.0, 1.0),
        MaterialDescription((0.9, 0.9, 0.9), 0.05, 0.2, 0.8, 0.0, 1.0),
        MaterialDescription((0.9, 0.95, 1.0), 0.05, 0.1, 0.0, 0.9, 1.5),
        MaterialDescription((0.3, 0.6, 0.8), 0.1, 0.7, 0.0, 0.0, 1.0),
    ]
    nx, ny = cells_x + 1, cells_y + 1
    x = -32.0 + 64.0 * np.arange(nx, dtype=np.float64) / cells_x
    y = -32.0 + 64.0 * np.arange(ny, dtype=np.float64) / cells_y
    X, Y = np.meshgrid(x, y, indexing="xy")  # [ny, nx]
    vid = (np.arange(ny, dtype=np.uint32)[:, None] * np.uint32(nx) + np.arange(nx, dtype=np.uint32)[None, :])
    noise = _pcg_hash(vid, seed).astype(np.float64) / 4294967296.0 - 0.5
    Z = 3.0 * np.sin(0.37 * X) * np.cos(0.29 * Y) + 0.25 * noise
    P = np.stack([X, Y, Z], axis=-1).astype(np.float32)  # [ny, nx, 3]
    v00, v10, v11, v01 = P[:-1, :-1], P[:-1, 1:], P[1:, 1:], P[1:, :-1]
    tris = np.empty((cells_y, cells_x, 2, 3, 3), np.float32)
    tris[:, :, 0, 0], tris[:, :, 0, 1], tris[:, :, 0, 2] = v00, v10, v11
    tris[:, :, 1, 0], tris[:, :, 1, 1], tris[:, :, 1, 2] = v00, v11, v01
    bx = (np.arange(cells_x) // 25)[None, :]
    by = (np.arange(cells_y) // 25)[:, None]
    mat = ((bx + by) % 4).astype(np.int32)
    mats = np.broadcast_to(mat[:, :, None], (cells_y, cells_x, 2))
    s.TriangleMeshes.append(TrianglesMesh(0, materials=np.ascontiguousarray(mats).reshape(-1), vertices=tris.reshape(-1, 3, 3)))
    return s


def sphere_grid_scene(n: int = 16) -> ObjectData:
    """C3: n x n unit spheres at (3(i-(n-1)/2), 3(j-(n-1)/2), 1), glass / mirror alternating by (i+j)&1, on a 64x64x1 floor box."""
    s = ObjectData()
    _sample_camera_and_light(s)
    s.Materials = [
        MaterialDescription((0.9, 0.95, 1.0), 0.05, 0.1, 0.0, 0.9, 1.5),   # glass
        MaterialDescription((0.9, 0.9, 0.9), 0.05, 0.2, 0.8, 0.0, 1.0),    # mirror
        MaterialDescription((0.8, 0.8, 0.8), 0.1, 0.7, 0.2, 0.0, 1.0),     # floor
    ]
    half = (n - 1) / 2.0
    for j in range(n):
        for i in range(n):
            s.Transformations.append(CompositeTransformation([TransformElement.Translation((3.0 * (i - half), 3.0 * (j - half), 1.0))]))
            s.Spheres.append(SphereDescription(len(s.Transformations) - 1, (i + j) & 1))
    s.Transformations.append(CompositeTransformation([TransformElement.Translation((0.0, 0.0, -0.5)), TransformElement.Scale((64.0, 64.0, 1.0))]))
    s.Boxes.append(BoxDescription(len(s.Transformations) - 1, 2))
    return s


SAMPLE_SCENES = ("test_scene_1", "test_scene_2", "eval_scene")


def save_scene_npz(path: str, s: ObjectData):
    elems = np.array([[e.Type, *e.XYZ, e.AngleDeg] for t in s.Transformations for e in t.Elements] or np.zeros((0, 5)), np.float64)
    offs = np.cumsum([0] + [len(t.Elements) for t in s.Transformations]).astype(np.int32)
    tri_mat = np.concatenate([m.materials for m in s.TriangleMeshes] or [np.zeros(0, np.int32)])
    tri_v = np.concatenate([m.vertices for m in s.TriangleMeshes] or [np.zeros((0, 3, 3), np.float32)])
    np.savez_compressed(
        path,
        image=np.array([s.Image.horizontal, s.Image.vertical, *s.Image.background[:3]] if s.Image else [], np.float64),
        camera=np.array([s.Camera.transformationIndex, s.Camera.distance, s.Camera.verticalFovDeg] if s.Camera else [], np.float64),
        xform_offsets=offs, xform_elems=elems.astype(np.float32),
        lights=np.array([[l.transformationIndex, *l.rgb[:3]] for l in s.Lights] or np.zeros((0, 4)), np.float32),
        materials=np.array([[*m.color[:3], m.ambient, m.diffuse, m.specular, m.refraction, m.ior] for m in s.Materials] or np.zeros((0, 8)), np.float32),
        meshes=np.array([[m.transformationIndex, m.materials.shape[0]] for m in s.TriangleMeshes] or np.zeros((0, 2)), np.int64),
        tri_mat=tri_mat.astype(np.int32), tri_v=tri_v.astype(np.float32),
        spheres=np.array([[p.transformationIndex, p.materialIndex] for p in s.Spheres] or np.zeros((0, 2)), np.int32),
        boxes=np.array([[p.transformationIndex, p.materialIndex] for p in s.Boxes] or np.zeros((0, 2)), np.int32))


def load_scene_npz(path: str) -> ObjectData:
    z = np.load(path)
    s = ObjectData()
    if z["image"].size:
        im = z["image"]
        s.Image = ImageSettings(int(im[0]), int(im[1]), (float(np.float32(im[2])), float(np.float32(im[3])), float(np.float32(im[4]))))
    if z["camera"].size:
        c = z["camera"]
        s.Camera = CameraSettings(int(c[0]), float(np.float32(c[1])), float(np.float32(c[2])))
    offs, el = z["xform_offsets"], z["xform_elems"]
    for i in range(len(offs) - 1):
        s.Transformations.append(CompositeTransformation([TransformElement(int(e[0]), (float(e[1]), float(e[2]), float(e[3])), float(e[4]))
                                                          for e in el[offs[i]:offs[i + 1]]]))
    for l in z["lights"]:
        s.Lights.append(LightSource(int(l[0]), (float(l[1]), float(l[2]), float(l[3]))))
    for m in z["materials"]:
        s.Materials.append(MaterialDescription((float(m[0]), float(m[1]), float(m[2])), *[float(v) for v in m[3:8]]))
    at = 0
    for xf, cnt in z["meshes"]:
        s.TriangleMeshes.append(TrianglesMesh(int(xf), materials=z["tri_mat"][at:at + cnt], vertices=z["tri_v"][at:at + cnt]))
        at += int(cnt)
    for p in z["spheres"]:
        s.Spheres.append(SphereDescription(int(p[0]), int(p[1])))
    for p in z["boxes"]:
        s.Boxes.append(BoxDescription(int(p[0]), int(p[1])))
    return s


def sample_scene(name: str = "test_scene_1") -> ObjectData:
    """One of the reference's three shipped scenes (Assets/Resources/Scenes/<name>.txt), from its packed copy."""
    return load_scene_npz(os.path.join(_GOLDEN, name + ".npz"))


def scene_to_text(s: ObjectData) -> str:
    """Serialises an ObjectData in the COSIG scene text format (SceneService.cs), CRLF + tabs like the shipped files."""
    out = []
    names = {abi.RTB_XF_T: "T", abi.RTB_XF_S: "S", abi.RTB_XF_RX: "Rx", abi.RTB_XF_RY: "Ry", abi.RTB_XF_RZ: "Rz"}
    g = lambda v: np.format_float_positional(np.float32(v), unique=True, trim="-")
    if s.Image:
        out += ["Image", "{", f"\t{s.Image.horizontal} {s.Image.vertical}", "\t" + " ".join(g(c) for c in s.Image.background[:3]), "}", ""]
    for t in s.Transformations:
        out += ["Transformation", "{"]
        for e in t.Elements:
            if e.Type in (abi.RTB_XF_T, abi.RTB_XF_S):
                out.append(f"\t{names[e.Type]} " + " ".join(g(c) for c in e.XYZ))
            else:
                out.append(f"\t{names[e.Type]} {g(e.AngleDeg)}")
        out += ["}", ""]
    if s.Camera:
        out += ["Camera", "{", f"\t{s.Camera.transformationIndex}", f"\t{g(s.Camera.distance)}", f"\t{g(s.Camera.verticalFovDeg)}", "}", ""]
    for l in s.Lights:
        out += ["Light", "{", f"\t{l.transformationIndex}", "\t" + " ".join(g(c) for c in l.rgb[:3]), "}", ""]
    for m in s.Materials:
        out += ["Material", "{", "\t" + " ".join(g(c) for c in m.color[:3]),
                "\t" + " ".join(g(c) for c in (m.ambient, m.diffuse, m.specular, m.refraction, m.ior)), "}", ""]
    for mesh in s.TriangleMeshes:
        out += ["Triangles", "{", f"\t{mesh.transformationIndex}"]
        for mi, v in zip(mesh.materials, mesh.vertices):
            out.append(f"\t{int(mi)}")
            for k in range(3):
                out.append("\t" + " ".join(g(c) for c in v[k]))
        out += ["}", ""]
    for p in s.Boxes:
        out += ["Box", "{", f"\t{p.transformationIndex}", f"\t{p.materialIndex}", "}", ""]
    for p in s.Spheres:
        out += ["Sphere", "{", f"\t{p.transformationIndex}", f"\t{p.materialIndex}", "}", ""]
    return "\r\n".join(out)
