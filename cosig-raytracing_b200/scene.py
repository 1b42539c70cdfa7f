"""Host-side mirror of the reference's scene model and render settings.

ObjectData and friends follow Assets/Models/ObjectData.cs:9-241 (same class and field names); RenderSettings follows
Assets/Models/RenderSettings.cs:7-70 (nullable overrides are `None`).  Large meshes are held as numpy arrays
(`TrianglesMesh.materials` int32[n], `.vertices` float32[n,3,3]) rather than per-triangle objects, which is the only
structural difference; `Triangle` remains available for small scenes and tests.

`pack_scene` turns an ObjectData into the flat `rtb_scene_desc` the C ABI takes (include/rtb.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi


@dataclass
class ImageSettings:  # ObjectData.cs:40-50
    horizontal: int = 0
    vertical: int = 0
    background: Tuple[float, float, float] = (0.0, 0.0, 0.0)


@dataclass
class TransformElement:  # ObjectData.cs:80-121
    Type: int = abi.RTB_XF_T
    XYZ: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    AngleDeg: float = 0.0

    @staticmethod
    def Translation(t):
        return TransformElement(abi.RTB_XF_T, tuple(t), 0.0)

    @staticmethod
    def Scale(s):
        return TransformElement(abi.RTB_XF_S, tuple(s), 0.0)

    @staticmethod
    def RotationX(a):
        return TransformElement(abi.RTB_XF_RX, (0.0, 0.0, 0.0), a)

    @staticmethod
    def RotationY(a):
        return TransformElement(abi.RTB_XF_RY, (0.0, 0.0, 0.0), a)

    @staticmethod
    def RotationZ(a):
        return TransformElement(abi.RTB_XF_RZ, (0.0, 0.0, 0.0), a)


@dataclass
class CompositeTransformation:  # ObjectData.cs:57-61
    Elements: List[TransformElement] = field(default_factory=list)


@dataclass
class CameraSettings:  # ObjectData.cs:128-138
    transformationIndex: int = 0
    distance: float = 1.0
    verticalFovDeg: float = 60.0


@dataclass
class LightSource:  # ObjectData.cs:144-151
    transformationIndex: int = 0
    rgb: Tuple[float, float, float] = (1.0, 1.0, 1.0)


@dataclass
class MaterialDescription:  # ObjectData.cs:158-176
    color: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    ambient: float = 0.0
    diffuse: float = 0.0
    specular: float = 0.0
    refraction: float = 0.0
    ior: float = 1.0


@dataclass
class Triangle:  # ObjectData.cs:196-215
    materialIndex: int
    v0: Tuple[float, float, float]
    v1: Tuple[float, float, float]
    v2: Tuple[float, float, float]


class TrianglesMesh:  # ObjectData.cs:183-190
    def __init__(self, transformationIndex: int = 0, Triangles: Optional[Sequence[Triangle]] = None,
                 materials: Optional[np.ndarray] = None, vertices: Optional[np.ndarray] = None):
        self.transformationIndex = transformationIndex
        if Triangles is not None:
            materials = np.array([t.materialIndex for t in Triangles], dtype=np.int32)
            vertices = np.array([[t.v0, t.v1, t.v2] for t in Triangles], dtype=np.float32).reshape(-1, 3, 3)
        self.materials = np.zeros(0, np.int32) if materials is None else np.ascontiguousarray(materials, np.int32)
        self.vertices = np.zeros((0, 3, 3), np.float32) if vertices is None else np.ascontiguousarray(vertices, np.float32)
        assert self.vertices.shape == (self.materials.shape[0], 3, 3)

    @property
    def Triangles(self) -> List[Triangle]:
        return [Triangle(int(m), tuple(v[0]), tuple(v[1]), tuple(v[2])) for m, v in zip(self.materials, self.vertices)]


@dataclass
class SphereDescription:  # ObjectData.cs:221-228
    transformationIndex: int = 0
    materialIndex: int = 0


@dataclass
class BoxDescription:  # ObjectData.cs:234-241
    transformationIndex: int = 0
    materialIndex: int = 0


@dataclass
class ObjectData:  # ObjectData.cs:9-34
    Image: Optional[ImageSettings] = None
    Transformations: List[CompositeTransformation] = field(default_factory=list)
    Camera: Optional[CameraSettings] = None
    Lights: List[LightSource] = field(default_factory=list)
    Materials: List[MaterialDescription] = field(default_factory=list)
    TriangleMeshes: List[TrianglesMesh] = field(default_factory=list)
    Spheres: List[SphereDescription] = field(default_factory=list)
    Boxes: List[BoxDescription] = field(default_factory=list)


@dataclass
class RenderSettings:  # RenderSettings.cs:7-70; defaults = what SceneBuilder.GetRenderSettingsFromUI produces
    ResolutionOverride: Optional[Tuple[int, int]] = None
    BackgroundColorOverride: Optional[Tuple[float, float, float]] = None
    LightIntensityScale: float = 1.0
    CameraPositionOverride: Optional[Tuple[float, float, float]] = None
    CameraRotationOverride: Optional[Tuple[float, float, float]] = None
    CameraFovOverride: Optional[float] = None
    MaxDepth: int = 2
    EnableAmbient: bool = True
    EnableDiffuse: bool = True
    EnableSpecular: bool = True
    EnableRefraction: bool = True
    IsOrthographic: bool = False
    AASamples: int = 1
    EnableSoftShadows: bool = False
    LightSize: float = 0.0
    EnableGlossy: bool = False
    SurfaceRoughness: float = 0.0
    EnableMotionBlur: bool = False
    ShutterSpeed: float = 0.0
    # additions of this library (rtb_render_params; zero = reference behaviour)
    DebugMode: int = 0
    SrgbEncode: bool = False

    def to_params(self) -> abi.RenderParams:
        p = abi.RenderParams()
        if self.ResolutionOverride is not None:
            p.has_resolution, p.width, p.height = 1, int(self.ResolutionOverride[0]), int(self.ResolutionOverride[1])
        if self.BackgroundColorOverride is not None:
            p.has_bg = 1
            p.bg[:] = [float(x) for x in self.BackgroundColorOverride[:3]]
        p.light_intensity = float(self.LightIntensityScale)
        if self.CameraPositionOverride is not None:
            p.has_cam_pos = 1
            p.cam_pos[:] = [float(x) for x in self.CameraPositionOverride]
        if self.CameraRotationOverride is not None:
            p.has_cam_rot = 1
            p.cam_rot_euler_deg[:] = [float(x) for x in self.CameraRotationOverride]
        if self.CameraFovOverride is not None:
            p.has_fov, p.fov_deg = 1, float(self.CameraFovOverride)
        p.max_depth = int(self.MaxDepth)
        p.enable_ambient = int(self.EnableAmbient)
        p.enable_diffuse = int(self.EnableDiffuse)
        p.enable_specular = int(self.EnableSpecular)
        p.enable_refraction = int(self.EnableRefraction)
        p.is_orthographic = int(self.IsOrthographic)
        p.aa_samples = int(self.AASamples)
        p.soft_shadows, p.light_size = int(self.EnableSoftShadows), float(self.LightSize)
        p.glossy, p.roughness = int(self.EnableGlossy), float(self.SurfaceRoughness)
        p.motion_blur, p.shutter_speed = int(self.EnableMotionBlur), float(self.ShutterSpeed)
        p.debug_mode = int(self.DebugMode)
        p.srgb_encode = int(self.SrgbEncode)
        return p


class PackedScene:
    """An rtb_scene_desc plus the numpy buffers it points into (kept alive for as long as this object lives)."""

    def __init__(self, scene: ObjectData):
        d = abi.SceneDesc()
        self._keep = []
        if scene.Image is not None:
            d.has_image, d.image_w, d.image_h = 1, int(scene.Image.horizontal), int(scene.Image.vertical)
            d.bg[:] = [float(x) for x in scene.Image.background[:3]]
        if scene.Camera is not None:
            d.has_camera, d.cam_xform = 1, int(scene.Camera.transformationIndex)
            d.cam_distance, d.cam_vfov_deg = float(scene.Camera.distance), float(scene.Camera.verticalFovDeg)
        offs = [0]
        elems = []
        for t in scene.Transformations:
            for e in t.Elements:
                elems.append((int(e.Type), float(e.XYZ[0]), float(e.XYZ[1]), float(e.XYZ[2]), float(e.AngleDeg)))
            offs.append(len(elems))
        self.xoff = np.array(offs, np.int32)
        self.xel = (abi.XformElem * max(1, len(elems)))(*[abi.XformElem(*e) for e in elems])
        d.n_xforms = len(scene.Transformations)
        d.xform_offsets = self.xoff.ctypes.data_as(C.POINTER(C.c_int32))
        d.xform_elems = C.cast(self.xel, C.POINTER(abi.XformElem))
        self.lxf = np.array([l.transformationIndex for l in scene.Lights] or [0], np.int32)
        self.lrgb = np.array([c for l in scene.Lights for c in l.rgb[:3]] or [0, 0, 0], np.float32)
        d.n_lights = len(scene.Lights)
        d.light_xforms = self.lxf.ctypes.data_as(C.POINTER(C.c_int32))
        d.light_rgb = self.lrgb.ctypes.data_as(C.POINTER(C.c_float))
        self.mats = np.array([[*m.color[:3], m.ambient, m.diffuse, m.specular, m.refraction, m.ior] for m in scene.Materials]
                             or [[0] * 8], np.float32)
        d.n_materials = len(scene.Materials)
        d.materials = self.mats.ctypes.data_as(C.POINTER(abi.Material))
        # triangles: rtb_triangle is {int32 material; float v[9]} = 10 words
        n = sum(m.materials.shape[0] for m in scene.TriangleMeshes)
        tri = np.zeros((max(1, n), 10), np.float32)
        meshes = (abi.Mesh * max(1, len(scene.TriangleMeshes)))()
        at = 0
        for i, m in enumerate(scene.TriangleMeshes):
            k = m.materials.shape[0]
            tri[at:at + k, 0] = m.materials.view(np.float32)
            tri[at:at + k, 1:] = m.vertices.reshape(k, 9)
            meshes[i].xform, meshes[i].first_tri, meshes[i].n_tris = int(m.transformationIndex), at, k
            at += k
        self.tri, self.meshes = tri, meshes
        d.n_meshes = len(scene.TriangleMeshes)
        d.meshes = C.cast(meshes, C.POINTER(abi.Mesh))
        d.n_triangles = n
        d.triangles = tri.ctypes.data_as(C.POINTER(abi.Triangle))
        self.sph = np.array([[s.transformationIndex, s.materialIndex] for s in scene.Spheres] or [[0, 0]], np.int32)
        self.box = np.array([[b.transformationIndex, b.materialIndex] for b in scene.Boxes] or [[0, 0]], np.int32)
        d.n_spheres, d.spheres = len(scene.Spheres), self.sph.ctypes.data_as(C.POINTER(abi.Prim))
        d.n_boxes, d.boxes = len(scene.Boxes), self.box.ctypes.data_as(C.POINTER(abi.Prim))
        self.desc = d

    def ptr(self):
        return C.byref(self.desc)


def pack_scene(scene: ObjectData) -> PackedScene:
    return PackedScene(scene)


def unpack_scene(desc: abi.SceneDesc) -> ObjectData:
    """rtb_scene_desc (e.g. from the native scene-file parser) -> ObjectData."""
    s = ObjectData()
    if desc.has_image:
        s.Image = ImageSettings(desc.image_w, desc.image_h, tuple(desc.bg))
    if desc.has_camera:
        s.Camera = CameraSettings(desc.cam_xform, desc.cam_distance, desc.cam_vfov_deg)
    for i in range(desc.n_xforms):
        ct = CompositeTransformation()
        for k in range(desc.xform_offsets[i], desc.xform_offsets[i + 1]):
            e = desc.xform_elems[k]
            ct.Elements.append(TransformElement(e.type, (e.x, e.y, e.z), e.angle_deg))
        s.Transformations.append(ct)
    for i in range(desc.n_lights):
        s.Lights.append(LightSource(desc.light_xforms[i], tuple(desc.light_rgb[3 * i + k] for k in range(3))))
    for i in range(desc.n_materials):
        m = desc.materials[i]
        s.Materials.append(MaterialDescription((m.r, m.g, m.b), m.ka, m.kd, m.ks, m.kr, m.ior))
    if desc.n_triangles:
        raw = np.ctypeslib.as_array(C.cast(desc.triangles, C.POINTER(C.c_float)), shape=(desc.n_triangles, 10))
    for i in range(desc.n_meshes):
        m = desc.meshes[i]
        blk = raw[m.first_tri:m.first_tri + m.n_tris] if m.n_tris else np.zeros((0, 10), np.float32)
        s.TriangleMeshes.append(TrianglesMesh(m.xform, materials=blk[:, 0].copy().view(np.int32),
                                              vertices=blk[:, 1:].copy().reshape(-1, 3, 3)))
    for i in range(desc.n_spheres):
        s.Spheres.append(SphereDescription(desc.spheres[i].xform, desc.spheres[i].material))
    for i in range(desc.n_boxes):
        s.Boxes.append(BoxDescription(desc.boxes[i].xform, desc.boxes[i].material))
    return s
