"""rtb200 — B200-native replacement of the per-pixel render path of mpoboas/cosig-raytracing.

The product is csrc/ (CUDA kernels + C ABI, built into librtb200.so); the Python modules mirror the reference's host
API (RayTracer, ObjectData, RenderSettings, SceneService) over that ABI.  Import with
importlib.import_module("cosig-raytracing_b200") (the directory name is the project's name and is not an identifier).
"""
from . import abi, bands, scene, synth  # noqa: F401
from .gif_generator import GifGenerator  # noqa: F401
from .raytracer import DeviceTexture, RayTracer, RtbError, SceneService, Texture2D  # noqa: F401
from .scene import (BoxDescription, CameraSettings, CompositeTransformation, ImageSettings, LightSource, MaterialDescription,  # noqa: F401
                    ObjectData, RenderSettings, SphereDescription, TransformElement, Triangle, TrianglesMesh)
