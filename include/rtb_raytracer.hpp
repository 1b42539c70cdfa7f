// rtb_raytracer.hpp — C++ host mirror of the reference's render API over the C ABI (include/rtb.h).
//
// The reference's host side is C# (Unity); no .NET / Mono / Unity toolchain exists in the build image, so the compiled-
// language host mirror is C++ (the C# P/Invoke stub a maintainer would add is in INTEGRATION.md and csharp/).  Class and
// member names follow the reference one to one:
//
//   ObjectData & friends     Assets/Models/ObjectData.cs:9-241
//   RenderSettings           Assets/Models/RenderSettings.cs:7-70   (nullable overrides -> std::optional)
//   RayTracer                Assets/Services/RayTracer.cs:17        InvalidateBVHCache :38, ReleaseBuffers :47,
//                                                                   ClearRenderTarget :65, RenderToTexture :82, RenderAsync :212
//   SceneService::LoadScene  Assets/Services/SceneService.cs:26
//
// Error behaviour: like the reference, RenderAsync / RenderToTexture return an empty result when no scene can be rendered
// (null shader -> null, RayTracer.cs:84-88) or the render was cancelled (:283); everything the reference would throw on
// raises rtb::Error.  Header-only; link with -lrtb200.
#pragma once
#include <array>
#include <cstring>
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "rtb.h"

namespace rtb {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error("rtb error " + std::to_string(c) + ": " + what), code(c) {}
};

struct Vector3 { float x = 0, y = 0, z = 0; };
struct Color { float r = 0, g = 0, b = 0; };

// ---- ObjectData.cs -----------------------------------------------------------------------------------------------------
struct ImageSettings { int horizontal = 0, vertical = 0; Color background{}; };
enum class TransformType { T = RTB_XF_T, Rx = RTB_XF_RX, Ry = RTB_XF_RY, Rz = RTB_XF_RZ, S = RTB_XF_S };
struct TransformElement {
  TransformType Type = TransformType::T;
  Vector3 XYZ{};
  float AngleDeg = 0;
  static TransformElement Translation(Vector3 t) { return {TransformType::T, t, 0}; }
  static TransformElement Scale(Vector3 s) { return {TransformType::S, s, 0}; }
  static TransformElement RotationX(float a) { return {TransformType::Rx, {}, a}; }
  static TransformElement RotationY(float a) { return {TransformType::Ry, {}, a}; }
  static TransformElement RotationZ(float a) { return {TransformType::Rz, {}, a}; }
};
struct CompositeTransformation { std::vector<TransformElement> Elements; };
struct CameraSettings { int transformationIndex = 0; float distance = 1.0f, verticalFovDeg = 60.0f; };
struct LightSource { int transformationIndex = 0; Color rgb{1, 1, 1}; };
struct MaterialDescription { Color color{1, 1, 1}; float ambient = 0, diffuse = 0, specular = 0, refraction = 0, ior = 1; };
struct Triangle { int materialIndex = 0; Vector3 v0, v1, v2; };
struct TrianglesMesh { int transformationIndex = 0; std::vector<Triangle> Triangles; };
struct SphereDescription { int transformationIndex = 0, materialIndex = 0; };
struct BoxDescription { int transformationIndex = 0, materialIndex = 0; };

struct ObjectData {
  std::optional<ImageSettings> Image;
  std::vector<CompositeTransformation> Transformations;
  std::optional<CameraSettings> Camera;
  std::vector<LightSource> Lights;
  std::vector<MaterialDescription> Materials;
  std::vector<TrianglesMesh> TriangleMeshes;
  std::vector<SphereDescription> Spheres;
  std::vector<BoxDescription> Boxes;
};

// ---- RenderSettings.cs ---------------------------------------------------------------------------------------------------
struct RenderSettings {
  std::optional<std::array<int, 2>> ResolutionOverride;
  std::optional<Color> BackgroundColorOverride;
  float LightIntensityScale = 1.0f;
  std::optional<Vector3> CameraPositionOverride, CameraRotationOverride;
  std::optional<float> CameraFovOverride;
  int MaxDepth = 2;
  bool EnableAmbient = true, EnableDiffuse = true, EnableSpecular = true, EnableRefraction = true;
  bool IsOrthographic = false;
  int AASamples = 1;
  bool EnableSoftShadows = false; float LightSize = 0;
  bool EnableGlossy = false; float SurfaceRoughness = 0;
  bool EnableMotionBlur = false; float ShutterSpeed = 0;

  rtb_render_params ToParams() const {
    rtb_render_params p;
    rtb_params_default(&p);
    if (ResolutionOverride) { p.has_resolution = 1; p.width = (*ResolutionOverride)[0]; p.height = (*ResolutionOverride)[1]; }
    if (BackgroundColorOverride) { p.has_bg = 1; p.bg[0] = BackgroundColorOverride->r; p.bg[1] = BackgroundColorOverride->g; p.bg[2] = BackgroundColorOverride->b; }
    p.light_intensity = LightIntensityScale;
    if (CameraPositionOverride) { p.has_cam_pos = 1; p.cam_pos[0] = CameraPositionOverride->x; p.cam_pos[1] = CameraPositionOverride->y; p.cam_pos[2] = CameraPositionOverride->z; }
    if (CameraRotationOverride) { p.has_cam_rot = 1; p.cam_rot_euler_deg[0] = CameraRotationOverride->x; p.cam_rot_euler_deg[1] = CameraRotationOverride->y; p.cam_rot_euler_deg[2] = CameraRotationOverride->z; }
    if (CameraFovOverride) { p.has_fov = 1; p.fov_deg = *CameraFovOverride; }
    p.max_depth = MaxDepth;
    p.enable_ambient = EnableAmbient; p.enable_diffuse = EnableDiffuse; p.enable_specular = EnableSpecular; p.enable_refraction = EnableRefraction;
    p.is_orthographic = IsOrthographic;
    p.aa_samples = AASamples;
    p.soft_shadows = EnableSoftShadows; p.light_size = LightSize;
    p.glossy = EnableGlossy; p.roughness = SurfaceRoughness;
    p.motion_blur = EnableMotionBlur; p.shutter_speed = ShutterSpeed;
    return p;
  }
};

// Flat view of an ObjectData for rtb_upload_scene; owns the arrays the rtb_scene_desc points into.
class PackedScene {
 public:
  explicit PackedScene(const ObjectData& s) {
    std::memset(&d_, 0, sizeof d_);
    if (s.Image) { d_.has_image = 1; d_.image_w = s.Image->horizontal; d_.image_h = s.Image->vertical; d_.bg[0] = s.Image->background.r; d_.bg[1] = s.Image->background.g; d_.bg[2] = s.Image->background.b; }
    if (s.Camera) { d_.has_camera = 1; d_.cam_xform = s.Camera->transformationIndex; d_.cam_distance = s.Camera->distance; d_.cam_vfov_deg = s.Camera->verticalFovDeg; }
    xoff_.push_back(0);
    for (const auto& t : s.Transformations) {
      for (const auto& e : t.Elements) xel_.push_back(rtb_xform_elem{(int32_t)e.Type, e.XYZ.x, e.XYZ.y, e.XYZ.z, e.AngleDeg});
      xoff_.push_back((int32_t)xel_.size());
    }
    for (const auto& l : s.Lights) { lxf_.push_back(l.transformationIndex); lrgb_.insert(lrgb_.end(), {l.rgb.r, l.rgb.g, l.rgb.b}); }
    for (const auto& m : s.Materials) mats_.push_back(rtb_material{m.color.r, m.color.g, m.color.b, m.ambient, m.diffuse, m.specular, m.refraction, m.ior});
    for (const auto& mesh : s.TriangleMeshes) {
      rtb_mesh rm{};
      rm.xform = mesh.transformationIndex; rm.first_tri = (int64_t)tris_.size(); rm.n_tris = (int64_t)mesh.Triangles.size();
      for (const auto& t : mesh.Triangles)
        tris_.push_back(rtb_triangle{t.materialIndex, {t.v0.x, t.v0.y, t.v0.z}, {t.v1.x, t.v1.y, t.v1.z}, {t.v2.x, t.v2.y, t.v2.z}});
      meshes_.push_back(rm);
    }
    for (const auto& p : s.Spheres) spheres_.push_back(rtb_prim{p.transformationIndex, p.materialIndex});
    for (const auto& p : s.Boxes) boxes_.push_back(rtb_prim{p.transformationIndex, p.materialIndex});
    d_.n_xforms = (int32_t)xoff_.size() - 1; d_.xform_offsets = xoff_.data(); d_.xform_elems = xel_.data();
    d_.n_lights = (int32_t)lxf_.size(); d_.light_xforms = lxf_.data(); d_.light_rgb = lrgb_.data();
    d_.n_materials = (int32_t)mats_.size(); d_.materials = mats_.data();
    d_.n_meshes = (int32_t)meshes_.size(); d_.meshes = meshes_.data();
    d_.n_triangles = (int64_t)tris_.size(); d_.triangles = tris_.data();
    d_.n_spheres = (int32_t)spheres_.size(); d_.spheres = spheres_.data();
    d_.n_boxes = (int32_t)boxes_.size(); d_.boxes = boxes_.data();
  }
  const rtb_scene_desc* desc() const { return &d_; }

 private:
  rtb_scene_desc d_;
  std::vector<int32_t> xoff_, lxf_;
  std::vector<rtb_xform_elem> xel_;
  std::vector<float> lrgb_;
  std::vector<rtb_material> mats_;
  std::vector<rtb_mesh> meshes_;
  std::vector<rtb_triangle> tris_;
  std::vector<rtb_prim> spheres_, boxes_;
};

// RGBA32 pixels, row 0 = bottom (Unity Texture2D convention).
struct Texture2D { int width = 0, height = 0; std::vector<uint8_t> pixels; };
// Frame left in device memory (RenderToTexture: no readback).
struct RenderTexture { void* device_ptr = nullptr; int width = 0, height = 0; };

class RayTracer {
 public:
  explicit RayTracer(const std::vector<int32_t>& devices = {}, int bvh_mode = RTB_BVH_REFERENCE, int primitive_mode = RTB_PRIM_TESSELLATED)
      : bvh_mode_(bvh_mode), primitive_mode_(primitive_mode) {
    const int rc = rtb_create(&ctx_, devices.empty() ? nullptr : devices.data(), (int32_t)devices.size());
    if (rc != RTB_OK) throw Error(rc, rtb_last_error(nullptr));
  }
  ~RayTracer() { rtb_destroy(ctx_); }
  RayTracer(const RayTracer&) = delete;
  RayTracer& operator=(const RayTracer&) = delete;

  void InvalidateBVHCache() { needs_rebuild_ = true; check(rtb_invalidate(ctx_)); }
  void ReleaseBuffers() { cached_ = nullptr; needs_rebuild_ = true; check(rtb_invalidate(ctx_)); check(rtb_clear_target(ctx_)); }
  void ClearRenderTarget() { check(rtb_clear_target(ctx_)); }

  // RayTracer.cs:212-380 (blocking).  `cancel` plays the CancellationToken: polled between wavefront depths.
  std::optional<Texture2D> RenderAsync(const ObjectData* scene, const RenderSettings& settings, const volatile int32_t* cancel = nullptr) {
    if (!ensure_scene(scene)) return std::nullopt;
    const rtb_render_params p = settings.ToParams();
    int32_t wh[2];
    check_create(rtb_resolve_frame(packed_->desc(), &p, nullptr, wh));
    Texture2D tex;
    tex.width = wh[0]; tex.height = wh[1];
    tex.pixels.resize((size_t)wh[0] * wh[1] * 4);
    rtb_set_cancel_flag(ctx_, cancel);
    const int rc = rtb_render(ctx_, &p, tex.pixels.data(), tex.pixels.size(), nullptr, nullptr);
    rtb_set_cancel_flag(ctx_, nullptr);
    if (rc == RTB_E_CANCELLED) return std::nullopt;
    check(rc);
    return tex;
  }

  // RayTracer.cs:82-202: the frame stays on the device.
  std::optional<RenderTexture> RenderToTexture(const ObjectData* scene, const RenderSettings& settings) {
    if (!ensure_scene(scene)) return std::nullopt;
    const rtb_render_params p = settings.ToParams();
    int32_t wh[2];
    check_create(rtb_resolve_frame(packed_->desc(), &p, nullptr, wh));
    RenderTexture rt;
    rt.width = wh[0]; rt.height = wh[1];
    check(rtb_frame_export(ctx_, (size_t)wh[0] * wh[1] * 4, &rt.device_ptr, nullptr));  // the context's own frame buffer; nothing is exported
    check(rtb_render_device(ctx_, &p, rt.device_ptr, (size_t)wh[0] * wh[1] * 4, 1));
    return rt;
  }

  rtb_stats Stats() { rtb_stats s; check(rtb_get_stats(ctx_, &s)); return s; }
  rtb_context* Context() { return ctx_; }
  bool EnsureScene(const ObjectData* scene) { return ensure_scene(scene); }  // uploads when the scene object changed (RayTracer.cs:118-123)

 private:
  bool ensure_scene(const ObjectData* scene) {
    if (!scene) return false;
    if (needs_rebuild_ || cached_ != scene) {  // RayTracer.cs:118-123: the cache key is the scene object's identity
      packed_ = std::make_unique<PackedScene>(*scene);
      check(rtb_upload_scene(ctx_, packed_->desc(), primitive_mode_, bvh_mode_));
      cached_ = scene;
      needs_rebuild_ = false;
    }
    return true;
  }
  void check(int rc) { if (rc != RTB_OK) throw Error(rc, rtb_last_error(ctx_)); }
  void check_create(int rc) { if (rc != RTB_OK) throw Error(rc, rtb_last_error(nullptr)); }

  rtb_context* ctx_ = nullptr;
  int bvh_mode_, primitive_mode_;
  const ObjectData* cached_ = nullptr;
  bool needs_rebuild_ = true;
  std::unique_ptr<PackedScene> packed_;
};

// GifGenerator, Assets/Services/GifGenerator.cs:17-31: the 36-frame rotation sweep and the GIF89a writer.
class GifGenerator {
 public:
  GifGenerator(RayTracer& rayTracer, const ObjectData* scene) : rayTracer_(rayTracer), scene_(scene) {}

  // GenerateRotationFrames :40-72: frame k renders with CameraRotationOverride = (base.x, base.y, 10 k), k = 0..35.
  std::vector<Texture2D> GenerateRotationFrames(const RenderSettings& baseSettings, const std::function<void(float, const std::string&)>& progress = nullptr,
                                                const volatile int32_t* cancel = nullptr) {
    std::vector<Texture2D> frames;
    const int totalFrames = 36;
    for (int angle = 0; angle < 360; angle += 10) {
      if (cancel && *cancel) break;
      const int frameIndex = angle / 10;
      if (progress) progress((float)frameIndex / totalFrames, "Rendering frame " + std::to_string(frameIndex + 1) + "/" + std::to_string(totalFrames) + " (Z=" + std::to_string(angle) + "\xC2\xB0)");
      RenderSettings frameSettings = baseSettings;
      const Vector3 baseRotation = baseSettings.CameraRotationOverride.value_or(Vector3{0, 0, 0});
      frameSettings.CameraRotationOverride = Vector3{baseRotation.x, baseRotation.y, (float)angle};
      auto frame = rayTracer_.RenderAsync(scene_, frameSettings, cancel);
      if (frame) frames.push_back(std::move(*frame));
    }
    return frames;
  }

  // SaveGif :160-184 / SaveGifAsync :82-155: palette mapping on the device, LZW on host threads inside the library.
  void SaveGif(const std::vector<Texture2D>& frames, const std::string& filePath, int frameDelay = 10) {
    if (frames.empty()) return;
    std::vector<const uint8_t*> ptrs;
    for (const Texture2D& f : frames) {
      if (f.width != frames[0].width || f.height != frames[0].height) throw Error(RTB_E_ARG, "frames differ in size");
      ptrs.push_back(f.pixels.data());
    }
    const int rc = rtb_gif_save(rayTracer_.Context(), filePath.c_str(), frames[0].width, frames[0].height, ptrs.data(), (int32_t)ptrs.size(), frameDelay, 0);
    if (rc != RTB_OK) throw Error(rc, rtb_last_error(rayTracer_.Context()));
  }

  // Both halves in one library call: frames leave the GPU as palette indices (1 byte per pixel) and are compressed while the
  // following frames render.
  void RenderRotationGif(const RenderSettings& baseSettings, const std::string& filePath, int frameDelay = 10, int totalFrames = 36, float stepDeg = 10.0f) {
    if (!rayTracer_.EnsureScene(scene_)) throw Error(RTB_E_NOSCENE, "no scene");
    const rtb_render_params p = baseSettings.ToParams();
    const int rc = rtb_gif_render_rotation(rayTracer_.Context(), &p, totalFrames, stepDeg, filePath.c_str(), frameDelay, 0);
    if (rc != RTB_OK) throw Error(rc, rtb_last_error(rayTracer_.Context()));
  }

 private:
  RayTracer& rayTracer_;
  const ObjectData* scene_;
};

// SceneService.cs:26 — a missing file yields an empty ObjectData, as in the reference (:28-33).
struct SceneService {
  static ObjectData LoadScene(const std::string& path) {
    rtb_scene* h = nullptr;
    char err[512] = {0};
    const int rc = rtb_scene_load(path.c_str(), &h, err, sizeof err);
    if (rc == RTB_E_IO) return ObjectData{};
    if (rc != RTB_OK) throw Error(rc, err);
    const rtb_scene_desc& d = *rtb_scene_get(h);
    ObjectData s;
    if (d.has_image) s.Image = ImageSettings{d.image_w, d.image_h, Color{d.bg[0], d.bg[1], d.bg[2]}};
    if (d.has_camera) s.Camera = CameraSettings{d.cam_xform, d.cam_distance, d.cam_vfov_deg};
    for (int i = 0; i < d.n_xforms; i++) {
      CompositeTransformation t;
      for (int k = d.xform_offsets[i]; k < d.xform_offsets[i + 1]; k++) {
        const rtb_xform_elem& e = d.xform_elems[k];
        t.Elements.push_back(TransformElement{(TransformType)e.type, Vector3{e.x, e.y, e.z}, e.angle_deg});
      }
      s.Transformations.push_back(std::move(t));
    }
    for (int i = 0; i < d.n_lights; i++) s.Lights.push_back(LightSource{d.light_xforms[i], Color{d.light_rgb[3 * i], d.light_rgb[3 * i + 1], d.light_rgb[3 * i + 2]}});
    for (int i = 0; i < d.n_materials; i++) { const rtb_material& m = d.materials[i]; s.Materials.push_back(MaterialDescription{Color{m.r, m.g, m.b}, m.ka, m.kd, m.ks, m.kr, m.ior}); }
    for (int i = 0; i < d.n_meshes; i++) {
      TrianglesMesh mesh;
      mesh.transformationIndex = d.meshes[i].xform;
      for (int64_t k = 0; k < d.meshes[i].n_tris; k++) {
        const rtb_triangle& t = d.triangles[d.meshes[i].first_tri + k];
        mesh.Triangles.push_back(Triangle{t.material, Vector3{t.v0[0], t.v0[1], t.v0[2]}, Vector3{t.v1[0], t.v1[1], t.v1[2]}, Vector3{t.v2[0], t.v2[1], t.v2[2]}});
      }
      s.TriangleMeshes.push_back(std::move(mesh));
    }
    for (int i = 0; i < d.n_spheres; i++) s.Spheres.push_back(SphereDescription{d.spheres[i].xform, d.spheres[i].material});
    for (int i = 0; i < d.n_boxes; i++) s.Boxes.push_back(BoxDescription{d.boxes[i].xform, d.boxes[i].material});
    rtb_scene_free(h);
    return s;
  }
};

}  // namespace rtb
