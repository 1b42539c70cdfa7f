// scene_host.hpp — host side of the render path: scene description copy, composite transforms, per-frame uniform resolve,
// the reference-shape (median split) BVH build used by parity mode, and the COSIG scene text parser.
//
// Reference files restated here (paths relative to the reference root):
//   Assets/Services/RayTracer.cs:221-222,238-267,302-355,410-437,455-499   (resolution, matrices, uniforms, materials)
//   Assets/Services/SceneGeometryConverter.cs:83-114,161-190               (BuildMatrix, unit-sphere vertex table)
//   Assets/Services/BVH/BVHBuilder.cs:76-238, AABB.cs:23-39,72              (median-split build + BFS flatten)
//   Assets/Services/SceneService.cs:26-334                                  (scene text format)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rtb.h"
#include "host_math.hpp"
#include "rtb_device.cuh"

namespace rtb {

// Deep copy of an rtb_scene_desc (ObjectData, ObjectData.cs:9-34).  `d` points into the vectors below.
struct HostScene {
  rtb_scene_desc d{};
  std::vector<int32_t> xform_offsets;
  std::vector<rtb_xform_elem> xform_elems;
  std::vector<int32_t> light_xforms;
  std::vector<float> light_rgb;
  std::vector<rtb_material> materials;
  std::vector<rtb_mesh> meshes;
  std::vector<rtb_triangle> triangles;
  std::vector<rtb_prim> spheres, boxes;

  // Returns an empty string on success, else what is wrong with `src`.
  std::string assign(const rtb_scene_desc& src, bool copy_triangles);
  void relink();
};

// BuildComposite, RayTracer.cs:410-437 == BuildMatrix, SceneGeometryConverter.cs:83-114.  Out-of-range index -> identity.
Mat4 composite_matrix(const rtb_scene_desc& s, int index);

// One entry per scene object, in emission order (meshes, boxes, spheres: SceneGeometryConverter.cs:23-48).
enum : int32_t { OBJ_MESH = 0, OBJ_BOX = 1, OBJ_SPHERE = 2,
                 OBJ_BOX_ANALYTIC = 3, OBJ_SPHERE_ANALYTIC = 4 };  // analytic mode: one bounding-box entry per primitive
struct FlattenObject {
  float m[12];        // rows 0..2 of the object's composite matrix (x' = m0*x + m1*y + m2*z + m3)
  float nm[9];        // rows 0..2 of (M^-1)^T, 3x3 part: sphere normals (SceneGeometryConverter.cs:258)
  int32_t kind;
  int32_t material;   // boxes / spheres
  int32_t out_first;  // index of the object's first emitted triangle
  int32_t src_first;  // meshes: index of its first input triangle; analytic primitives: index into the analytic table
  int32_t count;      // emitted triangles
  int32_t pad[2];
};
static_assert(sizeof(FlattenObject) == 112, "FlattenObject layout");

// Builds the object table; returns the total emitted entry count (or -1 if it exceeds int32).  With `analytic`, boxes and
// spheres are not tessellated: each becomes one entry plus a row of `prims` — 24 floats: objectToWorld rows 0..2, then
// worldToObject rows 0..2 (SphereInstance / BoxInstance, Assets/Services/BVH/HittableObjects.cs:16-20, 124-127).
int64_t build_object_table(const rtb_scene_desc& s, bool analytic, std::vector<FlattenObject>& out, std::vector<float>& prims);

// The 402 unit-sphere vertices of AddSphere (SceneGeometryConverter.cs:161-190), xyz per vertex.
const float* unit_sphere_table();  // 402 * 3 floats
constexpr int kSphereVerts = 402;  // (24 + 1) * 16 + 2, SceneGeometryConverter.cs:168
constexpr int kSphereTris = 768;
constexpr int kBoxTris = 12;

// SetupMaterialBuffer, RayTracer.cs:455-499: two float4 per material; an empty list yields the default material.
void pack_materials(const rtb_scene_desc& s, std::vector<float>& out8);

// Everything RayTracer.cs:221-355 resolves from (scene, settings).  Returns false with `err` set on bad parameters.
bool resolve_frame(const rtb_scene_desc& s, const rtb_render_params& p, FrameParams& out, std::string& err);

// --- reference-shape BVH (parity mode) -------------------------------------------------------------------------------
struct RefBvh {
  std::vector<float> nodes;    // 8 floats per node: min.xyz, leftOrFirst (int bits), max.xyz, count (int bits)
  std::vector<int32_t> perm;   // leaf order -> emission index
  int32_t max_leaf = 0;
  int32_t max_depth = 0;
};
// raw: 12 floats per triangle in emission order: v0.xyz, c.x, v1.xyz, c.y, v2.xyz, c.z (the flatten kernel's output).
void build_reference_bvh(const float* raw, int32_t n, RefBvh& out);

// --- scene text parser ------------------------------------------------------------------------------------------------
// SceneService.LoadScene, SceneService.cs:26-242.  Throws std::runtime_error on malformed numbers / truncated files
// (the reference would throw FormatException / IndexOutOfRange).
void parse_scene_text(const char* text, size_t len, HostScene& out);

}  // namespace rtb
