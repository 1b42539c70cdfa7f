// rtb_device.cuh — shared device-side definitions: FP32 arithmetic spec helpers, frame uniforms, scene/queue views.
//
// Arithmetic spec: IEEE binary32, one rounding per + - * / sqrt, no contraction.  Every translation unit
// of the library is compiled with --fmad=false --prec-div=true --prec-sqrt=true --ftz=false, so plain expressions are
// never fused.  HLSL intrinsics are fixed as in SURVEY.md App. D: dot = (x*x' + y*y') + z*z',
// normalize(v) = v * (1/sqrt(dot(v,v))), reflect(i,n) = i - (2*dot(n,i))*n, min/max return the non-NaN operand
// (fminf/fmaxf), pow(x,32) = five squarings.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtb {

#define RTB_INFINITY 3.402823466e+38f /* BVHRayTracing.compute:101 */
#define RTB_EPSILON 1e-4f             /* BVHRayTracing.compute:102 */
#define RTB_OFFSET (1e-4f * 100.0f)   /* "Epsilon * 100.0", BVHRayTracing.compute:396,442,447,454 */

#define RTB_STACK_REF 64   /* reference traversal: int stack (the reference has 32 unchecked, compute:235) */
/* float4 per tri_isect record: 3 = packed 48-byte records (three 128-bit loads per test).  4 = 64-byte records, 32-byte
   aligned, fetched with one 256-bit + one 128-bit load: measured 1 % SLOWER on C2-C4 (more bytes, same sectors), kept as a
   build option. */
#ifndef RTB_TRI_F4
#define RTB_TRI_F4 3
#endif
#define RTB_TOTALS 80     /* 8 counters + 16 depths x 4 timeline words (diagnostic builds, RTB_RAY_STATS) + packet counters */
#define RTB_TOT_PACKET_NODES 72 /* node records fetched by the packet kernels (one per warp and visit) */
#define RTB_TOT_ENTERED 74      /* primary rays that entered the BVH (the depth-0 queue) */
#define RTB_TOT_PACKET_TRIS 73  /* triangle records fetched by the packet kernels (one per warp and test) */
#define RTB_STACK_LBVH 96  /* LBVH ordered traversal: one deferred sibling per level */
#ifndef RTB_LEAF_MAX
#define RTB_LEAF_MAX 4     /* LBVH leaf size (<= 8) */
#endif
#define RTB_BVH_WIDE 2      /* kernel-side flavour: RTB_BVH_LBVH with 8-wide quantised node records (RTB_WIDE=1; the default keeps the binary records) */
#define RTB_POOL_SLOTS 64   /* rays a warp of k_traverse_pool keeps in flight */
#define RTB_WIDE_F4 6       /* float4 per 8-wide node record (96 bytes = three 32-byte sectors) */
#define RTB_REF_DONE ((int32_t)0x80000000) /* LBVH traversal: "no more work" reference (never a valid leaf: n < 2^28) */
#define RTB_REF_MISS ((int32_t)0x80000001) /* lbvh_visit: no child of the node was hit (the caller pops its stack) */
#ifndef RTB_LBVH_WIDTH
#define RTB_LBVH_WIDTH 2   /* children per LBVH node record: 2 (64-byte records) or 4 (128-byte records, binary tree collapsed by two levels) */
#endif

struct f3 { float x, y, z; };
__host__ __device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ f3 mk3(float4 v) { return mk3(v.x, v.y, v.z); }
__host__ __device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ f3 negate(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ float dot3(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__host__ __device__ __forceinline__ f3 cross3(f3 a, f3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ f3 hlsl_normalize(f3 v) { const float r = 1.0f / sqrtf(dot3(v, v)); return v * r; }
__device__ __forceinline__ float hlsl_length(f3 v) { return sqrtf(dot3(v, v)); }
__device__ __forceinline__ f3 hlsl_reflect(f3 i, f3 n) { const float k = 2.0f * dot3(n, i); return i - k * n; }
__device__ __forceinline__ float hlsl_frac(float x) { return x - floorf(x); }
__device__ __forceinline__ float pow32(float x) { x = x * x; x = x * x; x = x * x; x = x * x; x = x * x; return x; }
// Vector3.normalized (Unity): v / |v| when |v| > 1e-5, else zero.
__device__ __forceinline__ f3 unity_normalized(f3 v) {
  const float mag = sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
  if (mag > 1e-5f) return mk3(v.x / mag, v.y / mag, v.z / mag);
  return mk3(0.0f, 0.0f, 0.0f);
}

// Per-frame uniforms: what RayTracer.cs:302-355 uploads with ComputeShader.Set*, resolved on the host
// (scene_host.cpp: resolve_frame) and passed by value as a kernel parameter.
struct FrameParams {
  float cam[12];       // _CameraToWorld rows 0..2 (row-major 3x4), RayTracer.cs:352
  float cam_dist;      // _CameraDistance :354
  float tan_half;      // tan(radians(_CameraFOV)*0.5), BVHRayTracing.compute:292, evaluated once on the host
  float ortho_size;    // _OrthoSize :347-348
  float light[3];      // _LightPosition :325-337
  float bg[3];         // _BackgroundColor :322-323
  float light_intensity, light_size, roughness, shutter;
  int32_t width, height;
  int32_t spp, grid_w, grid_h;   // max(1,_AASamples), compute:283-287
  int32_t max_depth;
  int32_t en_ambient, en_diffuse, en_specular, en_refraction, ortho, soft, glossy, blur, debug, srgb;
  // tile sharding: bands of band_rows rows, band b belongs to rank b % band_world
  int32_t band_rank, band_world, band_rows;
  int32_t out_compact;
};

// Rows of the frame owned by rank `rank`: bands rank, rank+world, ...  Local rows are the owned rows packed in order.
__host__ __device__ __forceinline__ int32_t band_local_rows(int32_t height, int32_t rank, int32_t world, int32_t band_rows) {
  if (world <= 1) return height;
  const int32_t n_bands = (height + band_rows - 1) / band_rows;
  int32_t rows = 0;
  // bands owned: rank, rank+world, ... < n_bands
  const int32_t owned = n_bands > rank ? (n_bands - rank + world - 1) / world : 0;
  if (owned == 0) return 0;
  rows = owned * band_rows;
  const int32_t last_band = rank + (owned - 1) * world;
  const int32_t last_end = (last_band + 1) * band_rows;
  if (last_end > height) rows -= last_end - height;
  return rows;
}
__host__ __device__ __forceinline__ int32_t band_global_row(int32_t local_row, int32_t rank, int32_t world, int32_t band_rows) {
  if (world <= 1) return local_row;
  const int32_t lb = local_row / band_rows;
  return (lb * world + rank) * band_rows + (local_row - lb * band_rows);
}

// Work decomposition of one chunk: local rows [row0, row0+rows) of this rank, as 8x4 pixel tiles, spp slots per pixel.
// slot = ((tile * spp + sample) * 32 + lane), tile = ty * tiles_x + tx, lane = ly * 8 + lx.
struct ChunkView {
  int32_t row0, rows;      // local (per-rank) row range; row0 is a multiple of 4
  int32_t tiles_x;         // ceil(width / 8)
  int32_t n_slots;         // tiles_x * ceil(rows/4) * 32 * spp
};

// Geometry as laid out in HBM.
struct SceneView {
  const float4* __restrict__ tri_isect;  // RTB_TRI_F4 per triangle, leaf order: (v0, prim_id) (e1, material) (e2, 0); analytic primitive:
                                         //   (table index bits, -, -, prim_id) (-, -, -, material) (-, -, -, kind: 1 sphere, 2 box)
  const float4* __restrict__ tri_shade;  // 3 per triangle, leaf order: n0 n1 n2
  const float4* __restrict__ nodes;      // reference: 2 per node (min,leftOrFirst)(max,count); LBVH: 4 or 8 per node (see lbvh.cu)
                                         //   LBVH boxes are padded outward (lbvh.cu: k_emit) so the FMA slab test stays conservative
  const float4* __restrict__ materials;  // 2 per material: (r,g,b,ka) (kd,ks,kr,ior)
  const float4* __restrict__ prims;      // analytic mode: 6 per primitive: objectToWorld rows 0..2, worldToObject rows 0..2
  int32_t n_prims;
  int32_t n_tris, n_nodes, n_mats;
  int32_t root;                          // LBVH: root reference (>= 0 internal node, < 0 leaf, see lbvh_leaf_ref)
};

// LBVH child reference: >= 0 internal node index; < 0 leaf: ~ref = (first_triangle << 3) | (count - 1).
__host__ __device__ __forceinline__ int32_t lbvh_leaf_ref(int32_t first, int32_t count) { return ~((first << 3) | (count - 1)); }

// Wavefront queues.
// Ray queue entry (40 B): o = (origin, slot bits), d = (dir, attenuation.x), a = (attenuation.y, attenuation.z).
// Shadow queue entry (56 B): o = (origin, distToLight), d = (dir, slot bits), lit = (lit increment, unlit.x), un = (unlit.y, unlit.z):
// the two candidate increments of the slot's sampleColor (BVHRayTracing.compute:418 evaluated for both outcomes of the shadow test).
// No padding lanes are left: RTB_QUEUE_BYTES_PER_SLOT = 2 * 40 + 56 + 16 (hit record) + 16 (accumulator) = 168 (round 1: 192).
#define RTB_QUEUE_BYTES_PER_SLOT 168
struct QueueView {
  float4* ray_o[2]; float4* ray_d[2]; float2* ray_a[2];
  float4* sh_o; float4* sh_d; float4* sh_lit; float2* sh_un;
  float4* hits;           // per ray-queue entry of the current depth: (t, u, v, leaf-order triangle index bits; -1 = miss)
  float4* accum;          // per slot: running sampleColor
  int32_t* counters;      // RTB_CNT_BLOCKS blocks of depth_cap ints: ray queue sizes | shadow queue sizes | traverse fetch | (spare)
  unsigned long long* totals;  // [0] primary rays [1] continuation rays [2] shadow rays [3] primary hits [4] stack overflows
                               // [5] BVH nodes fetched [6] triangles tested (all rays of the frame)
  int32_t depth_cap;      // D (>= max_depth + 1)
};
#define RTB_CNT_RAY(q, d) ((q).counters[(d)])
#define RTB_CNT_SHADOW(q, d) ((q).counters[(q).depth_cap + (d)])
#define RTB_CNT_FETCH(q, d) ((q).counters[2 * (q).depth_cap + (d)])
#define RTB_CNT_BLOCKS 4

}  // namespace rtb
