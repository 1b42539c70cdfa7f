// kernels.hpp — host-callable launchers of the library's CUDA kernels (flatten.cu, lbvh.cu, trace.cu).
#pragma once
#include <cuda_runtime.h>

#include "../../include/rtb.h"
#include "rtb_device.cuh"
#include "scene_host.hpp"

namespace rtb {

// ---- flatten.cu (K1: scene upload) ------------------------------------------------------------------------------------
// tri_in: device copy of the caller's rtb_triangle array (10 words each); objs: device FlattenObject table;
// sphere_table: 402 float3 on the device.  Writes, in emission order, raw[3i..] = (v0,c.x)(v1,c.y)(v2,c.z) and
// nrm[3i..] = (n0,material)(n1,0)(n2,0).
void launch_flatten(const float* tri_in, const FlattenObject* objs, int n_objs, const float* sphere_table, int32_t n_out, float4* raw,
                    float4* nrm, cudaStream_t st);
// Gathers into leaf order: isect[RTB_TRI_F4 j..] = (v0,prim_id)(v1-v0,material)(v2-v0,0), shade[3j..] = n0 n1 n2, prim_id = perm[j].
void launch_pack(const float4* raw, const float4* nrm, const int32_t* perm, int32_t n, float4* isect, float4* shade, cudaStream_t st);

// ---- lbvh.cu (K2: Morton-code LBVH built on the GPU) -------------------------------------------------------------------
struct LbvhBuffers {            // all device memory, sized by lbvh_workspace_bytes / allocated by the caller
  float4* nodes;                // (4 or 8, see lbvh_node_f4; wide: 6) * max(1, n-1): worst case; root_out[1] tells how many records were written
  int32_t* perm;                // n
  void* workspace;              // lbvh_workspace_bytes(n)
  size_t workspace_bytes;
  int32_t* root_out;            // 2 ints on the device: root reference, number of node records
};
size_t lbvh_workspace_bytes(int32_t n);
constexpr int lbvh_node_f4 = RTB_LBVH_WIDTH == 4 ? 8 : 4;  // float4 per node record
// Returns cudaSuccess or the failing call's error.  Asynchronous on `st`.
// wide: emit 8-wide quantised records (RTB_WIDE_F4 float4 each, trace.cuh) instead of the binary two-box records; the wide build
// synchronises `st` (it reads its work-list size back between batches of tree levels).
cudaError_t lbvh_build(const float4* raw, int32_t n, const LbvhBuffers& b, cudaStream_t st, bool wide);
int node_record_f4(int bvh);  // float4 per node record of a kernel-side flavour (RTB_BVH_REFERENCE / RTB_BVH_LBVH / RTB_BVH_WIDE)

// ---- trace.cu (K3-K8: the per-pixel wavefront) --------------------------------------------------------------------------
void launch_raygen(int bvh, const FrameParams& f, const SceneView& s, const QueueView& q, const ChunkView& c, int grid, cudaStream_t st);
// `grid` = number of persistent blocks (sm_count * traverse_blocks_per_sm).
// smem_bytes > 0 selects the shared-memory staged variant (grid = sm_count blocks of 1024 threads); it must have been
// enabled for that size with traverse_enable_smem and equal traverse_smem_bytes(bvh, s).
// mode: bit 0 = the closest-hit rays of `depth`, bit 1 = the shadow rays emitted at depth - 1.
void launch_traverse(int bvh, const SceneView& s, const QueueView& q, int depth, int mode, int grid, size_t smem_bytes, cudaStream_t st);
// Packet kernels (one warp walks the BVH once for 32 neighbouring rays).  launch_primary: ray generation + depth-0 traversal
// fused, fills the depth-0 ray queue with the hits only (grid blocks of stream_block_threads() threads, one slot per thread and
// iteration).  launch_packet: kind 0 = closest-hit rays of `depth`, kind 1 = shadow rays emitted at `depth`.
void launch_primary(int bvh, const FrameParams& f, const SceneView& s, const QueueView& q, const ChunkView& c, int grid, cudaStream_t st);
void launch_packet(int bvh, const SceneView& s, const QueueView& q, int depth, int kind, int grid, cudaStream_t st);
int stream_block_threads();
// The regrouping per-lane kernel (k_traverse_pool, RTB_POOL=1): binary LBVH records only.  `scratch` = pool_scratch_bytes(grid)
// bytes of device memory private to the launching stream.
int pool_blocks_per_sm();
size_t pool_scratch_bytes(int grid);
void launch_traverse_pool(const SceneView& s, const QueueView& q, int depth, int mode, int grid, void* scratch, cudaStream_t st);
size_t traverse_smem_bytes(int bvh, const SceneView& s);
cudaError_t traverse_enable_smem(int bvh, size_t bytes);
// k_shade handles depth `depth` when its queue holds >= tail_max rays; otherwise k_tail runs the remaining paths to their
// end in one launch (both are launched every depth: the queue size lives on the device).
void launch_shade(const FrameParams& f, const SceneView& s, const QueueView& q, const ChunkView& c, int depth, int32_t tail_max, int grid,
                  cudaStream_t st);
void launch_tail(int bvh, const FrameParams& f, const SceneView& s, const QueueView& q, const ChunkView& c, int depth, int32_t tail_max, int grid,
                 cudaStream_t st);
void launch_resolve(const FrameParams& f, const QueueView& q, const ChunkView& c, void* dst, int grid, cudaStream_t st);
void launch_debug(int bvh, const FrameParams& f, const SceneView& s, const ChunkView& c, void* dst, int grid, cudaStream_t st);
void launch_aux(int bvh, const FrameParams& f, const SceneView& s, int32_t* prim, float* t, int32_t* mat, int grid, cudaStream_t st);
// ---- group.cu (hand-shake of the process-per-GPU frame ring) ---------------------------------------------------------------
void launch_group_post(uint32_t* flag, uint32_t value, cudaStream_t st);
// waits until flags[i * stride] >= target for i < n (n <= 32); sets *error after timeout_ns
void launch_group_wait(const uint32_t* flags, int n, int stride, uint32_t target, uint32_t* error, unsigned long long timeout_ns, cudaStream_t st);

// Resident blocks per SM of the persistent traversal kernel (occupancy query), by variant.
int traverse_blocks_per_sm(int bvh);

}  // namespace rtb
