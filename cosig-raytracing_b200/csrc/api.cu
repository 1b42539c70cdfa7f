// api.cu — the C ABI of librtb200.so (include/rtb.h): context, scene upload, the per-frame wavefront schedule, readback.
//
// This file is the B200-native counterpart of Assets/Services/RayTracer.cs: rtb_upload_scene = RebuildBVH (:386-404) +
// SetupMaterialBuffer (:455-499); rtb_render = RenderAsync (:212-380); rtb_render_device = RenderToTexture (:82-202).
// There is no CPU fallback: every entry point that computes needs a CUDA device and fails with RTB_E_CUDA otherwise.
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "gif.hpp"
#include "kernels.hpp"

using namespace rtb;

namespace {

thread_local std::string g_create_error;

struct DeviceScene {
  float4 *raw = nullptr, *nrm = nullptr, *isect = nullptr, *shade = nullptr, *nodes = nullptr, *materials = nullptr, *prims = nullptr;
  int32_t n_prims = 0;
  int32_t* perm = nullptr;
  int32_t n_tris = 0, n_nodes = 0, n_mats = 0, root = 0;
  int bvh_mode = RTB_BVH_REFERENCE;
  int flavour = RTB_BVH_REFERENCE;  // what the kernels dispatch on: RTB_BVH_REFERENCE, RTB_BVH_LBVH (binary records) or RTB_BVH_WIDE
  int node_floats = 8;
};

// Device memory of the scene arrays and of the upload's temporaries, kept across rtb_invalidate / rtb_upload_scene: a re-upload
// of a scene that fits does no cudaMalloc and no cudaFree (each of which synchronises the device and, next to gigabytes of live
// wavefront queues, cost tens of milliseconds: round 1 measured 601 ms for rtb_invalidate and 14-16 ms of allocation per upload).
struct Pool {
  void* p = nullptr;
  size_t cap = 0;
};
enum { kPoolMaterials, kPoolPrims, kPoolTriIn, kPoolObjs, kPoolRaw, kPoolNrm, kPoolIsect, kPoolShade, kPoolPerm, kPoolNodes, kPoolWorkspace, kPoolRoot, kPoolCount };
template <typename T>
cudaError_t pool_reserve(Pool& pl, T*& out, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (pl.cap < bytes) {
    if (pl.p) cudaFree(pl.p);
    pl.p = nullptr;
    pl.cap = 0;
    const size_t want = bytes + bytes / 8;  // headroom: a slightly larger scene next time still fits
    cudaError_t e = cudaMalloc(&pl.p, want);
    if (e != cudaSuccess) return e;
    pl.cap = want;
  }
  out = (T*)pl.p;
  return cudaSuccess;
}

struct Queues {
  float4* base = nullptr;      // one allocation: RTB_QUEUE_BYTES_PER_SLOT bytes per slot (queue_view carves it)
  int32_t* counters = nullptr;
  unsigned long long* totals = nullptr;
  int32_t capacity = 0, depth_cap = 0;
};

// A lane = a stream with its own wavefront queues.  Chunks (and, for asynchronous renders, whole frames) alternate between
// the lanes of a device, so the latency-bound tail of one chunk — a few very long rays — overlaps the bulk of the next.
struct LaneState {
  cudaStream_t stream = nullptr;
  Queues q;
  void* pool_scratch = nullptr;   // k_traverse_pool's stack scratch (RTB_POOL=1)
  size_t pool_scratch_bytes = 0;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_done = nullptr;
  uint64_t frame_id = 0;   // last frame that used this lane (its totals belong to that frame)
  bool used = false;       // work was enqueued since the last join
};

struct DeviceState {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;  // == lane[0].stream: uploads, readback, and what rtb_get_stream hands out
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_done = nullptr;  // scene upload timing / cross-device completion
  cudaEvent_t ev_resolve = nullptr;  // orders the resolve kernels of successive chunks / frames (they may share pixels)
  bool resolve_pending = false;
  static constexpr int kMaxLanes = 16;
  LaneState lane[kMaxLanes];
  int next_lane = 0;
  int last_lane = 0;              // lane of the last chunk enqueued (where a frame's readback is ordered)
  cudaStream_t copy_stream = nullptr;  // device->host copies of rtb_render_begin: the lanes go on with the next frames meanwhile
  cudaEvent_t ev_copy = nullptr;
  bool copy_pending = false;
  void* frame_async[kMaxLanes] = {};  // frames of rtb_render_begin, rotating, so a readback never races the following frames
  size_t frame_async_bytes[kMaxLanes] = {};
  struct ChunkRows { int lane, row0, rows; };
  std::vector<ChunkRows> last_chunks;  // chunks of the frame enqueued last (rtb_render copies them back one by one)
  void* gif_scratch = nullptr;        // rtb_gif_index_frame: one RGBA8 frame + its palette indices
  size_t gif_scratch_bytes = 0;
  void* index_async[kMaxLanes] = {};  // palette-index frames of rtb_render_begin_indexed (GIF sweep), same rotation
  size_t index_async_bytes[kMaxLanes] = {};
  float* sphere_table = nullptr;
  DeviceScene scene;
  Pool pool[kPoolCount];  // backing store of `scene` and of the upload's temporaries
  void* frame = nullptr;  // RGBA8 frame (device 0) — also the IPC-exported buffer
  size_t frame_bytes = 0;
  bool frame_exported = false;  // rtb_frame_export handed out an IPC handle: peers may have it mapped, so it must neither move nor go
  int32_t *aux_prim = nullptr, *aux_mat = nullptr;
  float* aux_t = nullptr;
  size_t aux_px = 0;
  int grid_traverse[3] = {0, 0, 0};
  int grid_pool = 0;
  size_t smem_limit = 0;       // opt-in dynamic shared memory per block
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_trace, prof_shadow, prof_resolve;
  size_t prof_used[3] = {0, 0, 0};
};

}  // namespace

struct rtb_context {
  std::vector<DeviceState> devs;
  HostScene host;       // description without the triangle array
  bool has_scene = false;
  int primitive_mode = 0, bvh_mode = 0;
  std::string err;
  rtb_stats stats{};
  const volatile int32_t* cancel = nullptr;
  bool profiling = false;
  int64_t chunk_slots_once = 0;  // rtb_render: chunk size of the frame being enqueued (a blocking frame is split over the lanes), 0 = chunk_slots
  int split_blocking = 1;        // RTB_SPLIT_BLOCKING: 0 = one chunk per blocking frame as long as it fits chunk_slots
  int64_t chunk_slots = 1 << 24;  // RTB_CHUNK_SLOTS: pixel-samples per chunk (RTB_QUEUE_BYTES_PER_SLOT = 168 B of queues each, per lane); C5 sweep: profiles/r1e_sweep_chunk_slots_c5.log
  int n_lanes = 6;            // RTB_LANES (1..8): chunks / async frames rotate over this many streams, each with its own queues (4 -> 6: pipelined e2e +2.5 %, profiles/r2_sweep_lanes.log)
  uint64_t frame_id = 0;
  int smem_mode = 1;          // RTB_SMEM: 1 = stage nodes + triangles in shared memory when they fit (small scenes), 0 = never
  int pool = 0;               // RTB_POOL=1: binary-LBVH scenes in global memory are traversed by the regrouping kernel k_traverse_pool
  int wide = 0;               // RTB_WIDE=1: RTB_BVH_LBVH scenes are stored as 8-wide quantised records instead of binary two-box records (measured slower: profiles/r2_sweep_wide.log)
  int packet_closest = -1;     // RTB_PACKET_CLOSEST: closest-hit rays of depth <= this go through the packet kernels (-1: none, k_raygen + per-lane)
  int packet_shadow = -1;     // RTB_PACKET_SHADOW: shadow rays emitted at depth <= this go through k_packet (-1: none)
  int32_t tail_max = 65536;   // queues smaller than this finish in k_tail (RTB_TAIL_MAX; 0 = pure wavefront); sweep: profiles/r1e_sweep_tail_max.log
  std::vector<void*> ipc_opened;
  // Process-per-GPU frame ring (rtb_group_*): n_buf frame buffers + sequence flags in rank 0's memory, mapped by every rank.
  struct Group {
    bool active = false, owner = false;
    int rank = 0, world = 1, n_buf = 0;
    size_t frame_bytes = 0, stride = 0;
    uint8_t* base = nullptr;      // rank 0: own allocation; other ranks: CUDA-IPC mapping of it
    uint32_t* flags = nullptr;    // behind the frames: stored[world][n_buf], then read_done[n_buf]
    uint32_t* error = nullptr;    // own device word, set by a wait that timed out
    uint64_t seq = 0;             // frames begun
    cudaStream_t gate = nullptr;  // the "buffer is free again" waits run here
    cudaEvent_t gate_done = nullptr;
    static constexpr int kTickets = 16;
    cudaEvent_t ticket[kTickets] = {};
    unsigned long long timeout_ns = 10000000000ull;
    // host ring (rtb_group_create_host): the frames lie in POSIX shared memory that every rank has page-locked; each rank copies
    // its own bands there over its own PCIe link, the flags are words of the same mapping
    bool host = false;
    uint8_t* hbase = nullptr;     // mmap of the shared object: one 4096-byte header, then n_buf frames of `stride` bytes
    uint8_t* hdev = nullptr;      // the same bytes as this device addresses them
    size_t hbytes = 0;
    bool registered = false;
    std::string shm_name;
    bool slot_open[8] = {};       // rank 0: the slot's frame has been begun and not yet ended (its reader has not even seen it)
  } group;
  struct External { cudaExternalMemory_t mem; void* ptr; size_t bytes; };
  std::vector<External> externals;  // rtb_external_import: graphics-API allocations mapped into device 0's address space
  static constexpr int kTickets = 16;  // frames in flight through rtb_render_begin
  cudaEvent_t ticket_event[kTickets] = {};   // frame complete in its host buffer
  cudaEvent_t frame_ready[kTickets] = {};    // frame complete in its device buffer (the copy stream waits for it)
  uint64_t tickets_issued = 0;
};

struct rtb_scene {
  HostScene h;
};

namespace {

int fail(rtb_context* ctx, int code, const std::string& what) {
  if (ctx) ctx->err = what; else g_create_error = what;
  return code;
}
#define CK(ctx, call)                                                                                                  \
  do {                                                                                                                 \
    cudaError_t e__ = (call);                                                                                          \
    if (e__ != cudaSuccess) {                                                                                          \
      std::ostringstream os__;                                                                                         \
      os__ << #call << " failed at " << __FILE__ << ":" << __LINE__ << ": " << cudaGetErrorString(e__);                \
      return fail(ctx, RTB_E_CUDA, os__.str());                                                                        \
    }                                                                                                                  \
  } while (0)

template <typename T>
void dfree(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

// Forgets the scene; its memory stays in the pool for the next upload.
void free_scene(DeviceState& d) { d.scene = DeviceScene(); }

// Returns the pool to the driver (rtb_destroy; rtb_clear_target when no scene is resident).
void release_scene_pool(DeviceState& d) {
  cudaSetDevice(d.device);
  d.scene = DeviceScene();
  for (auto& pl : d.pool) {
    if (pl.p) cudaFree(pl.p);
    pl = Pool();
  }
}

void device_sync(DeviceState& d) {
  cudaSetDevice(d.device);
  for (auto& l : d.lane)
    if (l.stream) cudaStreamSynchronize(l.stream);
  if (d.copy_stream) cudaStreamSynchronize(d.copy_stream);
  for (auto& l : d.lane) l.used = false;
  d.resolve_pending = false;
  d.copy_pending = false;
}

// lane 0's stream (the one the host sees) waits for everything enqueued on the other lanes
cudaError_t join_lanes(DeviceState& d) {
  for (int k = 1; k < DeviceState::kMaxLanes; k++)
    if (d.lane[k].used) {
      cudaError_t e = cudaStreamWaitEvent(d.lane[0].stream, d.lane[k].ev_done, 0);
      if (e != cudaSuccess) return e;
      d.lane[k].used = false;
    }
  if (d.copy_pending) {  // readbacks of rtb_render_begin still in flight on the copy stream
    cudaError_t e = cudaEventRecord(d.ev_copy, d.copy_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(d.lane[0].stream, d.ev_copy, 0);
    if (e != cudaSuccess) return e;
    d.copy_pending = false;
  }
  return cudaSuccess;
}

void free_targets(DeviceState& d) {
  cudaSetDevice(d.device);
  for (auto& l : d.lane) {
    dfree(l.q.base); dfree(l.q.counters); dfree(l.q.totals);
    l.q = Queues();
    dfree(l.pool_scratch); l.pool_scratch_bytes = 0;
  }
  dfree(d.frame); d.frame_bytes = 0; d.frame_exported = false;
  for (int k = 0; k < DeviceState::kMaxLanes; k++) { dfree(d.frame_async[k]); d.frame_async_bytes[k] = 0; }
  for (int k = 0; k < DeviceState::kMaxLanes; k++) { dfree(d.index_async[k]); d.index_async_bytes[k] = 0; }
  dfree(d.gif_scratch); d.gif_scratch_bytes = 0;
  dfree(d.aux_prim); dfree(d.aux_mat); dfree(d.aux_t); d.aux_px = 0;
}

int ensure_queues(rtb_context* ctx, LaneState& l, int32_t capacity, int32_t depth_cap) {
  if (l.q.capacity >= capacity && l.q.depth_cap >= depth_cap) return RTB_OK;
  CK(ctx, cudaStreamSynchronize(l.stream));
  dfree(l.q.base); dfree(l.q.counters); dfree(l.q.totals);
  l.q = Queues();
  CK(ctx, cudaMalloc(&l.q.base, (size_t)capacity * RTB_QUEUE_BYTES_PER_SLOT));
  CK(ctx, cudaMalloc(&l.q.counters, (size_t)depth_cap * RTB_CNT_BLOCKS * sizeof(int32_t)));
  CK(ctx, cudaMalloc(&l.q.totals, RTB_TOTALS * sizeof(unsigned long long)));
  l.q.capacity = capacity;
  l.q.depth_cap = depth_cap;
  l.frame_id = 0;
  return RTB_OK;
}

QueueView queue_view(const Queues& q) {
  QueueView v;
  float4* p = q.base;
  const size_t c = (size_t)q.capacity;
  // nine float4 arrays, then three float2 arrays (capacity is a multiple of 32, so everything stays 16-byte aligned)
  v.ray_o[0] = p; v.ray_o[1] = p + c; v.ray_d[0] = p + 2 * c; v.ray_d[1] = p + 3 * c;
  v.sh_o = p + 4 * c; v.sh_d = p + 5 * c; v.sh_lit = p + 6 * c;
  v.hits = p + 7 * c;
  v.accum = p + 8 * c;
  float2* h = (float2*)(p + 9 * c);
  v.ray_a[0] = h; v.ray_a[1] = h + c; v.sh_un = h + 2 * c;
  v.counters = q.counters;
  v.totals = q.totals;
  v.depth_cap = q.depth_cap;
  return v;
}

SceneView scene_view(const DeviceScene& s) {
  SceneView v;
  v.tri_isect = s.isect; v.tri_shade = s.shade; v.nodes = s.nodes; v.materials = s.materials;
  v.n_tris = s.n_tris; v.n_nodes = s.n_nodes; v.n_mats = s.n_mats; v.root = s.root;
  v.prims = s.prims; v.n_prims = s.n_prims;
  return v;
}

int ensure_frame(rtb_context* ctx, DeviceState& d, size_t bytes) {
  if (d.frame_bytes >= bytes) return RTB_OK;
  if (d.frame_exported)
    return fail(ctx, RTB_E_ARG, "the context's frame buffer is exported over IPC and too small for this frame: export a larger one first (peers would keep storing into freed memory)");
  dfree(d.frame);
  d.frame_bytes = 0;
  CK(ctx, cudaMalloc(&d.frame, bytes));
  d.frame_bytes = bytes;
  return RTB_OK;
}

void prof_pair(DeviceState& d, int family, cudaEvent_t& a, cudaEvent_t& b) {
  auto& pool = family == 0 ? d.prof_trace : (family == 1 ? d.prof_shadow : d.prof_resolve);
  size_t& used = d.prof_used[family];
  if (used == pool.size()) {
    cudaEvent_t x, y;
    cudaEventCreate(&x);
    cudaEventCreate(&y);
    pool.emplace_back(x, y);
  }
  a = pool[used].first;
  b = pool[used].second;
  used++;
}

// ---------------------------------------------------------------------------------------------------------------------
// Scene upload on one device
// ---------------------------------------------------------------------------------------------------------------------
int upload_on_device(rtb_context* ctx, DeviceState& d, const rtb_scene_desc& desc, const std::vector<FlattenObject>& objs, int32_t n_out,
                     const std::vector<float>& mats, const std::vector<float>& prims, int bvh_mode, float& ms_build) {
  const bool wide = bvh_mode == RTB_BVH_LBVH && ctx->wide != 0;
  CK(ctx, cudaSetDevice(d.device));
  device_sync(d);
  free_scene(d);
  DeviceScene& s = d.scene;
  s.bvh_mode = bvh_mode;
  s.flavour = wide ? RTB_BVH_WIDE : bvh_mode;
  s.n_tris = n_out;
  s.n_mats = (int32_t)(mats.size() / 8);
  CK(ctx, pool_reserve(d.pool[kPoolMaterials], s.materials, mats.size() * sizeof(float)));
  CK(ctx, cudaMemcpyAsync(s.materials, mats.data(), mats.size() * sizeof(float), cudaMemcpyHostToDevice, d.stream));
  if (n_out == 0) { CK(ctx, cudaStreamSynchronize(d.stream)); return RTB_OK; }
  s.n_prims = (int32_t)(prims.size() / 24);
  if (s.n_prims > 0) {
    CK(ctx, pool_reserve(d.pool[kPoolPrims], s.prims, prims.size() * sizeof(float)));
    CK(ctx, cudaMemcpyAsync(s.prims, prims.data(), prims.size() * sizeof(float), cudaMemcpyHostToDevice, d.stream));
  }

  float* tri_in = nullptr;
  FlattenObject* dobjs = nullptr;
  if (desc.n_triangles > 0) {
    CK(ctx, pool_reserve(d.pool[kPoolTriIn], tri_in, (size_t)desc.n_triangles * sizeof(rtb_triangle)));
    CK(ctx, cudaMemcpyAsync(tri_in, desc.triangles, (size_t)desc.n_triangles * sizeof(rtb_triangle), cudaMemcpyHostToDevice, d.stream));
  }
  CK(ctx, pool_reserve(d.pool[kPoolObjs], dobjs, objs.size() * sizeof(FlattenObject)));
  CK(ctx, cudaMemcpyAsync(dobjs, objs.data(), objs.size() * sizeof(FlattenObject), cudaMemcpyHostToDevice, d.stream));
  const size_t tri_bytes = (size_t)n_out * 3 * sizeof(float4);
  CK(ctx, pool_reserve(d.pool[kPoolRaw], s.raw, tri_bytes));
  CK(ctx, pool_reserve(d.pool[kPoolNrm], s.nrm, tri_bytes));
  CK(ctx, pool_reserve(d.pool[kPoolIsect], s.isect, (size_t)n_out * RTB_TRI_F4 * sizeof(float4)));
  CK(ctx, pool_reserve(d.pool[kPoolShade], s.shade, tri_bytes));
  CK(ctx, pool_reserve(d.pool[kPoolPerm], s.perm, (size_t)n_out * sizeof(int32_t)));

  CK(ctx, cudaEventRecord(d.ev_begin, d.stream));
  launch_flatten(tri_in, dobjs, (int)objs.size(), d.sphere_table, n_out, s.raw, s.nrm, d.stream);
  CK(ctx, cudaGetLastError());

  if (bvh_mode == RTB_BVH_REFERENCE) {
    // Parity mode: the reference's own tree shape (BVHBuilder.cs) decides tie winners, so it is rebuilt on the host from
    // the device-flattened triangles and uploaded.
    std::vector<float> raw_host((size_t)n_out * 12);
    CK(ctx, cudaMemcpyAsync(raw_host.data(), s.raw, tri_bytes, cudaMemcpyDeviceToHost, d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));
    RefBvh bvh;
    build_reference_bvh(raw_host.data(), n_out, bvh);
    s.n_nodes = (int32_t)(bvh.nodes.size() / 8);
    s.node_floats = 8;
    CK(ctx, pool_reserve(d.pool[kPoolNodes], s.nodes, bvh.nodes.size() * sizeof(float)));
    CK(ctx, cudaMemcpyAsync(s.nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(float), cudaMemcpyHostToDevice, d.stream));
    CK(ctx, cudaMemcpyAsync(s.perm, bvh.perm.data(), (size_t)n_out * sizeof(int32_t), cudaMemcpyHostToDevice, d.stream));
    launch_pack(s.raw, s.nrm, s.perm, n_out, s.isect, s.shade, d.stream);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaEventRecord(d.ev_end, d.stream));
    CK(ctx, cudaStreamSynchronize(d.stream));  // bvh vectors go out of scope
  } else {
    // The records are built in place in a worst-case sized array (n - 1 records; the binary tree uses about a third of them,
    // the rest is never touched): no second allocation, no copy.
    const int32_t max_nodes = n_out > 1 ? n_out - 1 : 1;
    const int rec_f4 = node_record_f4(s.flavour);
    s.node_floats = 4 * rec_f4;
    CK(ctx, pool_reserve(d.pool[kPoolNodes], s.nodes, (size_t)max_nodes * rec_f4 * sizeof(float4)));
    LbvhBuffers b;
    b.nodes = s.nodes;
    b.perm = s.perm;
    b.workspace_bytes = lbvh_workspace_bytes(n_out);
    void* ws = nullptr;
    int32_t* root_dev = nullptr;
    int32_t root_host[2] = {0, 0};
    CK(ctx, pool_reserve(d.pool[kPoolWorkspace], ws, b.workspace_bytes));
    CK(ctx, pool_reserve(d.pool[kPoolRoot], root_dev, 2 * sizeof(int32_t)));
    b.workspace = ws;
    b.root_out = root_dev;
    cudaError_t e = lbvh_build(s.raw, n_out, b, d.stream, wide);
    if (e == cudaSuccess) { launch_pack(s.raw, s.nrm, s.perm, n_out, s.isect, s.shade, d.stream); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(root_host, root_dev, sizeof root_host, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess) e = cudaEventRecord(d.ev_end, d.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    if (e != cudaSuccess) return fail(ctx, RTB_E_CUDA, std::string("LBVH build failed: ") + cudaGetErrorString(e));
    s.root = root_host[0];
    s.n_nodes = std::max(1, std::min(root_host[1], max_nodes));
  }
  CK(ctx, cudaEventElapsedTime(&ms_build, d.ev_begin, d.ev_end));
  return RTB_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// One frame on one device: chunks of <= chunk_slots pixel-samples, each a fixed sequence of persistent kernels.
// `dst` is a device pointer valid on d.device (local memory or a peer/IPC mapping).
// ---------------------------------------------------------------------------------------------------------------------
int render_on_device(rtb_context* ctx, DeviceState& d, const FrameParams& f, void* dst, int& launches, int& chunks) {
  CK(ctx, cudaSetDevice(d.device));
  const int bvh = d.scene.flavour;
  const SceneView sv = scene_view(d.scene);
  const int32_t local_rows = band_local_rows(f.height, f.band_rank, f.band_world, f.band_rows);
  const int32_t tiles_x = (f.width + 7) / 8;
  const int64_t slots_per_tile_row = (int64_t)tiles_x * 32 * f.spp;  // 4 pixel rows
  int64_t tile_rows_per_chunk = (ctx->chunk_slots_once > 0 ? std::min(ctx->chunk_slots_once, ctx->chunk_slots) : ctx->chunk_slots) / slots_per_tile_row;
  if (tile_rows_per_chunk < 1) tile_rows_per_chunk = 1;
  if (tile_rows_per_chunk * slots_per_tile_row > (int64_t)INT32_MAX - 64) return fail(ctx, RTB_E_ARG, "a 4-row strip of this frame exceeds 2^31 samples");
  const int32_t rows_per_chunk = (int32_t)std::min<int64_t>(tile_rows_per_chunk * 4, ((int64_t)local_rows + 3) / 4 * 4);
  const int32_t capacity = (int32_t)(((int64_t)rows_per_chunk / 4) * slots_per_tile_row);
  const int32_t depth_cap = f.max_depth + 2;
  if (!d.grid_traverse[bvh]) d.grid_traverse[bvh] = d.sm_count * traverse_blocks_per_sm(bvh);
  const int shade_grid = d.sm_count * 8;
  size_t smem_bytes = 0;  // small scenes: k_traverse works out of a shared-memory copy of nodes + triangles
  if (ctx->smem_mode != 0 && d.scene.n_tris > 0 && d.scene.n_prims == 0) {
    const size_t need = traverse_smem_bytes(bvh, sv);
    if (need + 1024 <= d.smem_limit) {
      CK(ctx, traverse_enable_smem(bvh, need));
      smem_bytes = need;
    }
  }
  d.prof_used[0] = d.prof_used[1] = d.prof_used[2] = 0;
  d.last_chunks.clear();
  for (int32_t row0 = 0; row0 < local_rows; row0 += rows_per_chunk) {
    LaneState& L = d.lane[d.next_lane];
    d.last_lane = d.next_lane;
    d.last_chunks.push_back({d.next_lane, row0, std::min(rows_per_chunk, local_rows - row0)});
    d.next_lane = ctx->profiling ? d.next_lane : (d.next_lane + 1) % ctx->n_lanes;  // profiling: one lane, so per-launch event intervals do not overlap
    cudaStream_t stream = L.stream;
    if (ctx->cancel && *ctx->cancel) { device_sync(d); return fail(ctx, RTB_E_CANCELLED, "cancelled"); }
    {
      const int rc = ensure_queues(ctx, L, capacity, depth_cap);
      if (rc != RTB_OK) return rc;
    }
    const QueueView qv = queue_view(L.q);
    const bool use_pool = ctx->pool != 0 && bvh == RTB_BVH_LBVH && smem_bytes == 0;
    int pool_grid = 0;
    if (use_pool) {
      if (!d.grid_pool) d.grid_pool = d.sm_count * pool_blocks_per_sm();
      pool_grid = d.grid_pool;
      const size_t need = pool_scratch_bytes(pool_grid);
      if (L.pool_scratch_bytes < need) {
        CK(ctx, cudaStreamSynchronize(stream));
        dfree(L.pool_scratch);
        L.pool_scratch_bytes = 0;
        CK(ctx, cudaMalloc(&L.pool_scratch, need));
        L.pool_scratch_bytes = need;
      }
    }
    if (L.frame_id != ctx->frame_id) {  // first chunk of this frame on this lane
      L.frame_id = ctx->frame_id;
      CK(ctx, cudaEventRecord(L.ev_begin, stream));
      CK(ctx, cudaMemsetAsync(L.q.totals, 0, RTB_TOTALS * sizeof(unsigned long long), stream));
    }
    L.used = true;
    auto timed = [&](int family, auto&& launch) {
      cudaEvent_t a = nullptr, b = nullptr;
      if (ctx->profiling) { prof_pair(d, family, a, b); cudaEventRecord(a, stream); }
      launch();
      if (ctx->profiling) cudaEventRecord(b, stream);
      launches++;
    };
    // successive resolves may write the same pixels (frames in flight on several lanes): keep them in issue order
    auto ordered_resolve = [&](auto&& launch) -> cudaError_t {
      cudaError_t e = cudaSuccess;
      if (d.resolve_pending) e = cudaStreamWaitEvent(stream, d.ev_resolve, 0);
      if (e != cudaSuccess) return e;
      timed(2, launch);
      d.resolve_pending = true;
      return cudaEventRecord(d.ev_resolve, stream);
    };
    ChunkView c;
    c.row0 = row0;
    c.rows = std::min(rows_per_chunk, local_rows - row0);
    c.tiles_x = tiles_x;
    c.n_slots = (int32_t)((int64_t)((c.rows + 3) / 4) * slots_per_tile_row);
    chunks++;
    const int resolve_grid = std::min(d.sm_count * 8, (c.rows * f.width + 255) / 256);
    if (f.debug != 0) {
      CK(ctx, ordered_resolve([&] { launch_debug(bvh, f, sv, c, dst, resolve_grid, stream); }));
    } else {
      if (f.max_depth <= 0) {  // the depth loop never runs: sampleColor stays 0
        CK(ctx, cudaMemsetAsync(qv.accum, 0, (size_t)c.n_slots * sizeof(float4), stream));
      } else {
        CK(ctx, cudaMemsetAsync(L.q.counters, 0, (size_t)L.q.depth_cap * RTB_CNT_BLOCKS * sizeof(int32_t), stream));
        // raygen fills the depth-0 queue; depth d: traverse (closest-hit rays of depth d + shadow rays emitted at depth d-1),
        // then shade (or, for short queues, k_tail).  One more traverse at the end serves the last depth's shadow rays.
        // Optional packet kernels (RTB_PACKET_CLOSEST / RTB_PACKET_SHADOW >= 0; off by default): depth 0 is coherent (8x4-pixel tiles),
        // and so are the shadow rays those pixels emit; k_primary = raygen + traversal fused, k_packet = 32 queue entries per walk.
        // Measured (profiles/r2_sweep_packet.log): at 4K over 1 M triangles a tile's rays share the top of the tree only — its union
        // of nodes is ~3x one ray's path — so packets lose 8-60 % there; they remain for scenes of few large triangles.
        const bool packets = bvh != RTB_BVH_WIDE;  // the packet kernels walk the reference tree and the binary LBVH records
        const int pk_closest = packets ? ctx->packet_closest : -1, pk_shadow = packets ? ctx->packet_shadow : -1;
        const int packet_grid = d.sm_count * 8;
        if (pk_closest >= 0) {
          const int tb = stream_block_threads();
          const int primary_grid = (int)std::min<int64_t>(((int64_t)c.n_slots + tb - 1) / tb, (int64_t)1 << 20);
          timed(0, [&] { launch_primary(bvh, f, sv, qv, c, std::max(primary_grid, 1), stream); });
        } else {
          timed(1, [&] { launch_raygen(bvh, f, sv, qv, c, shade_grid, stream); });
        }
        for (int depth = 0; depth <= f.max_depth; depth++) {
          if (depth > 0 && ctx->cancel && *ctx->cancel) { device_sync(d); return fail(ctx, RTB_E_CANCELLED, "cancelled"); }
          int mode = 0;
          if (depth < f.max_depth && !(depth == 0 && pk_closest >= 0)) {
            if (depth > 0 && depth <= pk_closest) timed(0, [&] { launch_packet(bvh, sv, qv, depth, 0, packet_grid, stream); });
            else mode |= 1;
          }
          if (depth > 0 && f.en_diffuse == 1) {
            if (depth - 1 <= pk_shadow) timed(0, [&] { launch_packet(bvh, sv, qv, depth - 1, 1, packet_grid, stream); });
            else mode |= 2;
          }
          if (mode != 0) {
            if (use_pool) timed(0, [&] { launch_traverse_pool(sv, qv, depth, mode, pool_grid, L.pool_scratch, stream); });
            else timed(0, [&] { launch_traverse(bvh, sv, qv, depth, mode, smem_bytes ? d.sm_count : d.grid_traverse[bvh], smem_bytes, stream); });
          }
          if (depth < f.max_depth) {
            timed(1, [&] { launch_shade(f, sv, qv, c, depth, ctx->tail_max, shade_grid, stream); });
            if (ctx->tail_max > 0) timed(0, [&] { launch_tail(bvh, f, sv, qv, c, depth, ctx->tail_max, d.sm_count * 4, stream); });
          }
        }
      }
      CK(ctx, ordered_resolve([&] { launch_resolve(f, qv, c, dst, resolve_grid, stream); }));
    }
    CK(ctx, cudaEventRecord(L.ev_end, stream));
    CK(ctx, cudaEventRecord(L.ev_done, stream));
  }
  CK(ctx, cudaGetLastError());
  return RTB_OK;
}

float sum_pairs(std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& pool, size_t used) {
  float total = 0.0f;
  for (size_t i = 0; i < used; i++) {
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, pool[i].first, pool[i].second) == cudaSuccess) total += ms;
  }
  return total;
}

// Renders the frame over the context's devices into `dst` (memory of device 0, or a peer mapping) and waits.
int render_frame(rtb_context* ctx, const rtb_render_params* p, void* dst, size_t dst_bytes, bool to_internal_frame, bool sync, FrameParams& f_out) {
  if (!ctx || !p) return fail(ctx, RTB_E_ARG, "null argument");
  if (!ctx->has_scene) return fail(ctx, RTB_E_NOSCENE, "no scene uploaded (rtb_upload_scene)");
  FrameParams f;
  std::string why;
  if (!resolve_frame(ctx->host.d, *p, f, why)) return fail(ctx, RTB_E_ARG, why);
  const int n_dev = (int)ctx->devs.size();
  if (n_dev > 1 && f.band_world > 1) return fail(ctx, RTB_E_ARG, "band sharding across processes needs a single-device context");
  f_out = f;
  if (to_internal_frame) {
    const size_t need = (size_t)f.width * f.height * 4;
    const int rc = ensure_frame(ctx, ctx->devs[0], need);
    if (rc != RTB_OK) return rc;
    dst = ctx->devs[0].frame;
    dst_bytes = ctx->devs[0].frame_bytes;
    f.out_compact = 0;
  }
  const size_t rows_out = f.out_compact ? (size_t)band_local_rows(f.height, f.band_rank, f.band_world, f.band_rows) : (size_t)f.height;
  if (!dst || dst_bytes < rows_out * (size_t)f.width * 4) return fail(ctx, RTB_E_SIZE, "output buffer too small for the resolved resolution");

  int launches = 0, chunks = 0;
  ctx->frame_id++;
  for (int k = 0; k < n_dev; k++) {
    FrameParams fk = f;
    if (n_dev > 1) { fk.band_world = n_dev; fk.band_rank = k; fk.out_compact = 0; }
    const int rc = render_on_device(ctx, ctx->devs[(size_t)k], fk, dst, launches, chunks);
    if (rc != RTB_OK) return rc;
  }
  // device 0 waits for the peers' stores (the NVLink gather is complete when their resolve kernels have finished)
  for (int k = 1; k < n_dev; k++) {
    DeviceState& dk = ctx->devs[(size_t)k];
    CK(ctx, cudaSetDevice(dk.device));
    CK(ctx, join_lanes(dk));
    CK(ctx, cudaEventRecord(dk.ev_done, dk.stream));
    CK(ctx, cudaSetDevice(ctx->devs[0].device));
    CK(ctx, cudaStreamWaitEvent(ctx->devs[0].stream, dk.ev_done, 0));
  }
  CK(ctx, cudaSetDevice(ctx->devs[0].device));
  rtb_stats& st = ctx->stats;
  st.width = f.width; st.height = f.height; st.spp = f.spp; st.chunks = chunks; st.kernel_launches = launches; st.n_devices = n_dev;
  st.h2d_bytes = (int64_t)sizeof(FrameParams) * launches;  // uniforms travel as kernel parameters
  st.d2h_bytes = 0;
  if (sync) {
    for (int k = 0; k < n_dev; k++) device_sync(ctx->devs[(size_t)k]);
    CK(ctx, cudaSetDevice(ctx->devs[0].device));
    CK(ctx, cudaGetLastError());
  }
  return RTB_OK;
}

int collect_stats(rtb_context* ctx) {
  rtb_stats& st = ctx->stats;
  st.rays_primary = st.rays_continuation = st.rays_shadow = st.paths_hit_primary = 0;
  st.ms_render_device = 0.0f;
  st.ms_traverse = st.ms_shade = st.ms_resolve = 0.0f;
  int64_t overflow = 0, nodes = 0, tris = 0, longest = 0, entered = 0, pk_nodes = 0, pk_tris = 0;
  for (auto& d : ctx->devs) {
    device_sync(d);
    CK(ctx, cudaGetLastError());
    for (auto& l : d.lane) {
      if (!l.q.totals || l.frame_id != ctx->frame_id || ctx->frame_id == 0) continue;
      unsigned long long t[RTB_TOTALS];
      CK(ctx, cudaMemcpy(t, l.q.totals, sizeof t, cudaMemcpyDeviceToHost));
      if (std::getenv("RTB_TIMELINE")) {  // diagnostic builds (-DRTB_RAY_STATS=1): per k_traverse launch, ns since its first block started
        for (int dd = 0; dd < 16; dd++) {
          const unsigned long long* w = t + 8 + 4 * dd;
          if (!w[0]) continue;
          const unsigned long long start = ~w[0];
          std::fprintf(stderr, "k_traverse depth %d: queue exhausted first at %.1f us, last warp saw it at %.1f us, last warp done at %.1f us\n", dd,
                       (~w[1] - start) * 1e-3, (w[3] - start) * 1e-3, (w[2] - start) * 1e-3);
        }
      }
      st.rays_primary += (int64_t)t[0]; st.rays_continuation += (int64_t)t[1]; st.rays_shadow += (int64_t)t[2];
      st.paths_hit_primary += (int64_t)t[3]; overflow += (int64_t)t[4]; nodes += (int64_t)t[5]; tris += (int64_t)t[6];
      longest = std::max(longest, (int64_t)t[7]);
      entered += (int64_t)t[RTB_TOT_ENTERED]; pk_nodes += (int64_t)t[RTB_TOT_PACKET_NODES]; pk_tris += (int64_t)t[RTB_TOT_PACKET_TRIS];
      float ms = 0.0f;
      if (cudaEventElapsedTime(&ms, l.ev_begin, l.ev_end) == cudaSuccess) st.ms_render_device = std::max(st.ms_render_device, ms);
      else cudaGetLastError();
    }
    if (ctx->profiling && &d == &ctx->devs[0]) {
      st.ms_traverse = sum_pairs(d.prof_trace, d.prof_used[0]);
      st.ms_shade = sum_pairs(d.prof_shadow, d.prof_used[1]);
      st.ms_resolve = sum_pairs(d.prof_resolve, d.prof_used[2]);
    }
  }
  st.reserved[0] = overflow;
  st.reserved[1] = nodes;
  st.reserved[2] = tris;
  st.reserved[3] = longest;
  st.rays_traversed = entered + st.rays_continuation + st.rays_shadow;
  st.packet_node_fetches = pk_nodes;
  st.packet_tri_fetches = pk_tris;
  st.bytes_per_slot = RTB_QUEUE_BYTES_PER_SLOT;
  cudaSetDevice(ctx->devs[0].device);
  return RTB_OK;
}

}  // namespace

namespace {
// Header of the host ring (first 4096 bytes of the shared object).  Every field is written by exactly one party: rank 0 writes
// the shape once (magic last) and `begun0`; `done[rank * 8 + buffer]` is written by that rank's device (group.cu: k_group_post,
// ordered behind the copy of its bands on its copy stream).  Sequence numbers are compared as signed differences.
struct HostRingHeader {
  uint32_t magic, world, n_buf, reserved;
  uint64_t frame_bytes, stride;
  alignas(64) uint32_t begun0;     // frames rank 0 has begun: slot k % n_buf may be overwritten with frame k once begun0 > k
  alignas(64) uint32_t done[32 * 8];
};
static_assert(sizeof(HostRingHeader) <= 4096, "the header is one page");
constexpr uint32_t kHostRingMagic = 0x52544248u;  // "RTBH"
constexpr size_t kHostRingHeaderBytes = 4096;

unsigned long long monotonic_ns() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}

// Host-side wait on a word of the shared header: spins briefly, then yields; false after timeout_ns.
bool host_wait_reached(const uint32_t* word, uint32_t target, unsigned long long timeout_ns) {
  const unsigned long long t0 = monotonic_ns();
  for (unsigned spins = 0;; spins++) {
    if ((int32_t)(__atomic_load_n(word, __ATOMIC_ACQUIRE) - target) >= 0) return true;
    if (spins > 2000) {
      if (monotonic_ns() - t0 > timeout_ns) return false;
      timespec nap = {0, 20000};
      nanosleep(&nap, nullptr);
    }
  }
}

void group_release(rtb_context* ctx) {
  rtb_context::Group& g = ctx->group;
  if (!g.active) return;
  cudaSetDevice(ctx->devs[0].device);
  for (auto& d : ctx->devs) device_sync(d);
  if (g.gate) { cudaStreamSynchronize(g.gate); cudaStreamDestroy(g.gate); }
  if (g.gate_done) cudaEventDestroy(g.gate_done);
  for (auto& e : g.ticket) if (e) cudaEventDestroy(e);
  if (g.base) { if (g.owner) cudaFree(g.base); else cudaIpcCloseMemHandle(g.base); }
  if (g.error) cudaFree(g.error);
  if (g.hbase) {
    if (g.registered) cudaHostUnregister(g.hbase);
    munmap(g.hbase, g.hbytes);
  }
  if (g.owner && !g.shm_name.empty()) shm_unlink(g.shm_name.c_str());  // also when creation failed half-way (nothing mapped yet)
  cudaGetLastError();
  g = rtb_context::Group();
}

// Host-ring form of rtb_group_render_begin: render this rank's bands into a device frame buffer of its own, copy exactly those
// bands into the frame's slot of the shared host ring (one strided copy: the bands of a rank are band_rows * width * 4 bytes every
// world * that many), post done[rank][slot] behind the copy.  No rank touches another rank's GPU or PCIe link.
int group_host_begin(rtb_context* ctx, const rtb_render_params* p, uint8_t* rgba8, int32_t* ticket) {
  rtb_context::Group& g = ctx->group;
  if (rgba8) return fail(ctx, RTB_E_ARG, "a host-ring group delivers frames in the ring (rtb_group_frame): pass rgba8 = NULL");
  if (!g.hbase || !g.hdev) return fail(ctx, RTB_E_ARG, "no group (rtb_group_create_host)");
  rtb_render_params pp = *p;
  pp.band_rank = g.rank; pp.band_world = g.world; pp.out_layout = RTB_OUT_FRAME;
  if (pp.band_rows <= 0) pp.band_rows = 8;
  FrameParams f;
  std::string why;
  if (!resolve_frame(ctx->host.d, pp, f, why)) return fail(ctx, RTB_E_ARG, why);
  const size_t row_bytes = (size_t)f.width * 4, need = row_bytes * (size_t)f.height;
  if (need > g.frame_bytes) return fail(ctx, RTB_E_SIZE, "frame larger than the group's buffers");
  DeviceState& d = ctx->devs[0];
  CK(ctx, cudaSetDevice(d.device));
  HostRingHeader* hdr = (HostRingHeader*)g.hbase;
  const uint64_t k = g.seq;
  const int j = (int)(k % (uint64_t)g.n_buf), slot = (int)(k % rtb_context::Group::kTickets);
  // slot j still holds frame k - n_buf until rank 0's caller has moved on, which it says by beginning frame k
  if (g.rank == 0 && g.slot_open[j]) return fail(ctx, RTB_E_ARG, "the ring is full: end the frame begun n_buffers calls ago (and read it) before beginning another");
  if (g.rank == 0) __atomic_store_n(&hdr->begun0, (uint32_t)(k + 1), __ATOMIC_RELEASE);
  else if (!host_wait_reached(&hdr->begun0, (uint32_t)(k + 1), g.timeout_ns))
    return fail(ctx, RTB_E_CUDA, "rank 0 of the group did not begin this frame in time (RTB_GROUP_TIMEOUT_MS)");
  if (!g.ticket[slot]) CK(ctx, cudaEventCreateWithFlags(&g.ticket[slot], cudaEventDisableTiming));
  else CK(ctx, cudaEventSynchronize(g.ticket[slot]));  // the ring of tickets is full: wait for the frame begun kTickets calls ago
  const int n_dev_buf = std::max(2, ctx->n_lanes), buf = (int)(k % (uint64_t)n_dev_buf);
  if (d.frame_async_bytes[buf] < need) {
    device_sync(d);
    dfree(d.frame_async[buf]);
    d.frame_async_bytes[buf] = 0;
    CK(ctx, cudaMalloc(&d.frame_async[buf], need));
    d.frame_async_bytes[buf] = need;
  }
  if (k >= (uint64_t)n_dev_buf) {  // the device buffer's previous frame must have left it
    cudaEvent_t prev = g.ticket[(k - (uint64_t)n_dev_buf) % rtb_context::Group::kTickets];
    for (auto& l : d.lane) if (l.stream) CK(ctx, cudaStreamWaitEvent(l.stream, prev, 0));
  }
  FrameParams f2;
  const int rc = render_frame(ctx, &pp, d.frame_async[buf], d.frame_async_bytes[buf], /*to_internal_frame=*/false, /*sync=*/false, f2);
  if (rc != RTB_OK) return rc;
  LaneState& last = d.lane[d.last_lane];
  for (int l = 0; l < DeviceState::kMaxLanes; l++)  // multi-chunk frame: every lane that carried one of its chunks
    if (l != d.last_lane && d.lane[l].stream && d.lane[l].frame_id == ctx->frame_id) CK(ctx, cudaStreamWaitEvent(last.stream, d.lane[l].ev_done, 0));
  CK(ctx, cudaEventRecord(last.ev_done, last.stream));
  last.used = true;
  CK(ctx, cudaStreamWaitEvent(d.copy_stream, last.ev_done, 0));
  // ReadPixels (RayTracer.cs:371-375) of this rank's rows only
  const uint8_t* src = (const uint8_t*)d.frame_async[buf];
  uint8_t* dst = g.hbase + kHostRingHeaderBytes + (size_t)j * g.stride;
  size_t copied = 0;
  if (g.world == 1) {
    CK(ctx, cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, d.copy_stream));
    copied = need;
  } else {
    const size_t band_bytes = row_bytes * (size_t)f.band_rows;
    const int n_bands = (f.height + f.band_rows - 1) / f.band_rows;
    const int owned = n_bands > g.rank ? (n_bands - g.rank + g.world - 1) / g.world : 0;
    if (owned > 0) {
      const int last_band = g.rank + (owned - 1) * g.world;
      const bool last_short = (size_t)(last_band + 1) * (size_t)f.band_rows > (size_t)f.height;
      const int full = last_short ? owned - 1 : owned;
      const size_t first = (size_t)g.rank * band_bytes, pitch = (size_t)g.world * band_bytes;
      if (full > 0) {
        CK(ctx, cudaMemcpy2DAsync(dst + first, pitch, src + first, pitch, band_bytes, (size_t)full, cudaMemcpyDeviceToHost, d.copy_stream));
        copied += band_bytes * (size_t)full;
      }
      if (last_short) {
        const size_t off = (size_t)last_band * band_bytes, rest = need - off;
        CK(ctx, cudaMemcpyAsync(dst + off, src + off, rest, cudaMemcpyDeviceToHost, d.copy_stream));
        copied += rest;
      }
    }
  }
  launch_group_post((uint32_t*)(g.hdev + offsetof(HostRingHeader, done)) + (size_t)g.rank * 8 + j, (uint32_t)(k + 1), d.copy_stream);
  CK(ctx, cudaGetLastError());
  CK(ctx, cudaEventRecord(g.ticket[slot], d.copy_stream));
  d.copy_pending = true;
  ctx->stats.d2h_bytes = (int64_t)copied;
  ctx->stats.kernel_launches += 1;
  *ticket = (int32_t)(k & 0x7fffffff);
  g.slot_open[j] = true;
  g.seq++;
  return RTB_OK;
}

int group_host_end(rtb_context* ctx, int32_t ticket) {
  rtb_context::Group& g = ctx->group;
  if (g.seq - (uint64_t)ticket <= (uint64_t)rtb_context::Group::kTickets) CK(ctx, cudaEventSynchronize(g.ticket[(uint64_t)ticket % rtb_context::Group::kTickets]));
  CK(ctx, cudaGetLastError());
  if (g.rank != 0) return RTB_OK;
  const HostRingHeader* hdr = (const HostRingHeader*)g.hbase;
  const int j = (int)((uint64_t)ticket % (uint64_t)g.n_buf);
  for (int r = 1; r < g.world; r++)
    if (!host_wait_reached(&hdr->done[(size_t)r * 8 + j], (uint32_t)ticket + 1u, g.timeout_ns))
      return fail(ctx, RTB_E_CUDA, "a rank of the group did not arrive in time (RTB_GROUP_TIMEOUT_MS)");
  if (g.seq - (uint64_t)ticket <= (uint64_t)g.n_buf) g.slot_open[j] = false;
  return RTB_OK;
}
}  // namespace

// =====================================================================================================================
// C ABI
// =====================================================================================================================
extern "C" {

int rtb_api_version(void) { return RTB_API_VERSION; }

void rtb_abi_sizes(int32_t* out, int32_t n) {
  const int32_t v[] = {(int32_t)sizeof(rtb_xform_elem), (int32_t)sizeof(rtb_material), (int32_t)sizeof(rtb_triangle), (int32_t)sizeof(rtb_mesh),
                       (int32_t)sizeof(rtb_prim), (int32_t)sizeof(rtb_scene_desc), (int32_t)sizeof(rtb_render_params), (int32_t)sizeof(rtb_stats),
                       (int32_t)(4 * lbvh_node_f4)};  // [8]: 32-bit words per LBVH node record (rtb_get_bvh)
  for (int32_t i = 0; i < n && i < (int32_t)(sizeof v / sizeof v[0]); i++) out[i] = v[i];
}

void rtb_params_default(rtb_render_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof *p);
  p->light_intensity = 1.0f;
  p->max_depth = 2;
  p->enable_ambient = p->enable_diffuse = p->enable_specular = p->enable_refraction = 1;
  p->aa_samples = 1;
}

int rtb_create(rtb_context** out, const int32_t* device_ids, int32_t n_devices) {
  if (!out) return fail(nullptr, RTB_E_ARG, "out is null");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    return fail(nullptr, RTB_E_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                         " (librtb200 has no CPU fallback)");
  }
  std::vector<int32_t> ids;
  if (!device_ids || n_devices <= 0) ids.push_back(0);
  else ids.assign(device_ids, device_ids + n_devices);
  for (int32_t id : ids)
    if (id < 0 || id >= count) return fail(nullptr, RTB_E_ARG, "device id out of range");
  std::unique_ptr<rtb_context> ctx(new rtb_context());
  if (const char* env = std::getenv("RTB_CHUNK_SLOTS")) {
    const long long v = std::atoll(env);
    if (v >= 1024) ctx->chunk_slots = v;
  }
  if (const char* env = std::getenv("RTB_LANES")) ctx->n_lanes = std::min((int)DeviceState::kMaxLanes, std::max(1, std::atoi(env)));
  if (const char* env = std::getenv("RTB_SMEM")) ctx->smem_mode = std::atoi(env);
  if (const char* env = std::getenv("RTB_SPLIT_BLOCKING")) ctx->split_blocking = std::atoi(env);
  if (const char* env = std::getenv("RTB_WIDE")) ctx->wide = std::atoi(env);
  if (const char* env = std::getenv("RTB_POOL")) ctx->pool = std::atoi(env);
  if (const char* env = std::getenv("RTB_PACKET_CLOSEST")) ctx->packet_closest = std::atoi(env);
  if (const char* env = std::getenv("RTB_PACKET_SHADOW")) ctx->packet_shadow = std::atoi(env);
  if (const char* env = std::getenv("RTB_TAIL_MAX")) ctx->tail_max = (int32_t)std::max(0LL, std::atoll(env));
  ctx->devs.resize(ids.size());
  for (size_t k = 0; k < ids.size(); k++) {
    DeviceState& d = ctx->devs[k];
    d.device = ids[k];
    CK(nullptr, cudaSetDevice(d.device));
    cudaDeviceProp prop;
    CK(nullptr, cudaGetDeviceProperties(&prop, d.device));
    if (prop.major < 10) return fail(nullptr, RTB_E_CUDA, "librtb200 is built for sm_100a (B200) only");
    d.sm_count = prop.multiProcessorCount;
    d.smem_limit = prop.sharedMemPerBlockOptin;
    for (auto& l : d.lane) {
      CK(nullptr, cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
      CK(nullptr, cudaEventCreate(&l.ev_begin));
      CK(nullptr, cudaEventCreate(&l.ev_end));
      CK(nullptr, cudaEventCreateWithFlags(&l.ev_done, cudaEventDisableTiming));
    }
    d.stream = d.lane[0].stream;
    CK(nullptr, cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    CK(nullptr, cudaEventCreateWithFlags(&d.ev_copy, cudaEventDisableTiming));
    CK(nullptr, cudaEventCreateWithFlags(&d.ev_resolve, cudaEventDisableTiming));
    CK(nullptr, cudaEventCreate(&d.ev_begin));
    CK(nullptr, cudaEventCreate(&d.ev_end));
    CK(nullptr, cudaEventCreateWithFlags(&d.ev_done, cudaEventDisableTiming));
    CK(nullptr, cudaMalloc(&d.sphere_table, kSphereVerts * 3 * sizeof(float)));
    CK(nullptr, cudaMemcpy(d.sphere_table, unit_sphere_table(), kSphereVerts * 3 * sizeof(float), cudaMemcpyHostToDevice));
  }
  // peers store their bands straight into device 0's frame
  for (size_t k = 1; k < ids.size(); k++) {
    int can = 0;
    CK(nullptr, cudaDeviceCanAccessPeer(&can, ids[k], ids[0]));
    if (!can) return fail(nullptr, RTB_E_CUDA, "peer access to device 0 is not available");
    CK(nullptr, cudaSetDevice(ids[k]));
    e = cudaDeviceEnablePeerAccess(ids[0], 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(nullptr, RTB_E_CUDA, cudaGetErrorString(e));
    cudaGetLastError();
  }
  cudaSetDevice(ids[0]);
  *out = ctx.release();
  return RTB_OK;
}

void rtb_destroy(rtb_context* ctx) {
  if (!ctx) return;
  for (auto& e : ctx->ticket_event)
    if (e) { cudaSetDevice(ctx->devs[0].device); cudaEventDestroy(e); }
  for (auto& e : ctx->frame_ready)
    if (e) { cudaSetDevice(ctx->devs[0].device); cudaEventDestroy(e); }
  group_release(ctx);
  for (auto& d : ctx->devs) device_sync(d);
  for (void* p : ctx->ipc_opened) {
    cudaSetDevice(ctx->devs[0].device);
    cudaIpcCloseMemHandle(p);
  }
  for (auto& x : ctx->externals) {
    cudaSetDevice(ctx->devs[0].device);
    cudaFree(x.ptr);
    cudaDestroyExternalMemory(x.mem);
  }
  for (auto& d : ctx->devs) {
    device_sync(d);
    release_scene_pool(d);
    free_targets(d);
    dfree(d.sphere_table);
    for (auto* pool : {&d.prof_trace, &d.prof_shadow, &d.prof_resolve})
      for (auto& pr : *pool) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    if (d.ev_begin) cudaEventDestroy(d.ev_begin);
    if (d.ev_end) cudaEventDestroy(d.ev_end);
    if (d.ev_done) cudaEventDestroy(d.ev_done);
    if (d.ev_resolve) cudaEventDestroy(d.ev_resolve);
    if (d.ev_copy) cudaEventDestroy(d.ev_copy);
    if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
    for (auto& l : d.lane) {
      if (l.ev_begin) cudaEventDestroy(l.ev_begin);
      if (l.ev_end) cudaEventDestroy(l.ev_end);
      if (l.ev_done) cudaEventDestroy(l.ev_done);
      if (l.stream) cudaStreamDestroy(l.stream);
    }
  }
  delete ctx;
}

const char* rtb_last_error(rtb_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rtb_upload_scene(rtb_context* ctx, const rtb_scene_desc* scene, int32_t primitive_mode, int32_t bvh_mode) {
  if (!ctx || !scene) return fail(ctx, RTB_E_ARG, "null argument");
  if (primitive_mode != RTB_PRIM_TESSELLATED && primitive_mode != RTB_PRIM_ANALYTIC) return fail(ctx, RTB_E_ARG, "unknown primitive_mode");
  if (bvh_mode != RTB_BVH_REFERENCE && bvh_mode != RTB_BVH_LBVH) return fail(ctx, RTB_E_ARG, "unknown bvh_mode");
  const std::string why = ctx->host.assign(*scene, /*copy_triangles=*/false);
  if (!why.empty()) return fail(ctx, RTB_E_ARG, why);
  ctx->has_scene = false;
  std::vector<FlattenObject> objs;
  std::vector<float> prims;
  const int64_t n_out = build_object_table(*scene, primitive_mode == RTB_PRIM_ANALYTIC, objs, prims);
  if (n_out < 0) return fail(ctx, RTB_E_ARG, "scene has more than 2^31 triangles");
  if (bvh_mode == RTB_BVH_LBVH && n_out >= (1 << 28)) return fail(ctx, RTB_E_ARG, "LBVH mode supports fewer than 2^28 triangles");
  std::vector<float> mats;
  pack_materials(*scene, mats);
  float ms_build = 0.0f;
  struct timespec a, b;
  clock_gettime(CLOCK_MONOTONIC, &a);
  for (auto& d : ctx->devs) {
    float ms = 0.0f;
    const int rc = upload_on_device(ctx, d, *scene, objs, (int32_t)n_out, mats, prims, bvh_mode, ms);
    if (rc != RTB_OK) return rc;
    ms_build = std::max(ms_build, ms);
  }
  clock_gettime(CLOCK_MONOTONIC, &b);
  cudaSetDevice(ctx->devs[0].device);
  ctx->primitive_mode = primitive_mode;
  ctx->bvh_mode = bvh_mode;
  ctx->has_scene = true;
  ctx->stats.n_triangles = n_out;
  ctx->stats.n_nodes = ctx->devs[0].scene.n_nodes;
  ctx->stats.ms_build = ms_build;
  ctx->stats.ms_upload = (float)((b.tv_sec - a.tv_sec) * 1e3 + (b.tv_nsec - a.tv_nsec) * 1e-6);
  return RTB_OK;
}

int rtb_invalidate(rtb_context* ctx) {
  if (!ctx) return RTB_E_ARG;
  for (auto& d : ctx->devs) { device_sync(d); free_scene(d); }  // the arrays stay pooled for the next upload (rtb_clear_target returns them)
  cudaSetDevice(ctx->devs[0].device);
  ctx->has_scene = false;
  return RTB_OK;
}

int rtb_clear_target(rtb_context* ctx) {
  if (!ctx) return RTB_E_ARG;
  if (!ctx->ipc_opened.empty() || ctx->group.active) return fail(ctx, RTB_E_ARG, "frame buffers are shared over IPC; destroy the context instead");
  for (auto& d : ctx->devs)
    if (d.frame_exported) return fail(ctx, RTB_E_ARG, "the frame buffer is exported over IPC (peers may have it mapped); destroy the context instead");
  for (auto& d : ctx->devs) {
    device_sync(d);
    free_targets(d);
    if (!ctx->has_scene) release_scene_pool(d);  // ReleaseBuffers (RayTracer.cs:47-59) = rtb_invalidate + rtb_clear_target: everything goes
  }
  cudaSetDevice(ctx->devs[0].device);
  return RTB_OK;
}

int rtb_render(rtb_context* ctx, const rtb_render_params* p, uint8_t* rgba8, size_t bytes, int32_t* out_w, int32_t* out_h) {
  if (!ctx || !p || !rgba8) return fail(ctx, RTB_E_ARG, "null argument");
  if (p->band_world > 1) return fail(ctx, RTB_E_ARG, "rtb_render renders whole frames; use rtb_render_device for band sharding");
  if (ctx->has_scene) {  // check the caller's capacity before any work
    FrameParams f;
    std::string why;
    if (!resolve_frame(ctx->host.d, *p, f, why)) return fail(ctx, RTB_E_ARG, why);
    if (out_w) *out_w = f.width;
    if (out_h) *out_h = f.height;
    if (bytes < (size_t)f.width * f.height * 4) return fail(ctx, RTB_E_SIZE, "rgba8 buffer too small for the resolved resolution");
  }
  FrameParams f;
  // A blocking frame has nothing else in flight to hide the end of its launches behind, so a large one is split into one row
  // range per lane: the ranges render concurrently and each is copied back as soon as its resolve kernel has finished (C4:
  // 5.2 -> 4.9 ms).  Small frames lose more to the four-fold launch count than they gain (1080p sample scene: 0.89 -> 1.24 ms),
  // hence the 4 M pixel-sample threshold.
  const bool split = ctx->split_blocking != 0 && ctx->devs.size() == 1 && ctx->n_lanes > 1 && !ctx->profiling && ctx->has_scene;
  if (split) {
    FrameParams g;
    std::string why;
    if (!resolve_frame(ctx->host.d, *p, g, why)) return fail(ctx, RTB_E_ARG, why);
    const int64_t slots = (int64_t)((g.width + 7) / 8) * ((g.height + 3) / 4) * 32 * g.spp;
    const int parts = std::min(ctx->n_lanes, 4);  // more than four row ranges cost more in launches than their overlap returns
    if (slots >= (1 << 22) && g.debug == 0) ctx->chunk_slots_once = (slots + parts - 1) / parts;
  }
  const int rc = render_frame(ctx, p, nullptr, 0, /*to_internal_frame=*/true, /*sync=*/false, f);
  const bool was_split = ctx->chunk_slots_once > 0;
  ctx->chunk_slots_once = 0;
  if (rc != RTB_OK) return rc;
  DeviceState& d0 = ctx->devs[0];
  const size_t need = (size_t)f.width * f.height * 4;
  CK(ctx, cudaSetDevice(d0.device));
  if (was_split && !d0.last_chunks.empty()) {  // ReadPixels (RayTracer.cs:371-375), one row range at a time, on the copy stream
    const size_t row_bytes = (size_t)f.width * 4;
    for (const auto& c : d0.last_chunks) {
      CK(ctx, cudaStreamWaitEvent(d0.copy_stream, d0.lane[c.lane].ev_done, 0));
      CK(ctx, cudaMemcpyAsync(rgba8 + (size_t)c.row0 * row_bytes, (const uint8_t*)d0.frame + (size_t)c.row0 * row_bytes, (size_t)c.rows * row_bytes,
                              cudaMemcpyDeviceToHost, d0.copy_stream));
    }
    d0.copy_pending = true;
  } else {
    CK(ctx, join_lanes(d0));
    CK(ctx, cudaMemcpyAsync(rgba8, d0.frame, need, cudaMemcpyDeviceToHost, d0.stream));  // ReadPixels, RayTracer.cs:371-375
  }
  for (auto& d : ctx->devs) device_sync(d);
  CK(ctx, cudaSetDevice(d0.device));
  CK(ctx, cudaGetLastError());
  ctx->stats.d2h_bytes = (int64_t)need;
  return RTB_OK;
}

// Pipelined form of rtb_render: enqueues the frame and its readback and returns; up to kTickets frames may be in flight.
// Successive frames rotate over the lanes (their tails overlap the next frame's bulk) and over as many device frame buffers.
// `indexed`: the frame is mapped to GIF palette indices on the device (gif.cu: k_palette) and 1 byte per pixel is read back.
namespace {
int begin_frame(rtb_context* ctx, const rtb_render_params* p, uint8_t* host_dst, size_t bytes, int32_t* ticket, bool indexed) {
  if (!ctx || !p || !host_dst || !ticket) return fail(ctx, RTB_E_ARG, "null argument");
  if (p->band_world > 1) return fail(ctx, RTB_E_ARG, "rtb_render_begin renders whole frames");
  if (ctx->group.active && ctx->group.host) return fail(ctx, RTB_E_ARG, "a host-ring group owns the pipelined frame buffers: use rtb_group_render_begin, or destroy the group");
  if (!ctx->has_scene) return fail(ctx, RTB_E_NOSCENE, "no scene uploaded (rtb_upload_scene)");
  FrameParams f;
  std::string why;
  if (!resolve_frame(ctx->host.d, *p, f, why)) return fail(ctx, RTB_E_ARG, why);
  const size_t need = (size_t)f.width * f.height * 4;
  const size_t need_out = indexed ? (size_t)f.width * f.height : need;
  if (bytes < need_out) return fail(ctx, RTB_E_SIZE, indexed ? "index buffer too small for the resolved resolution" : "rgba8 buffer too small for the resolved resolution");
  DeviceState& d = ctx->devs[0];
  CK(ctx, cudaSetDevice(d.device));
  const uint64_t n = ctx->tickets_issued;
  const int n_buf = std::max(2, ctx->n_lanes);
  const int slot = (int)(n % rtb_context::kTickets), buf = (int)(n % (uint64_t)n_buf);
  if (!ctx->frame_ready[slot]) CK(ctx, cudaEventCreateWithFlags(&ctx->frame_ready[slot], cudaEventDisableTiming));
  if (!ctx->ticket_event[slot]) CK(ctx, cudaEventCreateWithFlags(&ctx->ticket_event[slot], cudaEventDisableTiming));
  else CK(ctx, cudaEventSynchronize(ctx->ticket_event[slot]));  // the ring is full: wait for the frame issued kTickets calls ago
  if (d.frame_async_bytes[buf] < need) {
    device_sync(d);
    dfree(d.frame_async[buf]);
    d.frame_async_bytes[buf] = 0;
    CK(ctx, cudaMalloc(&d.frame_async[buf], need));
    d.frame_async_bytes[buf] = need;
  }
  if (indexed && d.index_async_bytes[buf] < need_out) {
    device_sync(d);
    dfree(d.index_async[buf]);
    d.index_async_bytes[buf] = 0;
    CK(ctx, cudaMalloc(&d.index_async[buf], need_out));
    d.index_async_bytes[buf] = need_out;
  }
  // this frame reuses the device buffer of the frame issued n_buf calls ago: its readback must have finished (on every device
  // of a multi-device context: the peers store their bands into the same buffer)
  if (n >= (uint64_t)n_buf) {
    cudaEvent_t prev = ctx->ticket_event[(n - (uint64_t)n_buf) % rtb_context::kTickets];
    for (auto& dev : ctx->devs) {
      CK(ctx, cudaSetDevice(dev.device));
      for (auto& l : dev.lane) CK(ctx, cudaStreamWaitEvent(l.stream, prev, 0));
    }
    CK(ctx, cudaSetDevice(d.device));
  }
  const int rc = render_frame(ctx, p, d.frame_async[buf], d.frame_async_bytes[buf], /*to_internal_frame=*/false, /*sync=*/false, f);
  if (rc != RTB_OK) return rc;
  LaneState& last = d.lane[d.last_lane];
  for (int k = 0; k < DeviceState::kMaxLanes; k++)  // multi-chunk frame: every lane that carried one of its chunks
    if (k != d.last_lane && d.lane[k].stream && d.lane[k].frame_id == ctx->frame_id) CK(ctx, cudaStreamWaitEvent(last.stream, d.lane[k].ev_done, 0));
  if (ctx->devs.size() > 1) {  // render_frame made device 0's primary stream wait for the peers' stores: order the readback behind it
    CK(ctx, cudaEventRecord(d.ev_done, d.stream));
    if (last.stream != d.stream) CK(ctx, cudaStreamWaitEvent(last.stream, d.ev_done, 0));
  }
  if (indexed) {  // ConvertToIndexed, GifGenerator.cs:346-369
    launch_palette(d.frame_async[buf], f.width, f.height, (uint8_t*)d.index_async[buf], last.stream);
    ctx->stats.kernel_launches++;
  }
  // the readback runs on its own stream, so this lane can start the frame after next while the copy is on the wire
  CK(ctx, cudaEventRecord(ctx->frame_ready[slot], last.stream));
  CK(ctx, cudaStreamWaitEvent(d.copy_stream, ctx->frame_ready[slot], 0));
  CK(ctx, cudaMemcpyAsync(host_dst, indexed ? d.index_async[buf] : d.frame_async[buf], need_out, cudaMemcpyDeviceToHost, d.copy_stream));
  CK(ctx, cudaEventRecord(ctx->ticket_event[slot], d.copy_stream));
  d.copy_pending = true;
  CK(ctx, cudaEventRecord(last.ev_done, last.stream));
  last.used = true;
  ctx->stats.d2h_bytes = (int64_t)need_out;
  *ticket = (int32_t)(n & 0x7fffffff);
  ctx->tickets_issued++;
  return RTB_OK;
}
}  // namespace

int rtb_render_begin(rtb_context* ctx, const rtb_render_params* p, uint8_t* rgba8, size_t bytes, int32_t* ticket) {
  return begin_frame(ctx, p, rgba8, bytes, ticket, /*indexed=*/false);
}

int rtb_render_begin_indexed(rtb_context* ctx, const rtb_render_params* p, uint8_t* indexed, size_t bytes, int32_t* ticket) {
  return begin_frame(ctx, p, indexed, bytes, ticket, /*indexed=*/true);
}

// Waits until the frame of `ticket` is complete in its host buffer.
int rtb_render_end(rtb_context* ctx, int32_t ticket) {
  if (!ctx) return RTB_E_ARG;
  const uint64_t issued = ctx->tickets_issued;
  if (ticket < 0 || (uint64_t)ticket >= issued) return fail(ctx, RTB_E_ARG, "unknown ticket");
  const uint64_t t = (uint64_t)ticket;
  if (issued - t > (uint64_t)rtb_context::kTickets) return RTB_OK;  // its ring slot was recycled, which waited for it
  CK(ctx, cudaSetDevice(ctx->devs[0].device));
  CK(ctx, cudaEventSynchronize(ctx->ticket_event[t % rtb_context::kTickets]));
  CK(ctx, cudaGetLastError());
  return RTB_OK;
}

int rtb_render_device(rtb_context* ctx, const rtb_render_params* p, void* dst_device, size_t bytes, int32_t sync) {
  if (!ctx || !p || !dst_device) return fail(ctx, RTB_E_ARG, "null argument");
  FrameParams f;
  return render_frame(ctx, p, dst_device, bytes, /*to_internal_frame=*/false, sync != 0, f);
}

int rtb_render_aux(rtb_context* ctx, const rtb_render_params* p, int32_t* prim_id, float* t, int32_t* material) {
  if (!ctx || !p) return fail(ctx, RTB_E_ARG, "null argument");
  if (!ctx->has_scene) return fail(ctx, RTB_E_NOSCENE, "no scene uploaded (rtb_upload_scene)");
  FrameParams f;
  std::string why;
  if (!resolve_frame(ctx->host.d, *p, f, why)) return fail(ctx, RTB_E_ARG, why);
  DeviceState& d = ctx->devs[0];
  CK(ctx, cudaSetDevice(d.device));
  const size_t n_px = (size_t)f.width * f.height;
  if (d.aux_px < n_px) {
    dfree(d.aux_prim); dfree(d.aux_mat); dfree(d.aux_t);
    d.aux_px = 0;
    CK(ctx, cudaMalloc(&d.aux_prim, n_px * 4));
    CK(ctx, cudaMalloc(&d.aux_mat, n_px * 4));
    CK(ctx, cudaMalloc(&d.aux_t, n_px * 4));
    d.aux_px = n_px;
  }
  device_sync(d);
  launch_aux(d.scene.flavour, f, scene_view(d.scene), d.aux_prim, d.aux_t, d.aux_mat, d.sm_count * 8, d.stream);
  CK(ctx, cudaGetLastError());
  if (prim_id) CK(ctx, cudaMemcpyAsync(prim_id, d.aux_prim, n_px * 4, cudaMemcpyDeviceToHost, d.stream));
  if (t) CK(ctx, cudaMemcpyAsync(t, d.aux_t, n_px * 4, cudaMemcpyDeviceToHost, d.stream));
  if (material) CK(ctx, cudaMemcpyAsync(material, d.aux_mat, n_px * 4, cudaMemcpyDeviceToHost, d.stream));
  CK(ctx, cudaStreamSynchronize(d.stream));
  return RTB_OK;
}

int rtb_get_triangles(rtb_context* ctx, float* v_n_18, int32_t* material, int64_t capacity, int64_t* n_out) {
  if (!ctx) return RTB_E_ARG;
  if (!ctx->has_scene) return fail(ctx, RTB_E_NOSCENE, "no scene uploaded (rtb_upload_scene)");
  DeviceState& d = ctx->devs[0];
  const int64_t n = d.scene.n_tris;
  if (n_out) *n_out = n;
  if (!v_n_18 && !material) return RTB_OK;
  if (capacity < n) return fail(ctx, RTB_E_SIZE, "triangle capacity too small");
  if (n == 0) return RTB_OK;
  CK(ctx, cudaSetDevice(d.device));
  std::vector<float> raw((size_t)n * 12), nrm((size_t)n * 12);
  CK(ctx, cudaMemcpy(raw.data(), d.scene.raw, raw.size() * 4, cudaMemcpyDeviceToHost));
  CK(ctx, cudaMemcpy(nrm.data(), d.scene.nrm, nrm.size() * 4, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; i++) {
    if (v_n_18)
      for (int k = 0; k < 3; k++)
        for (int c = 0; c < 3; c++) {
          v_n_18[i * 18 + k * 3 + c] = raw[(size_t)i * 12 + (size_t)k * 4 + (size_t)c];
          v_n_18[i * 18 + 9 + k * 3 + c] = nrm[(size_t)i * 12 + (size_t)k * 4 + (size_t)c];
        }
    if (material) std::memcpy(&material[i], &nrm[(size_t)i * 12 + 3], 4);
  }
  return RTB_OK;
}

// Parity/debug access to the acceleration structure: nodes as stored on the device (reference mode: 8 floats per node =
// GPUBVHNode; LBVH: 16 floats per node) and the leaf-order -> emission-order permutation.
int rtb_get_bvh(rtb_context* ctx, void* nodes, int64_t nodes_capacity_bytes, int64_t* n_nodes, int32_t* perm, int64_t perm_capacity) {
  if (!ctx) return RTB_E_ARG;
  if (!ctx->has_scene) return fail(ctx, RTB_E_NOSCENE, "no scene uploaded (rtb_upload_scene)");
  DeviceState& d = ctx->devs[0];
  CK(ctx, cudaSetDevice(d.device));
  if (n_nodes) *n_nodes = d.scene.n_tris > 0 ? d.scene.n_nodes : 0;
  const size_t node_bytes = d.scene.n_tris > 0 ? (size_t)d.scene.n_nodes * (size_t)d.scene.node_floats * 4 : 0;
  if (nodes && node_bytes) {
    if ((size_t)nodes_capacity_bytes < node_bytes) return fail(ctx, RTB_E_SIZE, "node capacity too small");
    CK(ctx, cudaMemcpy(nodes, d.scene.nodes, node_bytes, cudaMemcpyDeviceToHost));
  }
  if (perm && d.scene.n_tris) {
    if (perm_capacity < d.scene.n_tris) return fail(ctx, RTB_E_SIZE, "perm capacity too small");
    CK(ctx, cudaMemcpy(perm, d.scene.perm, (size_t)d.scene.n_tris * 4, cudaMemcpyDeviceToHost));
  }
  return RTB_OK;
}

int rtb_get_bvh_node_words(rtb_context* ctx) {
  if (!ctx || !ctx->has_scene) return 0;
  return ctx->devs[0].scene.node_floats;
}

int rtb_get_stats(rtb_context* ctx, rtb_stats* out) {
  if (!ctx || !out) return RTB_E_ARG;
  const int rc = collect_stats(ctx);
  if (rc != RTB_OK) return rc;
  *out = ctx->stats;
  return RTB_OK;
}

int rtb_set_profiling(rtb_context* ctx, int32_t enable) {
  if (!ctx) return RTB_E_ARG;
  ctx->profiling = enable != 0;
  return RTB_OK;
}

int rtb_set_cancel_flag(rtb_context* ctx, const volatile int32_t* flag) {
  if (!ctx) return RTB_E_ARG;
  ctx->cancel = flag;
  return RTB_OK;
}

int rtb_synchronize(rtb_context* ctx) {
  if (!ctx) return RTB_E_ARG;
  for (auto& d : ctx->devs) device_sync(d);
  cudaSetDevice(ctx->devs[0].device);
  CK(ctx, cudaGetLastError());
  return RTB_OK;
}

int rtb_flush(rtb_context* ctx) {
  if (!ctx) return RTB_E_ARG;
  for (auto& d : ctx->devs) { CK(ctx, cudaSetDevice(d.device)); CK(ctx, join_lanes(d)); }
  cudaSetDevice(ctx->devs[0].device);
  return RTB_OK;
}

void* rtb_get_stream(rtb_context* ctx, int32_t index) {
  if (!ctx || index < 0 || index >= (int32_t)ctx->devs.size()) return nullptr;
  return (void*)ctx->devs[(size_t)index].stream;
}

void* rtb_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void rtb_free_pinned(void* p) { if (p) cudaFreeHost(p); }

int rtb_frame_export(rtb_context* ctx, size_t bytes, void** dev_ptr, uint8_t handle64[64]) {
  if (!ctx || (!handle64 && !dev_ptr)) return fail(ctx, RTB_E_ARG, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DeviceState& d = ctx->devs[0];
  CK(ctx, cudaSetDevice(d.device));
  if (d.frame_exported && d.frame_bytes < bytes) return fail(ctx, RTB_E_ARG, "a smaller frame buffer was already exported; its importers would be left with a dangling mapping");
  const int rc = ensure_frame(ctx, d, bytes);
  if (rc != RTB_OK) return rc;
  if (handle64) {  // without a handle the call only names the context's own frame buffer (nothing leaves the process)
    d.frame_exported = true;
    cudaIpcMemHandle_t h;
    CK(ctx, cudaIpcGetMemHandle(&h, d.frame));
    std::memcpy(handle64, &h, 64);
  }
  if (dev_ptr) *dev_ptr = d.frame;
  return RTB_OK;
}

int rtb_frame_import(rtb_context* ctx, const uint8_t handle64[64], void** dev_ptr) {
  if (!ctx || !handle64 || !dev_ptr) return fail(ctx, RTB_E_ARG, "null argument");
  CK(ctx, cudaSetDevice(ctx->devs[0].device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* p = nullptr;
  CK(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  ctx->ipc_opened.push_back(p);
  *dev_ptr = p;
  return RTB_OK;
}

// ---- process-per-GPU frame ring ------------------------------------------------------------------------------------------------
// One frame split over the ranks of a job by row bands, gathered on rank 0 by the resolve kernels' peer stores, read
// back by rank 0, pipelined over n_buffers frames — all inside the library: every rank only enqueues (group.cu has the
// hand-shake), so frame k+1 .. k+n_buffers-1 render on all ranks while rank 0 copies frame k out.

int rtb_group_create(rtb_context* ctx, int32_t rank, int32_t world, size_t frame_bytes, int32_t n_buffers, uint8_t handle64[64]) {
  if (!ctx || !handle64) return fail(ctx, RTB_E_ARG, "null argument");
  if (world < 1 || world > 32 || rank < 0 || rank >= world || n_buffers < 1 || n_buffers > 8 || frame_bytes == 0) return fail(ctx, RTB_E_ARG, "bad group shape");
  if (ctx->devs.size() != 1) return fail(ctx, RTB_E_ARG, "a group member is a single-device context");
  group_release(ctx);
  rtb_context::Group& g = ctx->group;
  CK(ctx, cudaSetDevice(ctx->devs[0].device));
  g.active = true;  // from here on a failing call leaves a group that rtb_group_destroy / rtb_destroy cleans up
  g.rank = rank; g.world = world; g.n_buf = n_buffers; g.frame_bytes = frame_bytes;
  g.stride = (frame_bytes + 255) & ~(size_t)255;
  const size_t flag_bytes = 4096;
  const size_t total = g.stride * (size_t)n_buffers + flag_bytes;
  if (rank == 0) {
    CK(ctx, cudaMalloc(&g.base, total));
    g.owner = true;
    CK(ctx, cudaMemset(g.base + g.stride * (size_t)n_buffers, 0, flag_bytes));
    cudaIpcMemHandle_t h;
    CK(ctx, cudaIpcGetMemHandle(&h, g.base));
    std::memcpy(handle64, &h, 64);
  } else {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void* p = nullptr;
    CK(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g.base = (uint8_t*)p;
  }
  g.flags = (uint32_t*)(g.base + g.stride * (size_t)n_buffers);
  CK(ctx, cudaMalloc(&g.error, sizeof(uint32_t)));
  CK(ctx, cudaMemset(g.error, 0, sizeof(uint32_t)));
  CK(ctx, cudaStreamCreateWithFlags(&g.gate, cudaStreamNonBlocking));
  CK(ctx, cudaEventCreateWithFlags(&g.gate_done, cudaEventDisableTiming));
  if (const char* env = std::getenv("RTB_GROUP_TIMEOUT_MS")) g.timeout_ns = (unsigned long long)std::max(1LL, std::atoll(env)) * 1000000ull;
  g.seq = 0;
  return RTB_OK;
}

// Host ring: the frames of the job land in POSIX shared memory `shm_name` ("/name"; rank 0 creates it, the others attach once
// rank 0's call has returned), page-locked by every rank.  The readback then uses every GPU's own PCIe link instead of rank 0's
// alone, and no device memory is shared between the processes at all.
int rtb_group_create_host(rtb_context* ctx, int32_t rank, int32_t world, size_t frame_bytes, int32_t n_buffers, const char* shm_name) {
  if (!ctx || !shm_name) return fail(ctx, RTB_E_ARG, "null argument");
  if (world < 1 || world > 32 || rank < 0 || rank >= world || n_buffers < 1 || n_buffers > 8 || frame_bytes == 0) return fail(ctx, RTB_E_ARG, "bad group shape");
  if (shm_name[0] != '/' || std::strlen(shm_name) < 2 || std::strlen(shm_name) > 200) return fail(ctx, RTB_E_ARG, "shm_name must look like \"/name\"");
  if (ctx->devs.size() != 1) return fail(ctx, RTB_E_ARG, "a group member is a single-device context");
  group_release(ctx);
  rtb_context::Group& g = ctx->group;
  CK(ctx, cudaSetDevice(ctx->devs[0].device));
  g.active = true;
  g.host = true;
  g.rank = rank; g.world = world; g.n_buf = n_buffers; g.frame_bytes = frame_bytes;
  g.stride = (frame_bytes + 4095) & ~(size_t)4095;
  g.hbytes = kHostRingHeaderBytes + g.stride * (size_t)n_buffers;
  int fd = -1;
  if (rank == 0) {
    shm_unlink(shm_name);  // a stale object of a crashed job
    fd = shm_open(shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return fail(ctx, RTB_E_IO, std::string("shm_open(create) failed for ") + shm_name);
    g.owner = true;
    g.shm_name = shm_name;
    // posix_fallocate, not just ftruncate: a /dev/shm that is too small must fail HERE, not as SIGBUS when a page is first touched
    if (ftruncate(fd, (off_t)g.hbytes) != 0 || posix_fallocate(fd, 0, (off_t)g.hbytes) != 0) {
      close(fd);
      return fail(ctx, RTB_E_IO, "the host ring does not fit into /dev/shm");
    }
  } else {
    fd = shm_open(shm_name, O_RDWR, 0600);
    if (fd < 0) return fail(ctx, RTB_E_IO, std::string("shm_open failed for ") + shm_name + " (rank 0 creates it first)");
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (size_t)sb.st_size != g.hbytes) { close(fd); return fail(ctx, RTB_E_ARG, "the host ring has another shape than this call describes"); }
  }
  void* m = mmap(nullptr, g.hbytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return fail(ctx, RTB_E_IO, "mmap of the host ring failed");
  g.hbase = (uint8_t*)m;
  HostRingHeader* hdr = (HostRingHeader*)g.hbase;
  if (rank == 0) {
    std::memset(g.hbase, 0, kHostRingHeaderBytes);
    hdr->world = (uint32_t)world; hdr->n_buf = (uint32_t)n_buffers; hdr->frame_bytes = frame_bytes; hdr->stride = g.stride;
    __atomic_store_n(&hdr->magic, kHostRingMagic, __ATOMIC_RELEASE);
  } else if (__atomic_load_n(&hdr->magic, __ATOMIC_ACQUIRE) != kHostRingMagic || hdr->world != (uint32_t)world || hdr->n_buf != (uint32_t)n_buffers ||
             hdr->frame_bytes != frame_bytes) {
    return fail(ctx, RTB_E_ARG, "the host ring was created for another group shape");
  }
  CK(ctx, cudaHostRegister(g.hbase, g.hbytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  g.registered = true;
  void* dp = nullptr;
  CK(ctx, cudaHostGetDevicePointer(&dp, g.hbase, 0));
  g.hdev = (uint8_t*)dp;
  if (const char* env = std::getenv("RTB_GROUP_TIMEOUT_MS")) g.timeout_ns = (unsigned long long)std::max(1LL, std::atoll(env)) * 1000000ull;
  g.seq = 0;
  return RTB_OK;
}

// Rank 0 of a host-ring group: where the frame of `ticket` lies (after rtb_group_render_end(ticket)); it stays there until rank 0
// begins frame ticket + n_buffers.
int rtb_group_frame(rtb_context* ctx, int32_t ticket, const uint8_t** rgba8) {
  if (!ctx || !rgba8) return fail(ctx, RTB_E_ARG, "null argument");
  *rgba8 = nullptr;
  rtb_context::Group& g = ctx->group;
  if (!g.active || !g.host || !g.hbase) return fail(ctx, RTB_E_ARG, "no host-ring group (rtb_group_create_host)");
  if (g.rank != 0) return fail(ctx, RTB_E_ARG, "frames are complete on rank 0 only");
  if (ticket < 0 || (uint64_t)ticket >= g.seq || g.seq - (uint64_t)ticket > (uint64_t)g.n_buf) return fail(ctx, RTB_E_ARG, "that frame has left the ring");
  *rgba8 = g.hbase + kHostRingHeaderBytes + ((uint64_t)ticket % (uint64_t)g.n_buf) * g.stride;
  return RTB_OK;
}

int rtb_group_destroy(rtb_context* ctx) {
  if (!ctx) return RTB_E_ARG;
  group_release(ctx);
  return RTB_OK;
}

int rtb_group_render_begin(rtb_context* ctx, const rtb_render_params* p, uint8_t* rgba8, size_t bytes, int32_t* ticket) {
  if (!ctx || !p || !ticket) return fail(ctx, RTB_E_ARG, "null argument");
  rtb_context::Group& g = ctx->group;
  if (g.active && g.host) {
    if (!ctx->has_scene) return fail(ctx, RTB_E_NOSCENE, "no scene uploaded (rtb_upload_scene)");
    return group_host_begin(ctx, p, rgba8, ticket);
  }
  if (!g.active || !g.base || !g.error || !g.gate || !g.gate_done) return fail(ctx, RTB_E_ARG, "no group (rtb_group_create)");
  if (!ctx->has_scene) return fail(ctx, RTB_E_NOSCENE, "no scene uploaded (rtb_upload_scene)");
  rtb_render_params pp = *p;
  pp.band_rank = g.rank; pp.band_world = g.world; pp.out_layout = RTB_OUT_FRAME;
  if (pp.band_rows <= 0) pp.band_rows = 8;
  FrameParams f;
  std::string why;
  if (!resolve_frame(ctx->host.d, pp, f, why)) return fail(ctx, RTB_E_ARG, why);
  const size_t need = (size_t)f.width * f.height * 4;
  if (need > g.frame_bytes) return fail(ctx, RTB_E_SIZE, "frame larger than the group's buffers");
  if (g.rank == 0 && (!rgba8 || bytes < need)) return fail(ctx, RTB_E_SIZE, "rgba8 buffer too small for the resolved resolution");
  DeviceState& d = ctx->devs[0];
  CK(ctx, cudaSetDevice(d.device));
  const uint64_t k = g.seq;
  const int j = (int)(k % (uint64_t)g.n_buf), slot = (int)(k % rtb_context::Group::kTickets);
  uint32_t* stored = g.flags;                                   // [world][n_buf]
  uint32_t* read_done = g.flags + (size_t)g.world * g.n_buf;    // [n_buf]
  if (!g.ticket[slot]) CK(ctx, cudaEventCreateWithFlags(&g.ticket[slot], cudaEventDisableTiming));
  else CK(ctx, cudaEventSynchronize(g.ticket[slot]));  // the ring of tickets is full: wait for the frame begun kTickets calls ago
  // the buffer still holds frame k - n_buf until rank 0 has copied it out
  if (k >= (uint64_t)g.n_buf) {
    launch_group_wait(read_done + j, 1, 1, (uint32_t)(k - (uint64_t)g.n_buf + 1), g.error, g.timeout_ns, g.gate);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaEventRecord(g.gate_done, g.gate));
    for (int l = 0; l < ctx->n_lanes; l++) CK(ctx, cudaStreamWaitEvent(d.lane[l].stream, g.gate_done, 0));
  }
  uint8_t* dst = g.base + (size_t)j * g.stride;
  FrameParams f2;
  const int rc = render_frame(ctx, &pp, dst, g.stride, /*to_internal_frame=*/false, /*sync=*/false, f2);
  if (rc != RTB_OK) return rc;
  LaneState& last = d.lane[d.last_lane];
  for (int l = 0; l < DeviceState::kMaxLanes; l++)  // multi-chunk frame: every lane that carried one of its chunks
    if (l != d.last_lane && d.lane[l].stream && d.lane[l].frame_id == ctx->frame_id) CK(ctx, cudaStreamWaitEvent(last.stream, d.lane[l].ev_done, 0));
  launch_group_post(stored + (size_t)g.rank * g.n_buf + j, (uint32_t)(k + 1), last.stream);
  CK(ctx, cudaGetLastError());
  CK(ctx, cudaEventRecord(last.ev_done, last.stream));
  last.used = true;
  if (g.rank == 0) {
    launch_group_wait(stored + j, g.world, g.n_buf, (uint32_t)(k + 1), g.error, g.timeout_ns, d.copy_stream);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaMemcpyAsync(rgba8, dst, need, cudaMemcpyDeviceToHost, d.copy_stream));  // ReadPixels, RayTracer.cs:371-375
    launch_group_post(read_done + j, (uint32_t)(k + 1), d.copy_stream);
    CK(ctx, cudaGetLastError());
    CK(ctx, cudaEventRecord(g.ticket[slot], d.copy_stream));
    d.copy_pending = true;
    ctx->stats.d2h_bytes = (int64_t)need;
  } else {
    CK(ctx, cudaEventRecord(g.ticket[slot], last.stream));
  }
  ctx->stats.kernel_launches += g.rank == 0 ? 3 : 1;
  *ticket = (int32_t)(k & 0x7fffffff);
  g.seq++;
  return RTB_OK;
}

int rtb_group_render_end(rtb_context* ctx, int32_t ticket) {
  if (!ctx) return RTB_E_ARG;
  rtb_context::Group& g = ctx->group;
  if (!g.active) return fail(ctx, RTB_E_ARG, "no group (rtb_group_create)");
  if (ticket < 0 || (uint64_t)ticket >= g.seq) return fail(ctx, RTB_E_ARG, "unknown ticket");
  CK(ctx, cudaSetDevice(ctx->devs[0].device));
  if (g.host) return group_host_end(ctx, ticket);
  if (g.seq - (uint64_t)ticket <= (uint64_t)rtb_context::Group::kTickets) CK(ctx, cudaEventSynchronize(g.ticket[(uint64_t)ticket % rtb_context::Group::kTickets]));
  uint32_t err = 0;
  CK(ctx, cudaMemcpy(&err, g.error, sizeof err, cudaMemcpyDeviceToHost));
  if (err) return fail(ctx, RTB_E_CUDA, "a rank of the group did not arrive in time (RTB_GROUP_TIMEOUT_MS)");
  CK(ctx, cudaGetLastError());
  return RTB_OK;
}

// Zero-copy realtime path (RayTracer.RenderToTexture, RayTracer.cs:82-202, hands Unity a GPU texture it never reads back,
// SceneBuilder.cs:836-852): map a graphics-API allocation and let rtb_render_device's resolve kernel store into it.
int rtb_external_import(rtb_context* ctx, int32_t handle_type, void* handle, size_t bytes, int32_t dedicated, void** dev_ptr) {
  if (!ctx || !dev_ptr || bytes == 0) return fail(ctx, RTB_E_ARG, "null argument");
  *dev_ptr = nullptr;
  cudaExternalMemoryHandleDesc hd;
  std::memset(&hd, 0, sizeof hd);
  switch (handle_type) {
    case RTB_EXT_OPAQUE_FD: hd.type = cudaExternalMemoryHandleTypeOpaqueFd; hd.handle.fd = (int)(intptr_t)handle; break;
    case RTB_EXT_OPAQUE_WIN32: hd.type = cudaExternalMemoryHandleTypeOpaqueWin32; hd.handle.win32.handle = handle; break;
    case RTB_EXT_D3D12_HEAP: hd.type = cudaExternalMemoryHandleTypeD3D12Heap; hd.handle.win32.handle = handle; break;
    case RTB_EXT_D3D12_RESOURCE: hd.type = cudaExternalMemoryHandleTypeD3D12Resource; hd.handle.win32.handle = handle; break;
    default: return fail(ctx, RTB_E_ARG, "unknown external handle type");
  }
  hd.size = bytes;
  hd.flags = dedicated ? cudaExternalMemoryDedicated : 0;
  CK(ctx, cudaSetDevice(ctx->devs[0].device));
  cudaExternalMemory_t mem = nullptr;
  CK(ctx, cudaImportExternalMemory(&mem, &hd));
  cudaExternalMemoryBufferDesc bd;
  std::memset(&bd, 0, sizeof bd);
  bd.offset = 0;
  bd.size = bytes;
  void* p = nullptr;
  cudaError_t e = cudaExternalMemoryGetMappedBuffer(&p, mem, &bd);
  if (e != cudaSuccess) {
    cudaDestroyExternalMemory(mem);
    return fail(ctx, RTB_E_CUDA, std::string("cudaExternalMemoryGetMappedBuffer failed: ") + cudaGetErrorString(e));
  }
  ctx->externals.push_back({mem, p, bytes});
  *dev_ptr = p;
  return RTB_OK;
}

int rtb_external_release(rtb_context* ctx, void* dev_ptr) {
  if (!ctx || !dev_ptr) return fail(ctx, RTB_E_ARG, "null argument");
  for (size_t i = 0; i < ctx->externals.size(); i++)
    if (ctx->externals[i].ptr == dev_ptr) {
      for (auto& d : ctx->devs) device_sync(d);  // no kernel may still be storing into the mapping
      CK(ctx, cudaSetDevice(ctx->devs[0].device));
      cudaFree(ctx->externals[i].ptr);
      cudaDestroyExternalMemory(ctx->externals[i].mem);
      ctx->externals.erase(ctx->externals.begin() + (long)i);
      return RTB_OK;
    }
  return fail(ctx, RTB_E_ARG, "not a pointer returned by rtb_external_import");
}

// Copies device 0's internal frame (the buffer rtb_frame_export shares) to the host: the readback step after a
// multi-process peer-store gather.
int rtb_frame_read(rtb_context* ctx, uint8_t* rgba8, size_t bytes) {
  if (!ctx || !rgba8) return fail(ctx, RTB_E_ARG, "null argument");
  DeviceState& d = ctx->devs[0];
  if (!d.frame || d.frame_bytes < bytes) return fail(ctx, RTB_E_SIZE, "internal frame smaller than requested");
  CK(ctx, cudaSetDevice(d.device));
  CK(ctx, join_lanes(d));
  CK(ctx, cudaMemcpyAsync(rgba8, d.frame, bytes, cudaMemcpyDeviceToHost, d.stream));
  device_sync(d);
  ctx->stats.d2h_bytes = (int64_t)bytes;
  return RTB_OK;
}

// ---- host-only helpers (no device needed) -----------------------------------------------------------------------------

// The uniform block RayTracer.cs:221-355 would upload for (scene, settings): out25 = cameraToObject row-major 4x4,
// camDist, tanHalf, orthoSize, light xyz, bg xyz; wh = resolved width, height.
int rtb_resolve_frame(const rtb_scene_desc* scene, const rtb_render_params* p, float* out25, int32_t* wh) {
  if (!scene || !p) return RTB_E_ARG;
  FrameParams f;
  std::string why;
  if (!resolve_frame(*scene, *p, f, why)) { g_create_error = why; return RTB_E_ARG; }
  if (out25) {
    std::memcpy(out25, f.cam, 48);
    out25[12] = 0.0f; out25[13] = 0.0f; out25[14] = 0.0f; out25[15] = 1.0f;
    out25[16] = f.cam_dist; out25[17] = f.tan_half; out25[18] = f.ortho_size;
    std::memcpy(out25 + 19, f.light, 12);
    std::memcpy(out25 + 22, f.bg, 12);
  }
  if (wh) { wh[0] = f.width; wh[1] = f.height; }
  return RTB_OK;
}

// Host restatement of the reference BVH build on caller-provided triangles (12 floats each: v0,cx,v1,cy,v2,cz): lets
// the CPU test-suite compare the product's builder with the checker's without a GPU.
int rtb_build_reference_bvh(const float* raw12, int32_t n, float* nodes8, int64_t nodes_capacity, int64_t* n_nodes, int32_t* perm) {
  if (!raw12 && n > 0) return RTB_E_ARG;
  RefBvh b;
  build_reference_bvh(raw12, n, b);
  const int64_t nn = (int64_t)(b.nodes.size() / 8);
  if (n_nodes) *n_nodes = n > 0 ? nn : 0;
  if (n <= 0) return RTB_OK;
  if (nodes8) {
    if (nodes_capacity < nn) return RTB_E_SIZE;
    std::memcpy(nodes8, b.nodes.data(), b.nodes.size() * 4);
  }
  if (perm) std::memcpy(perm, b.perm.data(), (size_t)n * 4);
  return RTB_OK;
}

int rtb_scene_parse(const char* text, size_t len, rtb_scene** out, char* err, size_t err_cap) {
  if (!text || !out) return RTB_E_ARG;
  std::unique_ptr<rtb_scene> s(new rtb_scene());
  try {
    parse_scene_text(text, len, s->h);
  } catch (const std::exception& e) {
    if (err && err_cap) std::snprintf(err, err_cap, "%s", e.what());
    return RTB_E_PARSE;
  }
  *out = s.release();
  return RTB_OK;
}

int rtb_scene_load(const char* path, rtb_scene** out, char* err, size_t err_cap) {
  if (!path || !out) return RTB_E_ARG;
  std::ifstream in(path, std::ios::binary);
  if (!in) {  // the reference logs and returns an empty ObjectData (SceneService.cs:28-33); here it is an error code
    if (err && err_cap) std::snprintf(err, err_cap, "cannot open %s", path);
    return RTB_E_IO;
  }
  std::stringstream buf;
  buf << in.rdbuf();
  const std::string text = buf.str();
  return rtb_scene_parse(text.data(), text.size(), out, err, err_cap);
}

const rtb_scene_desc* rtb_scene_get(const rtb_scene* s) { return s ? &s->h.d : nullptr; }
void rtb_scene_free(rtb_scene* s) { delete s; }

// ---- GIF sweep (GifGenerator.cs; host coder and container in gif.cu) -------------------------------------------------------

int rtb_gif_index_frame(rtb_context* ctx, const uint8_t* rgba8, int32_t width, int32_t height, uint8_t* indexed) {
  if (!ctx || !rgba8 || !indexed || width <= 0 || height <= 0) return fail(ctx, RTB_E_ARG, "bad argument");
  DeviceState& d = ctx->devs[0];
  CK(ctx, cudaSetDevice(d.device));
  const size_t n_px = (size_t)width * height;
  if (d.gif_scratch_bytes < n_px * 5) {  // RGBA8 in + indices out, kept for the following frames of the same GIF
    CK(ctx, cudaStreamSynchronize(d.stream));
    dfree(d.gif_scratch);
    d.gif_scratch_bytes = 0;
    CK(ctx, cudaMalloc(&d.gif_scratch, n_px * 5));
    d.gif_scratch_bytes = n_px * 5;
  }
  uint8_t* src = (uint8_t*)d.gif_scratch;
  uint8_t* dst = src + n_px * 4;
  CK(ctx, cudaMemcpyAsync(src, rgba8, n_px * 4, cudaMemcpyHostToDevice, d.stream));
  launch_palette(src, width, height, dst, d.stream);
  CK(ctx, cudaGetLastError());
  CK(ctx, cudaMemcpyAsync(indexed, dst, n_px, cudaMemcpyDeviceToHost, d.stream));
  CK(ctx, cudaStreamSynchronize(d.stream));
  return RTB_OK;
}

int rtb_gif_index_device(rtb_context* ctx, const void* rgba8_device, int32_t width, int32_t height, void* indexed_device) {
  if (!ctx || !rgba8_device || !indexed_device || width <= 0 || height <= 0) return fail(ctx, RTB_E_ARG, "bad argument");
  DeviceState& d = ctx->devs[0];
  CK(ctx, cudaSetDevice(d.device));
  launch_palette(rgba8_device, width, height, (uint8_t*)indexed_device, d.stream);
  CK(ctx, cudaGetLastError());
  return RTB_OK;
}

int rtb_gif_save(rtb_context* ctx, const char* path, int32_t width, int32_t height, const uint8_t* const* rgba8_frames, int32_t n_frames,
                 int32_t frame_delay_cs, int32_t threads) {
  if (!ctx) return RTB_E_ARG;  // the palette mapping is a device kernel: there is no host fallback
  if (!path || !rgba8_frames || n_frames <= 0 || width <= 0 || height <= 0) return fail(ctx, RTB_E_ARG, "bad argument");
  const size_t n_px = (size_t)width * height;
  std::vector<std::vector<uint8_t>> indexed((size_t)n_frames, std::vector<uint8_t>(n_px));
  std::vector<const uint8_t*> ptrs((size_t)n_frames);
  for (int k = 0; k < n_frames; k++) {
    if (!rgba8_frames[k]) return fail(ctx, RTB_E_ARG, "null frame");
    const int rc_k = rtb_gif_index_frame(ctx, rgba8_frames[k], width, height, indexed[(size_t)k].data());
    if (rc_k != RTB_OK) return rc_k;
    ptrs[(size_t)k] = indexed[(size_t)k].data();
  }
  const int rc = rtb_gif_save_indexed(path, width, height, ptrs.data(), n_frames, frame_delay_cs, threads);
  return rc != RTB_OK ? fail(ctx, rc, std::string("cannot write ") + path) : rc;
}

int rtb_gif_render_rotation(rtb_context* ctx, const rtb_render_params* base, int32_t n_frames, float step_deg, const char* path,
                            int32_t frame_delay_cs, int32_t threads) {
  if (!ctx || !base || !path || n_frames <= 0) return fail(ctx, RTB_E_ARG, "bad argument");
  if (!ctx->has_scene) return fail(ctx, RTB_E_NOSCENE, "no scene uploaded (rtb_upload_scene)");
  FrameParams f;
  std::string why;
  if (!resolve_frame(ctx->host.d, *base, f, why)) return fail(ctx, RTB_E_ARG, why);
  const int width = f.width, height = f.height;
  const size_t n_px = (size_t)width * height;
  uint8_t* host = (uint8_t*)rtb_alloc_pinned(n_px * (size_t)n_frames);  // every frame's indices: 1 byte per pixel
  if (!host) return fail(ctx, RTB_E_CUDA, "cannot allocate pinned host memory for the frames");
  // frame k becomes ready when its rtb_render_end returned; workers compress ready frames in order of arrival
  std::vector<std::vector<uint8_t>> compressed((size_t)n_frames);
  std::mutex mu;
  std::condition_variable cv;
  int n_ready = 0;          // guarded by mu
  bool abort_flag = false;  // guarded by mu
  std::atomic<int> next{0};
  auto work = [&] {
    for (int k = next.fetch_add(1); k < n_frames; k = next.fetch_add(1)) {
      {
        std::unique_lock<std::mutex> lock(mu);
        cv.wait(lock, [&] { return n_ready > k || abort_flag; });
        if (abort_flag) return;
      }
      compressed[(size_t)k].resize(gif_lzw_bound(n_px));
      compressed[(size_t)k].resize(gif_lzw(host + (size_t)k * n_px, n_px, compressed[(size_t)k].data()));
    }
  };
  auto publish = [&](int ready, bool failed) {
    { std::lock_guard<std::mutex> lock(mu); n_ready = ready; abort_flag = abort_flag || failed; }
    cv.notify_all();
  };
  const int t = gif_threads(threads, n_frames);
  std::vector<std::thread> pool;
  for (int i = 0; i < t; i++) pool.emplace_back(work);
  const int in_flight = std::max(2, ctx->n_lanes);
  std::vector<int32_t> tickets((size_t)n_frames, -1);
  int rc = RTB_OK, ended = 0;
  for (int k = 0; k < n_frames && rc == RTB_OK; k++) {
    rtb_render_params p = *base;  // GifGenerator.cs:59-61: (base.x, base.y, angle)
    const float bx = base->has_cam_rot ? base->cam_rot_euler_deg[0] : 0.0f, by = base->has_cam_rot ? base->cam_rot_euler_deg[1] : 0.0f;
    p.has_cam_rot = 1;
    p.cam_rot_euler_deg[0] = bx; p.cam_rot_euler_deg[1] = by; p.cam_rot_euler_deg[2] = (float)k * step_deg;
    if (k - ended >= in_flight) {
      rc = rtb_render_end(ctx, tickets[(size_t)ended]);
      if (rc == RTB_OK) publish(++ended, false);
    }
    if (rc == RTB_OK) rc = rtb_render_begin_indexed(ctx, &p, host + (size_t)k * n_px, n_px, &tickets[(size_t)k]);
  }
  while (rc == RTB_OK && ended < n_frames) {
    rc = rtb_render_end(ctx, tickets[(size_t)ended]);
    if (rc == RTB_OK) publish(++ended, false);
  }
  if (rc != RTB_OK) publish(ended, true);
  for (auto& th : pool) th.join();
  if (rc != RTB_OK) { rtb_synchronize(ctx); rtb_free_pinned(host); return rc; }
  std::vector<uint8_t> file;
  gif_append_prologue(file, width, height);
  for (int k = 0; k < n_frames; k++) gif_append_frame(file, width, height, compressed[(size_t)k].data(), compressed[(size_t)k].size(), frame_delay_cs);
  file.push_back(0x3B);
  rtb_free_pinned(host);
  FILE* out = std::fopen(path, "wb");
  if (!out) return fail(ctx, RTB_E_IO, std::string("cannot open ") + path);
  const size_t wrote = std::fwrite(file.data(), 1, file.size(), out);
  if (std::fclose(out) != 0 || wrote != file.size()) return fail(ctx, RTB_E_IO, std::string("cannot write ") + path);
  return RTB_OK;
}

}  // extern "C"
