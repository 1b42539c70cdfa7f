// lbvh.cu — K2: Morton-code LBVH built on the GPU (replaces the recursive CPU median-split build of
// Assets/Services/BVH/BVHBuilder.cs:76-238 for triangle-heavy scenes; the exact reference-shape build stays available
// on the host for parity mode, scene_host.cpp).
//
// Pipeline (all on one stream, no host round trip):
//   k_centroid_bounds   min/max of the triangle centroids (warp shuffles + one atomic per block and component)
//   k_morton            63-bit Morton key (21 bits per axis, one scale: the centroids' bounding cube), value = emission index
//   k_sort_*            keys + values: stable LSD radix sort, 8 bits per pass (hand-written: histogram per tile, one scan of the
//                       256 x tiles counts, stable scatter ranked with __match_any_sync); k_sort_check refuses an unsorted result
//   k_hierarchy         Karras 2012 binary radix tree over the sorted keys (ties broken by index), parent links, ranges
//   k_refit             bottom-up AABBs: each leaf thread climbs, the second arrival at a node merges its children
//   k_emit / k_emit4    traversal records: every tree node whose range holds <= RTB_LEAF_MAX triangles collapses into a leaf
//                       (its triangles are contiguous in sorted order).  RTB_LBVH_WIDTH 2: every larger node becomes a
//                       64-byte record carrying BOTH children's boxes (default); RTB_LBVH_WIDTH 4: every larger node of
//                       even depth becomes a 128-byte record carrying its (up to) four grandchildren's boxes, compacted
// The sorted value array is the leaf order (perm) used by launch_pack.
#include "kernels.hpp"

namespace rtb {
namespace {

constexpr int kBlock = 256;

// Order-preserving float <-> uint map for atomicMin/Max.
__device__ __forceinline__ unsigned f2o(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float o2f(unsigned o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

struct Workspace {
  unsigned* bounds;          // 6: min xyz, max xyz (ordered-uint encoded)
  unsigned long long* keys;  // n
  unsigned long long* keys_sorted;
  int32_t* vals;             // n
  int2* child;               // n-1: Karras children (>= 0 internal, < 0: ~leaf index)
  int2* range;               // n-1: [first, last] in sorted order
  int32_t* parent;           // 2n-1: parents of internal nodes [0,n-1) then of leaves [n-1, 2n-1)
  int32_t* flag;             // n-1
  float4* box;               // 2*(2n-1): min,max of internal nodes then leaves
  int32_t* idx4;             // n-1: compact index of each emitted 4-wide node (exclusive scan of the emit flags)
  void* wide_queue[2];       // n-1 WideItem each: work lists of the wide collapse (one tree level per launch, ping-pong)
  int32_t* wide_counters;    // 4
  unsigned long long* keys_tmp;  // n: ping-pong partner of keys_sorted
  int32_t* vals_tmp;             // n: ping-pong partner of the caller's perm
  int32_t* sort_hist;            // 256 * sort_tiles(n): digit counts per tile, digit-major; scanned in place
  int32_t* scan_sums[3];         // block totals of the scan, one array per level
  int32_t* sort_error;           // 1 int: set when the sorted keys are not ascending
};

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// ---- sizes of the sort / scan scratch ----
constexpr int kSortThreads = 256, kSortRounds = 16, kSortTile = kSortThreads * kSortRounds;  // keys per block and pass
constexpr int kScanBlock = 1024;                                                              // elements per block of the scan
inline int32_t sort_tiles(int32_t n) { return (n + kSortTile - 1) / kSortTile; }
inline size_t scan_level_count(size_t n) { return (n + kScanBlock - 1) / kScanBlock; }

Workspace carve(void* base, int32_t n) {
  char* p = (char*)base;
  Workspace w;
  auto take = [&](size_t bytes) { void* r = p; p += align_up(bytes); return r; };
  const size_t ni = (size_t)(n > 1 ? n - 1 : 1), nt = (size_t)n, na = 2 * nt;
  w.bounds = (unsigned*)take(6 * sizeof(unsigned));
  w.keys = (unsigned long long*)take(nt * 8);
  w.keys_sorted = (unsigned long long*)take(nt * 8);
  w.vals = (int32_t*)take(nt * 4);
  w.child = (int2*)take(ni * 8);
  w.range = (int2*)take(ni * 8);
  w.parent = (int32_t*)take(na * 4);
  w.flag = (int32_t*)take(ni * 4);
  w.box = (float4*)take(na * 2 * sizeof(float4));
  w.idx4 = (int32_t*)take(ni * 4);
  w.wide_queue[0] = take(ni * 8);
  w.wide_queue[1] = take(ni * 8);
  w.wide_counters = (int32_t*)take(16);
  w.keys_tmp = (unsigned long long*)take(nt * 8);
  w.vals_tmp = (int32_t*)take(nt * 4);
  const size_t hist = 256 * (size_t)sort_tiles(n);
  w.sort_hist = (int32_t*)take(hist * 4);
  size_t level = scan_level_count(hist > ni ? hist : ni);
  for (int k = 0; k < 3; k++) { w.scan_sums[k] = (int32_t*)take(level * 4); level = scan_level_count(level); }
  w.sort_error = (int32_t*)take(4);
  return w;
}


// ---- exclusive prefix sum of int32 (hand-written; replaces a library scan) ------------------------------------------------------
// Three phases per level: every block scans kScanBlock elements in shared memory and writes its total; the totals are scanned the
// same way (recursively: three levels cover 2^30 elements); every block adds its offset.  In place is allowed (out == in).
__global__ void __launch_bounds__(kScanBlock) k_scan_block(const int32_t* __restrict__ in, int32_t* __restrict__ out, int32_t n, int32_t* __restrict__ sums) {
  __shared__ int32_t warp_total[kScanBlock / 32];
  const int32_t i = blockIdx.x * kScanBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int32_t v = i < n ? in[i] : 0;
  int32_t x = v;  // inclusive scan inside the warp
  for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) warp_total[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int32_t t = warp_total[lane];
    for (int o = 1; o < 32; o <<= 1) { const int32_t y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
    warp_total[lane] = t;  // inclusive over the warps
  }
  __syncthreads();
  const int32_t before = warp == 0 ? 0 : warp_total[warp - 1];
  if (i < n) out[i] = before + x - v;
  if (threadIdx.x == kScanBlock - 1 && sums) sums[blockIdx.x] = before + x;
}
__global__ void __launch_bounds__(kScanBlock) k_scan_add(int32_t* __restrict__ out, int32_t n, const int32_t* __restrict__ sums) {
  const int32_t i = blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n && blockIdx.x > 0) out[i] += sums[blockIdx.x];
}
void exclusive_scan(const int32_t* in, int32_t* out, int32_t n, int32_t* const (&sums)[3], int level, cudaStream_t st) {
  if (n <= 0) return;
  const int32_t blocks = (int32_t)scan_level_count((size_t)n);
  k_scan_block<<<blocks, kScanBlock, 0, st>>>(in, out, n, blocks > 1 ? sums[level] : nullptr);
  if (blocks > 1) {
    exclusive_scan(sums[level], sums[level], blocks, sums, level + 1, st);  // level <= 2: 1024^3 elements
    k_scan_add<<<blocks, kScanBlock, 0, st>>>(out, n, sums[level]);
  }
}

// ---- stable LSD radix sort of (63-bit key, int32 value) pairs, 8 bits per pass (hand-written; replaces a library sort) -----------
// A block owns a tile of kSortTile consecutive keys.  k_sort_hist counts the tile's digits; one exclusive scan over the digit-major
// table hist[digit][tile] turns the counts into each (digit, tile)'s first output position; k_sort_scatter walks the tile again in
// kSortRounds rounds of 256 keys and places every key at base[digit] + (keys of that digit in earlier rounds) + (in earlier warps of
// this round) + (in earlier lanes of this warp) — earlier in every sense that the input order defines, hence stable, which is what
// makes the value (the triangle's emission index) the tie-break of equal Morton keys.
__global__ void __launch_bounds__(kSortThreads) k_sort_hist(const unsigned long long* __restrict__ keys, int32_t n, int shift, int32_t tiles, int32_t* __restrict__ hist) {
  __shared__ int32_t count[256];
  count[threadIdx.x] = 0;
  __syncthreads();
  const int32_t base = blockIdx.x * kSortTile;
  for (int r = 0; r < kSortRounds; r++) {
    const int32_t i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&count[(unsigned)(keys[i] >> shift) & 255u], 1);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * tiles + blockIdx.x] = count[threadIdx.x];
}
__global__ void __launch_bounds__(kSortThreads) k_sort_scatter(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ vals, int32_t n, int shift,
                                                               int32_t tiles, const int32_t* __restrict__ first, unsigned long long* __restrict__ keys_out,
                                                               int32_t* __restrict__ vals_out) {
  __shared__ int32_t running[256];                      // position of the next key of each digit
  __shared__ int32_t warp_count[kSortThreads / 32][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  running[threadIdx.x] = first[(size_t)threadIdx.x * tiles + blockIdx.x];
  const int32_t base = blockIdx.x * kSortTile;
  for (int r = 0; r < kSortRounds; r++) {
    for (int w = 0; w < kSortThreads / 32; w++) warp_count[w][threadIdx.x] = 0;
    __syncthreads();
    const int32_t i = base + r * kSortThreads + threadIdx.x;
    const bool valid = i < n;
    const unsigned long long key = valid ? keys[i] : 0ull;
    const unsigned digit = valid ? (unsigned)(key >> shift) & 255u : 256u + (unsigned)lane;  // padding lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0) warp_count[warp][digit] = __popc(peers);
    __syncthreads();
    if (valid) {
      int32_t at = running[digit] + rank_in_warp;
      for (int w = 0; w < warp; w++) at += warp_count[w][digit];
      keys_out[at] = key;
      vals_out[at] = vals[i];
    }
    __syncthreads();
    int32_t total = 0;
    for (int w = 0; w < kSortThreads / 32; w++) total += warp_count[w][threadIdx.x];
    running[threadIdx.x] += total;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kBlock) k_sort_check(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ vals, int32_t n, int32_t* error) {
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += gridDim.x * blockDim.x) {
    const unsigned long long a = keys[i], b = keys[i + 1];
    if (a > b || (a == b && vals[i] >= vals[i + 1])) atomicExch(error, 1);  // ascending keys, ties in input order
  }
}

__global__ void k_init_bounds(unsigned* bounds) {
  if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(kBlock) k_centroid_bounds(const float4* __restrict__ raw, int32_t n, unsigned* bounds) {
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    for (int a = 0; a < 3; a++) {
      const float c = raw[3 * (size_t)i + a].w;
      mn[a] = fminf(mn[a], c);
      mx[a] = fmaxf(mx[a], c);
    }
  for (int a = 0; a < 3; a++)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  __shared__ float s_mn[3][kBlock / 32], s_mx[3][kBlock / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int a = 0; a < 3; a++) { s_mn[a][warp] = mn[a]; s_mx[a][warp] = mx[a]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int a = threadIdx.x;
    float lo = s_mn[a][0], hi = s_mx[a][0];
    for (int w = 1; w < kBlock / 32; w++) { lo = fminf(lo, s_mn[a][w]); hi = fmaxf(hi, s_mx[a][w]); }
    atomicMin(&bounds[a], f2o(lo));
    atomicMax(&bounds[3 + a], f2o(hi));
  }
}

__device__ __forceinline__ unsigned long long spread21(unsigned v) {  // 21 bits -> every third bit of 63
  unsigned long long x = v & 0x1fffffull;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

__global__ void __launch_bounds__(kBlock) k_morton(const float4* __restrict__ raw, int32_t n, const unsigned* __restrict__ bounds,
                                                   unsigned long long* __restrict__ keys, int32_t* __restrict__ vals) {
  // One scale for all three axes (the bounding CUBE of the centroids): Morton cells stay cubical, so a flat scene does not
  // spend a third of its split planes on its thin axis.
  float lo[3], scale[3], ext_max = 0.0f;
  for (int a = 0; a < 3; a++) {
    lo[a] = o2f(bounds[a]);
    ext_max = fmaxf(ext_max, o2f(bounds[3 + a]) - lo[a]);
  }
  for (int a = 0; a < 3; a++) scale[a] = ext_max > 0.0f ? 2097152.0f / ext_max : 0.0f;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned q[3];
    for (int a = 0; a < 3; a++) {
      const float v = (raw[3 * (size_t)i + a].w - lo[a]) * scale[a];
      q[a] = (unsigned)fminf(fmaxf(v, 0.0f), 2097151.0f);
    }
    keys[i] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    vals[i] = i;
  }
}

// Length of the common prefix of keys i and j (ties extended by the indices); -1 when j is outside [0, n).
__device__ __forceinline__ int common_prefix(const unsigned long long* __restrict__ keys, int32_t n, int32_t i, int32_t j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a != b) return __clzll((long long)(a ^ b));
  return 64 + __clz(i ^ j);
}

__global__ void __launch_bounds__(kBlock) k_hierarchy(const unsigned long long* __restrict__ keys, int32_t n, int2* __restrict__ child,
                                                      int2* __restrict__ range, int32_t* __restrict__ parent) {
  const int32_t n_int = n - 1;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_int; i += gridDim.x * blockDim.x) {
    const int d = common_prefix(keys, n, i, i + 1) - common_prefix(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int min_prefix = common_prefix(keys, n, i, i - d);
    int32_t l_max = 2;
    while (common_prefix(keys, n, i, i + l_max * d) > min_prefix) l_max <<= 1;
    int32_t l = 0;
    for (int32_t t = l_max >> 1; t >= 1; t >>= 1)
      if (common_prefix(keys, n, i, i + (l + t) * d) > min_prefix) l += t;
    const int32_t j = i + l * d;
    const int node_prefix = common_prefix(keys, n, i, j);
    int32_t s = 0;
    for (int32_t t = (l + 1) >> 1;; t = (t + 1) >> 1) {
      if (common_prefix(keys, n, i, i + (s + t) * d) > node_prefix) s += t;
      if (t == 1) break;
    }
    const int32_t split = i + s * d + min(d, 0);
    const int32_t first = min(i, j), last = max(i, j);
    const int32_t left = (first == split) ? ~split : split;
    const int32_t right = (last == split + 1) ? ~(split + 1) : split + 1;
    child[i] = make_int2(left, right);
    range[i] = make_int2(first, last);
    parent[left >= 0 ? left : n_int + ~left] = i;
    parent[right >= 0 ? right : n_int + ~right] = i;
    if (i == 0) parent[0] = -1;
  }
}

__global__ void __launch_bounds__(kBlock) k_refit(const float4* __restrict__ raw, const int32_t* __restrict__ vals, int32_t n,
                                                  const int2* __restrict__ child, const int32_t* __restrict__ parent, int32_t* flag,
                                                  float4* box) {
  const int32_t n_int = n - 1;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int32_t src = vals[i];
    const float4 a = raw[3 * (size_t)src], b = raw[3 * (size_t)src + 1], c = raw[3 * (size_t)src + 2];
    float4 mn = make_float4(fminf(fminf(a.x, b.x), c.x), fminf(fminf(a.y, b.y), c.y), fminf(fminf(a.z, b.z), c.z), 0.0f);
    float4 mx = make_float4(fmaxf(fmaxf(a.x, b.x), c.x), fmaxf(fmaxf(a.y, b.y), c.y), fmaxf(fmaxf(a.z, b.z), c.z), 0.0f);
    box[2 * (size_t)(n_int + i)] = mn;
    box[2 * (size_t)(n_int + i) + 1] = mx;
    if (n_int == 0) continue;
    int32_t node = parent[n_int + i];
    while (node >= 0) {
      __threadfence();
      if (atomicAdd(&flag[node], 1) == 0) break;  // first arrival: the sibling subtree is not finished yet
      const int2 ch = child[node];
      const size_t li = ch.x >= 0 ? (size_t)ch.x : (size_t)(n_int + ~ch.x), ri = ch.y >= 0 ? (size_t)ch.y : (size_t)(n_int + ~ch.y);
      const float4 lmn = __ldcg(&box[2 * li]), lmx = __ldcg(&box[2 * li + 1]), rmn = __ldcg(&box[2 * ri]), rmx = __ldcg(&box[2 * ri + 1]);
      mn = make_float4(fminf(lmn.x, rmn.x), fminf(lmn.y, rmn.y), fminf(lmn.z, rmn.z), 0.0f);
      mx = make_float4(fmaxf(lmx.x, rmx.x), fmaxf(lmx.y, rmx.y), fmaxf(lmx.z, rmx.z), 0.0f);
      __stcg(&box[2 * (size_t)node], mn);
      __stcg(&box[2 * (size_t)node + 1], mx);
      node = parent[node];
    }
  }
}

// Outward padding of an emitted box face: the traversal's FMA slab test (trace.cuh: slab_entry_fma) rounds differently
// from the exact (bound - origin) * inv form by a few ulp of |origin * inv|, i.e. ~1e-7 * |origin| in space; 1e-6 of
// (|coordinate| + scene radius) covers cameras within several scene radii and grows boxes by far less than a triangle.
__device__ __forceinline__ float pad_lo(float v, float radius) { return v - 1e-6f * (fabsf(v) + radius); }
__device__ __forceinline__ float pad_hi(float v, float radius) { return v + 1e-6f * (fabsf(v) + radius); }

#if RTB_LBVH_WIDTH != 4
// Only the tree nodes whose range holds more than RTB_LEAF_MAX triangles become records (about a third of the n - 1 Karras
// nodes); k_mark_live flags them and an exclusive scan gives each its slot, so the records are DENSE: a 128-byte cache line
// holds two live records instead of 0.7 on average, which is what L1 and L2 capacity are spent on.  Karras order is kept (a
// subtree's records stay contiguous), only the holes go.
__global__ void __launch_bounds__(kBlock) k_mark_live(int32_t n, const int2* __restrict__ range, int32_t* __restrict__ flag) {
  const int32_t n_int = n - 1;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_int; i += gridDim.x * blockDim.x) {
    const int2 rg = range[i];
    flag[i] = rg.y - rg.x + 1 > RTB_LEAF_MAX ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kBlock) k_emit(int32_t n, const int2* __restrict__ child, const int2* __restrict__ range,
                                                 const float4* __restrict__ box, const unsigned* __restrict__ bounds, const int32_t* __restrict__ flag,
                                                 const int32_t* __restrict__ slot, float4* __restrict__ nodes, int32_t* __restrict__ root_out) {
  const int32_t n_int = n - 1;
  float radius = 0.0f;
  for (int a = 0; a < 6; a++) radius = fmaxf(radius, fabsf(o2f(bounds[a])));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    root_out[0] = (n <= RTB_LEAF_MAX) ? lbvh_leaf_ref(0, n) : 0;  // the root is node 0 and, being live, gets slot 0
    root_out[1] = n_int > 0 ? max(1, slot[n_int - 1] + flag[n_int - 1]) : 1;
  }
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_int; i += gridDim.x * blockDim.x) {
    if (!flag[i]) continue;  // inside a collapsed leaf: never referenced
    const int2 ch = child[i];
    int32_t ref[2];
    size_t bi[2];
    const int32_t cc[2] = {ch.x, ch.y};
    for (int k = 0; k < 2; k++) {
      if (cc[k] < 0) { ref[k] = lbvh_leaf_ref(~cc[k], 1); bi[k] = (size_t)(n_int + ~cc[k]); }
      else {
        const int2 cr = range[cc[k]];
        const int32_t cnt = cr.y - cr.x + 1;
        ref[k] = cnt <= RTB_LEAF_MAX ? lbvh_leaf_ref(cr.x, cnt) : slot[cc[k]];
        bi[k] = (size_t)cc[k];
      }
    }
    const float4 lmn = box[2 * bi[0]], lmx = box[2 * bi[0] + 1], rmn = box[2 * bi[1]], rmx = box[2 * bi[1] + 1];
    float4* rec = nodes + 4 * (size_t)slot[i];
    rec[0] = make_float4(pad_lo(lmn.x, radius), pad_lo(lmn.y, radius), pad_lo(lmn.z, radius), __int_as_float(ref[0]));
    rec[1] = make_float4(pad_hi(lmx.x, radius), pad_hi(lmx.y, radius), pad_hi(lmx.z, radius), __int_as_float(ref[1]));
    rec[2] = make_float4(pad_lo(rmn.x, radius), pad_lo(rmn.y, radius), pad_lo(rmn.z, radius), 0.0f);
    rec[3] = make_float4(pad_hi(rmx.x, radius), pad_hi(rmx.y, radius), pad_hi(rmx.z, radius), 0.0f);
  }
}

#else
// ---- 4-wide records (RTB_LBVH_WIDTH == 4): the binary tree collapsed by two levels -----------------------------------------
// A record is emitted for every internal node of even depth whose range holds > RTB_LEAF_MAX triangles; its slots are its
// grandchildren (or a child, where that child is already a leaf).  Records are compacted with an exclusive scan of the flags.
// Layout (8 x float4 = 128 B = one cache line): min.x[4] min.y[4] min.z[4] max.x[4] max.y[4] max.z[4] ref[4] (unused);
// unused slots carry ref = RTB_REF_DONE and a copy of slot 0's box.
__device__ __forceinline__ bool is_big(const int2* __restrict__ range, int32_t i) { const int2 r = range[i]; return r.y - r.x + 1 > RTB_LEAF_MAX; }

__global__ void __launch_bounds__(kBlock) k_mark4(int32_t n, const int2* __restrict__ range, const int32_t* __restrict__ parent, int32_t* __restrict__ flag4) {
  const int32_t n_int = n - 1;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_int; i += gridDim.x * blockDim.x) {
    int depth = 0;
    for (int32_t p = parent[i]; p >= 0; p = parent[p]) depth++;
    flag4[i] = (is_big(range, i) && (depth & 1) == 0) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kBlock) k_emit4(int32_t n, const int2* __restrict__ child, const int2* __restrict__ range, const float4* __restrict__ box,
                                                  const unsigned* __restrict__ bounds, const int32_t* __restrict__ flag4, const int32_t* __restrict__ idx4,
                                                  float4* __restrict__ nodes, int32_t* __restrict__ root_out) {
  const int32_t n_int = n - 1;
  float radius = 0.0f;
  for (int a = 0; a < 6; a++) radius = fmaxf(radius, fabsf(o2f(bounds[a])));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    root_out[0] = (n <= RTB_LEAF_MAX) ? lbvh_leaf_ref(0, n) : 0;
    root_out[1] = n_int > 0 ? idx4[n_int - 1] + flag4[n_int - 1] : 0;
  }
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_int; i += gridDim.x * blockDim.x) {
    if (!flag4[i]) continue;
    int ns = 0;
    int32_t sref[4];
    size_t sbox[4];
    auto add = [&](int32_t c) {  // c: Karras child reference (>= 0 internal, < 0 leaf)
      if (c < 0) { sref[ns] = lbvh_leaf_ref(~c, 1); sbox[ns] = (size_t)(n_int + ~c); }
      else if (!is_big(range, c)) { const int2 cr = range[c]; sref[ns] = lbvh_leaf_ref(cr.x, cr.y - cr.x + 1); sbox[ns] = (size_t)c; }
      else { sref[ns] = idx4[c]; sbox[ns] = (size_t)c; }  // even depth, big: a record of its own
      ns++;
    };
    const int2 ch = child[i];
    const int32_t cc[2] = {ch.x, ch.y};
    for (int k = 0; k < 2; k++) {
      if (cc[k] >= 0 && is_big(range, cc[k])) { const int2 g = child[cc[k]]; add(g.x); add(g.y); }
      else add(cc[k]);
    }
    float mn[3][4], mx[3][4];
    int32_t refs[4];
    for (int k = 0; k < 4; k++) {
      const int src = k < ns ? k : 0;
      const float4 bmn = box[2 * sbox[src]], bmx = box[2 * sbox[src] + 1];
      mn[0][k] = pad_lo(bmn.x, radius); mn[1][k] = pad_lo(bmn.y, radius); mn[2][k] = pad_lo(bmn.z, radius);
      mx[0][k] = pad_hi(bmx.x, radius); mx[1][k] = pad_hi(bmx.y, radius); mx[2][k] = pad_hi(bmx.z, radius);
      refs[k] = k < ns ? sref[k] : RTB_REF_DONE;
    }
    float4* out = nodes + 8 * (size_t)idx4[i];
    for (int a = 0; a < 3; a++) {
      out[a] = make_float4(mn[a][0], mn[a][1], mn[a][2], mn[a][3]);
      out[3 + a] = make_float4(mx[a][0], mx[a][1], mx[a][2], mx[a][3]);
    }
    out[6] = make_float4(__int_as_float(refs[0]), __int_as_float(refs[1]), __int_as_float(refs[2]), __int_as_float(refs[3]));
    out[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
}

#endif


// ---- 8-wide quantised records (RTB_BVH_WIDE; layout and arithmetic: trace.cuh) ----------------------------------------------------
// The binary radix tree is collapsed top-down, one tree level of WIDE nodes per launch: a wide node starts from the two
// children of its binary node and keeps replacing the child with the largest surface area by that child's two children until it
// has eight (children whose range holds <= RTB_LEAF_MAX triangles are leaves and stay).  The children that remain inner nodes
// get consecutive record indices (one atomicAdd) and form the next level's work list.  Children go to the slot whose octant
// best matches their position relative to the node's centre (greedy assignment), which is what lets the traversal order them
// by (slot XOR ray octant).  Triangles keep the Morton order (leaf references address `perm` positions), so the record order
// may vary from run to run but nothing a ray returns does.
struct WideItem { int32_t bnode, wnode; };

__device__ __forceinline__ float half_area(const float4 mn, const float4 mx) {
  const float dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
  return dx * dy + dy * dz + dz * dx;
}

// Quantisation grid of one axis: the biased exponent byte of S = 2^15 * cell, cell = the smallest power of two with 255 cells >= extent.
__device__ __forceinline__ unsigned grid_exponent(float extent, float& inv_cell) {
  int e = -100;
  if (extent > 0.0f) e = ilogbf(extent / 255.0f) + 1;
  e = e < -100 ? -100 : (e > 100 ? 100 : e);
  while (e < 100 && ldexpf(255.0f, e) < extent) e++;
  inv_cell = ldexpf(1.0f, -e);
  return (unsigned)(e + 15 + 127);
}

__global__ void __launch_bounds__(128) k_wide_level(int32_t n, const int2* __restrict__ child, const int2* __restrict__ range, const float4* __restrict__ box,
                                                    const WideItem* __restrict__ in, const int32_t* __restrict__ n_in, WideItem* __restrict__ out,
                                                    int32_t* n_out, int32_t* n_nodes, float4* __restrict__ nodes, int32_t capacity, int32_t* error) {
  const int32_t n_int = n - 1;
  const int32_t count_in = *n_in;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count_in; i += gridDim.x * blockDim.x) {
    const WideItem it = in[i];
    int32_t ref[8];  // Karras references: >= 0 inner node, < 0 ~leaf
    int ns = 2;
    { const int2 ch = child[it.bnode]; ref[0] = ch.x; ref[1] = ch.y; }
    auto big = [&](int32_t r) { if (r < 0) return false; const int2 rg = range[r]; return rg.y - rg.x + 1 > RTB_LEAF_MAX; };
    auto box_at = [&](int32_t r) { return r >= 0 ? (size_t)r : (size_t)(n_int + ~r); };
    while (ns < 8) {
      int best = -1;
      float best_area = -1.0f;
      for (int k = 0; k < ns; k++)
        if (big(ref[k])) {
          const size_t b = box_at(ref[k]);
          const float a = half_area(box[2 * b], box[2 * b + 1]);
          if (a > best_area) { best_area = a; best = k; }
        }
      if (best < 0) break;
      const int2 ch = child[ref[best]];
      ref[best] = ch.x;
      ref[ns++] = ch.y;
    }
    int n_inner = 0;
    for (int k = 0; k < ns; k++) n_inner += big(ref[k]) ? 1 : 0;
    int32_t wbase = 0, qbase = 0;
    if (n_inner > 0) {
      wbase = atomicAdd(n_nodes, n_inner);
      qbase = atomicAdd(n_out, n_inner);
      if (wbase + n_inner > capacity) { atomicExch(error, 1); n_inner = 0; }
    }
    // slot assignment: cost[s][k] = sigma(s) . (centre_k - centre_node), take the largest remaining pair each round
    const float4 nmn = box[2 * (size_t)it.bnode], nmx = box[2 * (size_t)it.bnode + 1];
    const float ccx = 0.5f * (nmn.x + nmx.x), ccy = 0.5f * (nmn.y + nmx.y), ccz = 0.5f * (nmn.z + nmx.z);
    float vx[8], vy[8], vz[8];
    for (int k = 0; k < ns; k++) {
      const size_t b = box_at(ref[k]);
      const float4 mn = box[2 * b], mx = box[2 * b + 1];
      vx[k] = 0.5f * (mn.x + mx.x) - ccx; vy[k] = 0.5f * (mn.y + mx.y) - ccy; vz[k] = 0.5f * (mn.z + mx.z) - ccz;
    }
    int child_of_slot[8];
    for (int sidx = 0; sidx < 8; sidx++) child_of_slot[sidx] = -1;
    unsigned child_done = 0u;
    for (int round = 0; round < ns; round++) {
      int bs = -1, bk = -1;
      float bc = -INFINITY;
      for (int sidx = 0; sidx < 8; sidx++) {
        if (child_of_slot[sidx] >= 0) continue;
        for (int k = 0; k < ns; k++) {
          if (child_done & (1u << k)) continue;
          const float c = ((sidx & 4) ? vx[k] : -vx[k]) + ((sidx & 2) ? vy[k] : -vy[k]) + ((sidx & 1) ? vz[k] : -vz[k]);
          if (c > bc) { bc = c; bs = sidx; bk = k; }
        }
      }
      if (bs < 0) {  // every remaining cost compared false (NaN coordinates): any free slot and any remaining child will do
        for (int sidx = 0; sidx < 8 && bs < 0; sidx++) if (child_of_slot[sidx] < 0) bs = sidx;
        for (int k = 0; k < ns && bk < 0; k++) if (!(child_done & (1u << k))) bk = k;
      }
      child_of_slot[bs] = bk;
      child_done |= 1u << bk;
    }
    // quantise and write the record
    float icx, icy, icz;
    const unsigned esx = grid_exponent(nmx.x - nmn.x, icx), esy = grid_exponent(nmx.y - nmn.y, icy), esz = grid_exponent(nmx.z - nmn.z, icz);
    unsigned q[6][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}};  // lo.x lo.y lo.z hi.x hi.y hi.z, 8 bytes each
    int32_t out_ref[8];
    unsigned valid = 0u;
    int inner_seen = 0;
    for (int sidx = 0; sidx < 8; sidx++) {
      const int k = child_of_slot[sidx];
      unsigned lo[3] = {255u, 255u, 255u}, hi[3] = {0u, 0u, 0u};  // empty slot: inverted box
      out_ref[sidx] = RTB_REF_DONE;
      if (k >= 0) {
        valid |= 1u << sidx;
        const size_t b = box_at(ref[k]);
        const float4 mn = box[2 * b], mx = box[2 * b + 1];
        const float l[3] = {floorf((mn.x - nmn.x) * icx), floorf((mn.y - nmn.y) * icy), floorf((mn.z - nmn.z) * icz)};
        const float h[3] = {ceilf((mx.x - nmn.x) * icx), ceilf((mx.y - nmn.y) * icy), ceilf((mx.z - nmn.z) * icz)};
        for (int a = 0; a < 3; a++) {
          lo[a] = (unsigned)fminf(fmaxf(l[a], 0.0f), 255.0f);
          hi[a] = (unsigned)fminf(fmaxf(h[a], 0.0f), 255.0f);
        }
        if (ref[k] < 0) out_ref[sidx] = lbvh_leaf_ref(~ref[k], 1);
        else if (!big(ref[k])) { const int2 rg = range[ref[k]]; out_ref[sidx] = lbvh_leaf_ref(rg.x, rg.y - rg.x + 1); }
        else if (inner_seen < n_inner) {
          out_ref[sidx] = wbase + inner_seen;
          out[qbase + inner_seen] = WideItem{ref[k], wbase + inner_seen};
          inner_seen++;
        }
      }
      for (int a = 0; a < 3; a++) {
        q[a][sidx >> 2] |= lo[a] << (8 * (sidx & 3));
        q[3 + a][sidx >> 2] |= hi[a] << (8 * (sidx & 3));
      }
    }
    float4* rec = nodes + RTB_WIDE_F4 * (size_t)it.wnode;
    rec[0] = make_float4(nmn.x, nmn.y, nmn.z, __uint_as_float(esx | (esy << 8) | (esz << 16) | (valid << 24)));
    rec[1] = make_float4(__int_as_float(out_ref[0]), __int_as_float(out_ref[1]), __int_as_float(out_ref[2]), __int_as_float(out_ref[3]));
    rec[2] = make_float4(__uint_as_float(q[0][0]), __uint_as_float(q[0][1]), __uint_as_float(q[1][0]), __uint_as_float(q[1][1]));
    rec[3] = make_float4(__uint_as_float(q[2][0]), __uint_as_float(q[2][1]), __uint_as_float(q[3][0]), __uint_as_float(q[3][1]));
    rec[4] = make_float4(__uint_as_float(q[4][0]), __uint_as_float(q[4][1]), __uint_as_float(q[5][0]), __uint_as_float(q[5][1]));
    rec[5] = make_float4(__int_as_float(out_ref[4]), __int_as_float(out_ref[5]), __int_as_float(out_ref[6]), __int_as_float(out_ref[7]));
  }
}

__global__ void k_wide_init(int32_t n, WideItem* queue, int32_t* counters, int32_t* root_out) {
  // counters: [0] items in queue A, [1] items in queue B, [2] records allocated, [3] error flag
  const bool has_root = n > RTB_LEAF_MAX;
  queue[0] = WideItem{0, 0};
  counters[0] = has_root ? 1 : 0;
  counters[1] = 0;
  counters[2] = has_root ? 1 : 0;
  counters[3] = 0;
  root_out[0] = has_root ? 0 : lbvh_leaf_ref(0, n);
  root_out[1] = 0;
}

inline int grid_for(int64_t n) {
  const int64_t g = (n + kBlock - 1) / kBlock;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

static size_t sort_scratch_bytes(int32_t n) {
  const size_t nt = (size_t)n, ni = (size_t)(n > 1 ? n - 1 : 1), hist = 256 * (size_t)sort_tiles(n);
  size_t total = align_up(nt * 8) + align_up(nt * 4) + align_up(hist * 4) + align_up(4);
  size_t level = scan_level_count(hist > ni ? hist : ni);
  for (int k = 0; k < 3; k++) { total += align_up(level * 4); level = scan_level_count(level); }
  return total;
}

size_t lbvh_workspace_bytes(int32_t n) {
  if (n <= 0) return 256;
  const size_t ni = (size_t)(n > 1 ? n - 1 : 1), nt = (size_t)n, na = 2 * nt;
  return align_up(24) + 2 * align_up(nt * 8) + align_up(nt * 4) + 2 * align_up(ni * 8) + align_up(na * 4) + align_up(ni * 4) +
         align_up(na * 2 * sizeof(float4)) + align_up(ni * 4) + 2 * align_up(ni * 8) + align_up(16) + sort_scratch_bytes(n) + 256;
}

cudaError_t lbvh_build(const float4* raw, int32_t n, const LbvhBuffers& b, cudaStream_t st, bool wide) {
  if (n <= 0) return cudaSuccess;
  Workspace w = carve(b.workspace, n);
  k_init_bounds<<<1, 32, 0, st>>>(w.bounds);
  k_centroid_bounds<<<grid_for(n), kBlock, 0, st>>>(raw, n, w.bounds);
  k_morton<<<grid_for(n), kBlock, 0, st>>>(raw, n, w.bounds, w.keys, w.vals);
  // eight stable passes of eight bits; the buffers ping-pong so that the last pass lands in (keys_sorted, perm)
  cudaError_t e = cudaMemsetAsync(w.sort_error, 0, sizeof(int32_t), st);
  if (e != cudaSuccess) return e;
  {
    const int32_t tiles = sort_tiles(n);
    const unsigned long long* kin = w.keys;
    const int32_t* vin = w.vals;
    for (int pass = 0; pass < 8; pass++) {
      unsigned long long* kout = (pass & 1) ? w.keys_sorted : w.keys_tmp;
      int32_t* vout = (pass & 1) ? b.perm : w.vals_tmp;
      k_sort_hist<<<tiles, kSortThreads, 0, st>>>(kin, n, 8 * pass, tiles, w.sort_hist);
      exclusive_scan(w.sort_hist, w.sort_hist, 256 * tiles, w.scan_sums, 0, st);
      k_sort_scatter<<<tiles, kSortThreads, 0, st>>>(kin, vin, n, 8 * pass, tiles, w.sort_hist, kout, vout);
      kin = kout;
      vin = vout;
    }
    k_sort_check<<<grid_for(n), kBlock, 0, st>>>(w.keys_sorted, b.perm, n, w.sort_error);
    int32_t bad = 0;
    e = cudaMemcpyAsync(&bad, w.sort_error, sizeof bad, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    if (bad) return cudaErrorAssert;  // never build a hierarchy over keys that are not sorted
  }
  if (n > 1) {
    e = cudaMemsetAsync(w.flag, 0, (size_t)(n - 1) * 4, st);
    if (e != cudaSuccess) return e;
    k_hierarchy<<<grid_for(n - 1), kBlock, 0, st>>>(w.keys_sorted, n, w.child, w.range, w.parent);
  }
  k_refit<<<grid_for(n), kBlock, 0, st>>>(raw, b.perm, n, w.child, w.parent, w.flag, w.box);
  if (wide) {
    // Collapse into 8-wide records, one level per launch.  The number of levels is not known on the host: launch them in batches
    // and read the work-list size back after each batch (typically one round trip: 12 levels reach 8^12 leaves).
    WideItem* qa = (WideItem*)w.wide_queue[0];
    WideItem* qb = (WideItem*)w.wide_queue[1];
    int32_t* cnt = w.wide_counters;
    k_wide_init<<<1, 1, 0, st>>>(n, qa, cnt, b.root_out);
    const int32_t capacity = n > 1 ? n - 1 : 1;
    int side = 0;
    for (int batch = 0; batch < 16; batch++) {
      for (int level = 0; level < 12; level++) {
        cudaError_t e2 = cudaMemsetAsync(cnt + (side ^ 1), 0, sizeof(int32_t), st);
        if (e2 != cudaSuccess) return e2;
        k_wide_level<<<148 * 4, 128, 0, st>>>(n, w.child, w.range, w.box, side ? qb : qa, cnt + side, side ? qa : qb, cnt + (side ^ 1), cnt + 2,
                                               b.nodes, capacity, cnt + 3);
        side ^= 1;
      }
      int32_t host[4] = {0, 0, 0, 0};
      cudaError_t e2 = cudaMemcpyAsync(host, cnt, sizeof host, cudaMemcpyDeviceToHost, st);
      if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
      if (e2 != cudaSuccess) return e2;
      if (host[3] != 0) return cudaErrorMemoryAllocation;
      if (host[side] == 0) {
        e2 = cudaMemcpyAsync(b.root_out + 1, cnt + 2, sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
        return e2 != cudaSuccess ? e2 : cudaGetLastError();
      }
    }
    return cudaErrorUnknown;  // deeper than 192 wide levels: not a tree this builder produces
  }
#if RTB_LBVH_WIDTH == 4
  if (n > 1) {
    k_mark4<<<grid_for(n - 1), kBlock, 0, st>>>(n, w.range, w.parent, w.flag);
    exclusive_scan(w.flag, w.idx4, n - 1, w.scan_sums, 0, st);
  }
  k_emit4<<<grid_for(n > 1 ? n - 1 : 1), kBlock, 0, st>>>(n, w.child, w.range, w.box, w.bounds, w.flag, w.idx4, b.nodes, b.root_out);
#else
  if (n > 1) {  // w.flag has done its job in k_refit: reuse it for the live flags
    k_mark_live<<<grid_for(n - 1), kBlock, 0, st>>>(n, w.range, w.flag);
    exclusive_scan(w.flag, w.idx4, n - 1, w.scan_sums, 0, st);
  }
  k_emit<<<grid_for(n > 1 ? n - 1 : 1), kBlock, 0, st>>>(n, w.child, w.range, w.box, w.bounds, w.flag, w.idx4, b.nodes, b.root_out);
#endif
  return cudaGetLastError();
}

}  // namespace rtb
