// trace.cu — the wavefront kernels of the per-pixel path (replaces CSMain, Assets/Shaders/BVHRayTracing.compute:273-511).
//
//   k_raygen<BVH>              primary rays :283-349, one warp per 8x4 pixel tile at full SIMD width; rays that miss the
//                              scene's root box get the background at once (:364-368), the others are compacted into
//                              the depth-0 ray queue.
//
// One depth d of the reference's `for depth` loop (:360-473) is then two launches:
//
//   k_traverse_{ref,lbvh}      every BVH query pending at this point, in ONE persistent kernel: the closest-hit rays of
//                              depth d (:362) and the shadow rays emitted at depth d-1 (:395-406, as any-hit queries).
//                              Warps claim up to 64 consecutive work items with one atomicAdd (fewer towards the end of the
//                              queue) and refill their lanes as soon as 16 of them are idle; traversal is "while-while":
//                              all lanes descend inner nodes, then all lanes intersect their leaf.  A closest-hit ray ends by writing a 16-byte hit record, a
//                              shadow ray by adding the lit or unlit increment of :418 to its slot's sampleColor.
//   k_shade                    miss/background :364-368, shading :370-418 (emits the shadow ray with both candidate
//                              increments), continuation :420-473, compacted into the next depth's ray queue: warp ballots,
//                              then one atomicAdd per BLOCK and queue (block_reserve), which keeps 256-entry runs contiguous.
//
//   k_resolve                  :475-478,510: average the spp slots of a pixel, saturate, UNORM8, store (possibly into a peer
//                              GPU's frame: the NVLink gather is this store).
//   k_debug<BVH>, k_aux<BVH>   debug views :484-508 and the primary-hit maps used by the parity tests.
//
// Queue sizes stay on the device (kernels read the counters the previous launch filled), so a frame is a fixed launch
// sequence with no host round trip.
#include "kernels.hpp"
#include "trace.cuh"

namespace rtb {

namespace {

constexpr int kBlock = 256;
// Persistent traversal kernels (global-memory variant): block size and resident blocks per SM the register allocation is
// bounded for.  Measured on the B200: round 1 (profiles/r1e_sweep_occupancy.log) found 128 x 10 = 40 warps per SM
// at 48 registers ahead of 256 x 4 = 32 warps at 64 registers by 3-4 % on C3 / C4, and 48 warps (40 registers) behind again;
// with round 2's per-ray slab constants (three more live registers) 128 x 9 = 36 warps at 56 registers, which keeps the leaf
// loop free of spills, is ahead of 10 by 0.7 % (C4) / 1.3 % (C3) and of 8 by 3.5 % (profiles/r2_sweep_occupancy.log).
// Build-time knobs so the sweep can be repeated (tools/build_variants.sh).
#ifndef RTB_TRAVERSE_BLOCK
#define RTB_TRAVERSE_BLOCK 128
#endif
#ifndef RTB_TRAVERSE_MIN_BLOCKS
#define RTB_TRAVERSE_MIN_BLOCKS 9
#endif
constexpr int kTravBlock = RTB_TRAVERSE_BLOCK;
// Streaming kernels that compact into queues (k_raygen, k_shade): threads per block = length of a contiguous queue run.
#ifndef RTB_STREAM_BLOCK
#define RTB_STREAM_BLOCK 256
#endif
constexpr int kStreamBlock = RTB_STREAM_BLOCK;
constexpr int kBlockSmem = 1024;  // shared-memory staged k_traverse: one block per SM
constexpr unsigned kFull = 0xffffffffu;

// slot -> pixel of the chunk.  Returns false for padding slots (outside the image or the chunk's rows).
__device__ __forceinline__ bool slot_to_pixel(const FrameParams& f, const ChunkView& c, int32_t slot, int& px, int& local_row, int& sample) {
  const int lane = slot & 31;
  const int32_t ts = slot >> 5;
  sample = ts % f.spp;
  const int32_t tile = ts / f.spp;
  const int tx = tile % c.tiles_x, ty = tile / c.tiles_x;
  px = tx * 8 + (lane & 7);
  const int r = ty * 4 + (lane >> 3);
  local_row = c.row0 + r;
  return px < f.width && r < c.rows;
}
__device__ __forceinline__ int32_t pixel_to_slot(const FrameParams& f, const ChunkView& c, int px, int r, int sample) {
  const int tile = (r >> 2) * c.tiles_x + (px >> 3);
  return ((tile * f.spp + sample) << 5) + ((r & 3) << 3) + (px & 7);
}
__device__ __forceinline__ bool primary_ray_of_slot(const FrameParams& f, const ChunkView& c, int32_t slot, Ray& ray, int& px, int& py, int& sample) {
  int local_row;
  if (!slot_to_pixel(f, c, slot, px, local_row, sample)) return false;
  py = band_global_row(local_row, f.band_rank, f.band_world, f.band_rows);
  ray = generate_ray(f, px, py, sample);
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Queue reservation for the streaming kernels (k_raygen, k_shade).  One atomicAdd per WARP on a single queue counter makes
// 10^5 same-address atomics per launch, which serialise in L2: ncu showed them as the top stall of k_shade and k_raygen
// (profiles/r1e_ncu_shade_raygen.csv).  Here the warps of a block add their counts up in shared memory and ONE thread per
// queue reserves the block's range: 8 times fewer atomics.  Every thread of the block must call it the same number of times
// (block-strided loops); `it` is the caller's iteration number (buffers alternate, so two barriers per call suffice).
// Returns the first queue index of the calling warp's entries for queue Q.
// ---------------------------------------------------------------------------------------------------------------------
template <int NQ>
struct BlockReserve {
  int32_t count[2][NQ][kStreamBlock / 32];
  int32_t base[2][NQ];
};
template <int NQ>
__device__ __forceinline__ void block_reserve(BlockReserve<NQ>& sh, int it, const unsigned (&mask)[NQ], int32_t* const (&counter)[NQ], int32_t (&first)[NQ]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, buf = it & 1;
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < NQ; q++) sh.count[buf][q][warp] = __popc(mask[q]);
  }
  __syncthreads();
  if (threadIdx.x < NQ) {
    const int q = threadIdx.x;
    int32_t total = 0;
#pragma unroll
    for (int w = 0; w < kStreamBlock / 32; w++) { const int32_t c = sh.count[buf][q][w]; sh.count[buf][q][w] = total; total += c; }
    sh.base[buf][q] = total ? atomicAdd(counter[q], total) : 0;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; q++) first[q] = sh.base[buf][q] + sh.count[buf][q][warp];
}

// ---------------------------------------------------------------------------------------------------------------------
// Work distribution of k_traverse: work items [0, n_closest) are closest-hit rays, [n_closest, n_closest + n_shadow) are
// shadow rays.  A warp owns a pool [pool_at, pool_end) claimed 32 items at a time; idle lanes take consecutive items.
// ---------------------------------------------------------------------------------------------------------------------
struct WorkPool {
  int32_t at = 0, end = 0;
  int32_t seen = 0;    // first item of this warp's latest claim: what is known of the queue's progress
  bool exhausted = false;
};
// Items reserved per atomicAdd (a multiple of 32).  Consecutive queue entries are neighbouring rays, so a warp that keeps
// drawing from one contiguous run reuses the nodes its SM's L1 already holds: larger claims raise the steady-state throughput
// (C4 +3 %, C3 +5 % at a fixed 128).  A fixed large claim lengthens the end of every launch (the last warps sit on 128 rays while the
// others idle), so the claim shrinks with what is left of the queue — "guided" scheduling: RTB_CLAIM_MAX at most, at least
// ~4 claims per resident warp remaining, never below 32.  Measured (profiles/r1e_sweep_claim*.log): guided 64 keeps the
// single-frame latency of 32 and gains 1.3-2 % throughput; guided 128 gains another 1 % but costs 4 % latency.
#ifndef RTB_CLAIM_MAX
#define RTB_CLAIM_MAX 64
#endif
__device__ __forceinline__ int32_t claim_size(int32_t remaining) {
  const int32_t warps = (int32_t)((gridDim.x * blockDim.x) >> 5);
  const int32_t c = (remaining / (warps * 4)) & ~31;
  return c < 32 ? 32 : (c > RTB_CLAIM_MAX ? RTB_CLAIM_MAX : c);
}
// Gives every lane with `want` an item index (or -1 when the work list is drained).  Warp-uniform control flow.
__device__ __forceinline__ int32_t pool_take(WorkPool& pool, int32_t* fetch_counter, int32_t total, bool want, int lane) {
  const unsigned m = __ballot_sync(kFull, want);
  const int need = __popc(m);
  if (need == 0) return -1;
  int32_t item = -1;
  const int rank = __popc(m & ((1u << lane) - 1u));
  // first serve from what is left of the pool, then claim a fresh batch for the remainder
  const int have = pool.end - pool.at;
  if (want && rank < have) item = pool.at + rank;
  if (need > have) {
    pool.at = pool.end;
    if (!pool.exhausted) {
      int32_t base = 0;
      const int32_t claim = claim_size(total - pool.seen);
      if (lane == 0) base = atomicAdd(fetch_counter, claim);
      base = __shfl_sync(kFull, base, 0);
      if (base >= total) pool.exhausted = true;
      else {
        pool.seen = base;
        pool.at = base;
        pool.end = min(base + claim, total);
        const int r2 = rank - have;
        if (want && r2 >= 0 && pool.at + r2 < pool.end) item = pool.at + r2;
        pool.at = min(pool.at + (need - have), pool.end);
      }
    }
  } else {
    pool.at += need;
  }
  return item;
}

// Lane state shared by both traversal flavours.
struct Lane {
  f3 o, d, inv;       // inv: exact reciprocal (reference mode) or safe_inverse (LBVH mode)
  float t, u, v;      // closest hit so far; for shadow rays t = nextafter(distToLight): "t < L.t" <=> "t <= distToLight"
  int32_t tri;        // closest: leaf-order triangle or -1; shadow: 0 when an occluder was found, -1 otherwise
  int32_t item;       // work item, -1 = idle
  bool shadow;
  bool done;          // traversal finished, result not yet published (published in batches, see RTB_REFILL_MIN)
};

// Lanes are refilled (and finished lanes published) only when at least this many of the warp's 32 lanes are out of work.
// Measured on the B200 (profiles/r1e_sweep_*.log).  With 128-bit node loads, 32 — a warp takes 32 fresh rays when all of
// its lanes are done — beat 8 / 16 / 24 by 4-20 %: every active lane costs the L1 data pipe one wavefront per load
// instruction, so filling the idle lanes only moved the bottleneck.  With 256-bit node loads (half the wavefronts per node
// visit) 16 wins on the global-memory variant (C3 +7 %, C4 +3 %; 8 and 2 lose again); the shared-memory variant (C2) still
// prefers 32.
#ifndef RTB_REFILL_MIN
#define RTB_REFILL_MIN 16
#endif
// Queues shorter than RTB_MIN_BATCHES 32-ray batches per resident warp leave the surplus blocks idle (0 = always use the whole grid).
#ifndef RTB_RAY_STATS
#define RTB_RAY_STATS 0  /* 1: diagnostic build, rtb_stats.reserved[3] = node visits of the frame's longest ray; RTB_TIMELINE=1 prints launch timelines */
#endif
#if RTB_RAY_STATS
__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
#ifndef RTB_MIN_BATCHES
#define RTB_MIN_BATCHES 0
#endif
template <bool SMEM>
struct RefillMin { static constexpr int value = SMEM ? 32 : RTB_REFILL_MIN; };

// Loads work item `item` into the lane.
__device__ __forceinline__ void lane_load(Lane& L, const QueueView& q, int32_t item, int32_t n_closest, int in_q) {
  L.item = item;
  L.done = false;
  L.u = 0.0f; L.v = 0.0f; L.tri = -1;
  if (item < n_closest) {
    const float4 o = __ldcs(&q.ray_o[in_q][item]), d = __ldcs(&q.ray_d[in_q][item]);
    L.shadow = false;
    L.t = RTB_INFINITY;
    L.o = mk3(o); L.d = mk3(d);
  } else {
    const int32_t j = item - n_closest;
    const float4 o = __ldcs(&q.sh_o[j]), d = __ldcs(&q.sh_d[j]);
    L.shadow = true;
    L.o = mk3(o); L.d = mk3(d);
    // compute:406 shadows iff hit.t <= distToLight; with the bound one ulp up, the closest-hit comparisons ("entry >= bound"
    // skips a box, "t < bound" accepts a triangle) decide exactly that.
    L.t = nextafterf(o.w, INFINITY);
  }
}

// A finished lane publishes its result.
__device__ __forceinline__ void lane_finish(Lane& L, const QueueView& q, int32_t n_closest) {
  if (!L.shadow) {
    __stcs(&q.hits[L.item], make_float4(L.t, L.u, L.v, __int_as_float(L.tri)));
  } else {  // lit <=> no occluder with Epsilon < t <= distToLight (compute:406)
    const int32_t j = L.item - n_closest;
    const float4 lit = __ldcs(&q.sh_lit[j]);  // (lit increment, unlit.x)
    float3 inc = make_float3(lit.x, lit.y, lit.z);
    if (L.tri == 0) { const float2 un = __ldcs(&q.sh_un[j]); inc = make_float3(lit.w, un.x, un.y); }
    const int32_t slot = __float_as_int(__ldcs(&q.sh_d[j]).w);
    const float4 prev = q.accum[slot];
    q.accum[slot] = make_float4(prev.x + inc.x, prev.y + inc.y, prev.z + inc.z, 0.0f);
  }
  L.item = -1;
  L.done = false;
}

// Stages the traversal arrays (nodes, then tri_isect) into dynamic shared memory; returns the two base pointers.
__device__ __forceinline__ void stage_scene(const SceneView& s, int node_f4, float4* sm, const float4*& nodes, const float4*& tris) {
  const int n_node = s.n_nodes * node_f4, n_tri = s.n_tris * RTB_TRI_F4;
  for (int i = threadIdx.x; i < n_node; i += blockDim.x) sm[i] = __ldg(&s.nodes[i]);
  for (int i = threadIdx.x; i < n_tri; i += blockDim.x) sm[n_node + i] = __ldg(&s.tri_isect[i]);
  __syncthreads();
  nodes = sm;
  tris = sm + n_node;
}

// Triangle test shared by both flavours.  Returns true when a shadow ray found its occluder.  TIE: see closer_hit (trace.cuh);
// a shadow lane keeps tri = -1 until it is occluded, so for it the rule reduces to "t < bound".
template <bool SMEM, bool ANALYTIC, bool TIE>
__device__ __forceinline__ bool lane_test_triangle(Lane& L, const SceneView& s, const float4* tri_isect, int32_t tri) {
  float4 a, b;
#if RTB_TRI_F4 == 4
  ld8<SMEM>(&tri_isect[RTB_TRI_F4 * tri], a, b);
#else
  const float4* rec = tri_isect + RTB_TRI_F4 * (size_t)(uint32_t)tri;  // one widening multiply-add for the record's address
  a = ld4<SMEM>(rec); b = ld4<SMEM>(rec + 1);
#endif
#if RTB_TRI_F4 == 4
  const float4 c = ld4<SMEM>(&tri_isect[RTB_TRI_F4 * tri + 2]);
#else
  const float4 c = ld4<SMEM>(rec + 2);
#endif
  Ray r; r.o = L.o; r.d = L.d;
  float t, u, v;
  if (ANALYTIC && __float_as_int(c.w) != 0) {  // analytic primitive: hit record = (t_world, t_object, face code)
    int face;
    if (!intersect_analytic(&s.prims[6 * __float_as_int(a.x)], __float_as_int(c.w), L.o, L.d, t, u, face)) return false;
    v = (float)face;
  } else if (!moller_trumbore(r, mk3(a), mk3(b), mk3(c), t, u, v)) return false;
  if (!closer_hit<TIE>(t, tri, L.t, L.tri)) return false;
  if (L.shadow) { L.tri = 0; return true; }
  L.t = t; L.u = u; L.v = v; L.tri = tri;
  return false;
}

// ---------------------------------------------------------------------------------------------------------------------
// k_traverse, LBVH flavour: ordered traversal over 64-byte two-box nodes (see lbvh.cu for the layout).
// ---------------------------------------------------------------------------------------------------------------------
template <bool SMEM, bool ANALYTIC>
__global__ void __launch_bounds__(SMEM ? kBlockSmem : kTravBlock, SMEM ? 1 : RTB_TRAVERSE_MIN_BLOCKS) k_traverse_lbvh(const SceneView s, const QueueView q, const int depth, const int mode) {
  extern __shared__ float4 sm_scene[];
  const float4* nodes = s.nodes;
  const float4* tri_isect = s.tri_isect;
  if (SMEM) stage_scene(s, lbvh_node_f4, sm_scene, nodes, tri_isect);
  const int lane = threadIdx.x & 31;
  // mode: bit 0 = serve the closest-hit rays of this depth, bit 1 = the shadow rays emitted at depth - 1 (the packet
  // kernels below may have taken either)
  const int32_t n_closest = (mode & 1) ? RTB_CNT_RAY(q, depth) : 0;
  const int32_t n_shadow = (depth == 0 || !(mode & 2)) ? 0 : RTB_CNT_SHADOW(q, depth - 1);
  const int32_t total = n_closest + n_shadow;
  const int in_q = depth & 1;
#if RTB_MIN_BATCHES > 0
  if (!SMEM) {  // short queue: fewer resident warps, each with several batches, balance better than one batch on every warp
    const int32_t want_warps = max((total + 32 * RTB_MIN_BATCHES - 1) / (32 * RTB_MIN_BATCHES), (int32_t)(gridDim.x / RTB_TRAVERSE_MIN_BLOCKS) * 4);
    if ((int32_t)(blockIdx.x * (blockDim.x >> 5)) >= want_warps && blockIdx.x != 0) return;
  }
#endif
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (depth == 0 && n_closest > 0) atomicAdd(&q.totals[RTB_TOT_ENTERED], (unsigned long long)n_closest);
    if (depth > 0 && n_closest > 0) atomicAdd(&q.totals[1], (unsigned long long)n_closest);
    if (n_shadow > 0) atomicAdd(&q.totals[2], (unsigned long long)n_shadow);
  }
  float2 stack[RTB_STACK_LBVH];  // deferred children: (entry distance, node / leaf reference) in one 8-byte local-memory access
  int sp = 0;
  int32_t cur = RTB_REF_DONE;
#if RTB_RAY_STATS
  unsigned ray_steps = 0;  // diagnostic build: node visits of the lane's current ray; totals[7] = the longest ray of the frame
  // ... and a timeline of the launch in totals[8 + 4 depth ..]: ~(first start), ~(first exhaustion seen), last end, last exhaustion seen
  unsigned long long* tl = q.totals + 8 + 4 * (depth < 16 ? depth : 15);
  bool saw_exhausted = false;
  if (lane == 0 && total > 0) atomicMax(&tl[0], ~global_ns());
#endif
  SlabRay sr;
  sr.inv = sr.ood_mn = sr.ood_mx = mk3(0.0f, 0.0f, 0.0f);
  Lane L;
  L.item = -1; L.shadow = false; L.done = false; L.t = 0.0f; L.u = 0.0f; L.v = 0.0f; L.tri = -1;
  L.o = L.d = L.inv = mk3(0.0f, 0.0f, 0.0f);
  WorkPool pool;
  unsigned overflow = 0, n_nodes = 0, n_tris = 0;

  for (;;) {
    // ---- publish finished lanes and refill, in batches ----
    const int n_out = __popc(__ballot_sync(kFull, L.item < 0 || L.done));
    if (n_out >= RefillMin<SMEM>::value) {
      if (L.item >= 0 && L.done) lane_finish(L, q, n_closest);
      const int32_t item = pool_take(pool, &RTB_CNT_FETCH(q, depth), total, L.item < 0, lane);
#if RTB_RAY_STATS
      if (pool.exhausted && !saw_exhausted) {
        saw_exhausted = true;
        if (lane == 0) { const unsigned long long now = global_ns(); atomicMax(&tl[1], ~now); atomicMax(&tl[3], now); }
      }
#endif
      if (item >= 0) {
#if RTB_RAY_STATS
        if (L.item >= 0 || ray_steps) atomicMax(&q.totals[7], (unsigned long long)ray_steps);
        ray_steps = 0;
#endif
        lane_load(L, q, item, n_closest, in_q);
        sp = 0;
        cur = s.n_tris > 0 ? s.root : RTB_REF_DONE;
        L.done = cur == RTB_REF_DONE;
        sr = make_slab_ray(L.o, L.d);
      }
      if (__ballot_sync(kFull, L.item >= 0 && !L.done) == 0) {
        if (__ballot_sync(kFull, L.item >= 0) == 0 && pool.exhausted) break;
        continue;  // only padding slots / trivially finished items were drawn: publish and take more
      }
    }

    // ---- inner nodes: descend until this lane holds a leaf (or runs out of work) ----
    while (cur >= 0) {
      n_nodes++;
#if RTB_RAY_STATS
      ray_steps++;
#endif
      // closest: a box is skipped when entry >= best t (compute:246); shadow rays carry nextafter(distToLight) as bound
      const int32_t next = lbvh_visit<SMEM>(nodes, cur, sr, L.t, stack, sp, overflow);
      if (next != RTB_REF_MISS) cur = next;
      else {
        cur = RTB_REF_DONE;
        while (sp > 0) {
          sp--;
          const float2 e = stack[sp];
          if (!(e.x > L.t)) { cur = __float_as_int(e.y); break; }
        }
      }
#if RTB_PREFETCH_LEAF
      if (!SMEM && cur < 0 && cur != RTB_REF_DONE) prefetch_leaf(tri_isect, cur);  // this lane now waits for the others: warm its triangles up
#endif
    }

    // ---- leaf ----
    if (cur != RTB_REF_DONE) {
      const int32_t code = ~cur;
      const int32_t first = code >> 3, count = (code & 7) + 1;
      bool occluded = false;
      n_tris += count;
      for (int32_t i = 0; i < count && !occluded; i++) occluded = lane_test_triangle<SMEM, ANALYTIC, true>(L, s, tri_isect, first + i);
      cur = RTB_REF_DONE;
      if (!occluded)
        while (sp > 0) {
          sp--;
          const float2 e = stack[sp];
          if (!(e.x > L.t)) { cur = __float_as_int(e.y); break; }
        }
    }
    if (cur == RTB_REF_DONE && L.item >= 0) { sp = 0; L.done = true; }
  }

#if RTB_RAY_STATS
  if (ray_steps) atomicMax(&q.totals[7], (unsigned long long)ray_steps);
  if (lane == 0 && total > 0) atomicMax(&tl[2], global_ns());
#endif
  for (int o = 16; o > 0; o >>= 1) {
    overflow += __shfl_xor_sync(kFull, overflow, o);
    n_nodes += __shfl_xor_sync(kFull, n_nodes, o);
    n_tris += __shfl_xor_sync(kFull, n_tris, o);
  }
  if (lane == 0) {
    if (overflow) atomicAdd(&q.totals[4], (unsigned long long)overflow);
    if (n_nodes) atomicAdd(&q.totals[5], (unsigned long long)n_nodes);
    if (n_tris) atomicAdd(&q.totals[6], (unsigned long long)n_tris);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_traverse, reference flavour: TraverseBVH compute:225-267 — LIFO stack, left child first, no distance ordering, node
// culled when its own slab entry >= best t; leaves of any size.  Same results as the reference for every ray.
// ---------------------------------------------------------------------------------------------------------------------
template <bool SMEM, bool ANALYTIC>
__global__ void __launch_bounds__(SMEM ? kBlockSmem : kTravBlock, SMEM ? 1 : RTB_TRAVERSE_MIN_BLOCKS) k_traverse_ref(const SceneView s, const QueueView q, const int depth, const int mode) {
  extern __shared__ float4 sm_scene[];
  const float4* nodes = s.nodes;
  const float4* tri_isect = s.tri_isect;
  if (SMEM) stage_scene(s, 2, sm_scene, nodes, tri_isect);
  const int lane = threadIdx.x & 31;
  // mode: bit 0 = serve the closest-hit rays of this depth, bit 1 = the shadow rays emitted at depth - 1 (the packet
  // kernels below may have taken either)
  const int32_t n_closest = (mode & 1) ? RTB_CNT_RAY(q, depth) : 0;
  const int32_t n_shadow = (depth == 0 || !(mode & 2)) ? 0 : RTB_CNT_SHADOW(q, depth - 1);
  const int32_t total = n_closest + n_shadow;
  const int in_q = depth & 1;
#if RTB_MIN_BATCHES > 0
  if (!SMEM) {  // short queue: fewer resident warps, each with several batches, balance better than one batch on every warp
    const int32_t want_warps = max((total + 32 * RTB_MIN_BATCHES - 1) / (32 * RTB_MIN_BATCHES), (int32_t)(gridDim.x / RTB_TRAVERSE_MIN_BLOCKS) * 4);
    if ((int32_t)(blockIdx.x * (blockDim.x >> 5)) >= want_warps && blockIdx.x != 0) return;
  }
#endif
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (depth == 0 && n_closest > 0) atomicAdd(&q.totals[RTB_TOT_ENTERED], (unsigned long long)n_closest);
    if (depth > 0 && n_closest > 0) atomicAdd(&q.totals[1], (unsigned long long)n_closest);
    if (n_shadow > 0) atomicAdd(&q.totals[2], (unsigned long long)n_shadow);
  }
  int32_t stack[RTB_STACK_REF];
  int sp = 0;
  int32_t leaf_first = 0, leaf_count = 0;
  Lane L;
  L.item = -1; L.shadow = false; L.done = false; L.t = 0.0f; L.u = 0.0f; L.v = 0.0f; L.tri = -1;
  L.o = L.d = L.inv = mk3(0.0f, 0.0f, 0.0f);
  WorkPool pool;
  unsigned overflow = 0, n_nodes = 0, n_tris = 0;

  for (;;) {
    const int n_out = __popc(__ballot_sync(kFull, L.item < 0 || L.done));
    if (n_out >= RefillMin<SMEM>::value) {
      if (L.item >= 0 && L.done) lane_finish(L, q, n_closest);
      const int32_t item = pool_take(pool, &RTB_CNT_FETCH(q, depth), total, L.item < 0, lane);
      if (item >= 0) {
        lane_load(L, q, item, n_closest, in_q);
        L.inv = mk3(1.0f / L.d.x, 1.0f / L.d.y, 1.0f / L.d.z);  // CreateRay :142 / :398
        sp = 0; leaf_count = 0;
        if (s.n_nodes > 0) stack[sp++] = 0;
        else L.done = true;
      }
      if (__ballot_sync(kFull, L.item >= 0 && !L.done) == 0) {
        if (__ballot_sync(kFull, L.item >= 0) == 0 && pool.exhausted) break;
        continue;
      }
    }

    while (leaf_count == 0 && sp > 0) {
      const int32_t ni = stack[--sp];
      n_nodes++;
      const float4 lo = ld4<SMEM>(&nodes[2 * ni]), hi = ld4<SMEM>(&nodes[2 * ni + 1]);
      Ray r; r.o = L.o; r.d = L.d; r.inv = L.inv;
      const float dst = slab_entry(r, mk3(lo), mk3(hi));
      if (dst >= L.t) continue;  // compute:246
      const int32_t count = __float_as_int(hi.w), left_or_first = __float_as_int(lo.w);
      if (count > 0) { leaf_first = left_or_first; leaf_count = count; }
      else if (sp + 2 <= RTB_STACK_REF) { stack[sp++] = left_or_first + 1; stack[sp++] = left_or_first; }
      else overflow++;
    }
    if (leaf_count > 0) {
      bool occluded = false;
      n_tris += leaf_count;
      for (int32_t i = 0; i < leaf_count && !occluded; i++) occluded = lane_test_triangle<SMEM, ANALYTIC, false>(L, s, tri_isect, leaf_first + i);
      leaf_count = 0;
      if (occluded) sp = 0;
    }
    if (sp == 0 && leaf_count == 0 && L.item >= 0) L.done = true;
  }

  for (int o = 16; o > 0; o >>= 1) {
    overflow += __shfl_xor_sync(kFull, overflow, o);
    n_nodes += __shfl_xor_sync(kFull, n_nodes, o);
    n_tris += __shfl_xor_sync(kFull, n_tris, o);
  }
  if (lane == 0) {
    if (overflow) atomicAdd(&q.totals[4], (unsigned long long)overflow);
    if (n_nodes) atomicAdd(&q.totals[5], (unsigned long long)n_nodes);
    if (n_tris) atomicAdd(&q.totals[6], (unsigned long long)n_tris);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// k_traverse, wide flavour: the same persistent per-lane loop over the 8-wide quantised records (trace.cuh: wide_visit).
// ---------------------------------------------------------------------------------------------------------------------
#ifndef RTB_WIDE_MIN_BLOCKS
#define RTB_WIDE_MIN_BLOCKS 8
#endif
template <bool SMEM, bool ANALYTIC>
__global__ void __launch_bounds__(SMEM ? kBlockSmem : kTravBlock, SMEM ? 1 : RTB_WIDE_MIN_BLOCKS) k_traverse_wide(const SceneView s, const QueueView q, const int depth, const int mode) {
  extern __shared__ float4 sm_scene[];
  const float4* nodes = s.nodes;
  const float4* tri_isect = s.tri_isect;
  if (SMEM) stage_scene(s, RTB_WIDE_F4, sm_scene, nodes, tri_isect);
  const int lane = threadIdx.x & 31;
  const int32_t n_closest = (mode & 1) ? RTB_CNT_RAY(q, depth) : 0;
  const int32_t n_shadow = (depth == 0 || !(mode & 2)) ? 0 : RTB_CNT_SHADOW(q, depth - 1);
  const int32_t total = n_closest + n_shadow;
  const int in_q = depth & 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (depth == 0 && n_closest > 0) atomicAdd(&q.totals[RTB_TOT_ENTERED], (unsigned long long)n_closest);
    if (depth > 0 && n_closest > 0) atomicAdd(&q.totals[1], (unsigned long long)n_closest);
    if (n_shadow > 0) atomicAdd(&q.totals[2], (unsigned long long)n_shadow);
  }
  uint2 stack[RTB_STACK_WIDE];  // deferred sibling groups: (node, octant-ordered hit mask)
  int sp = 0;
  int32_t cur = RTB_REF_DONE;
  unsigned neg = 0u;
  Lane L;
  L.item = -1; L.shadow = false; L.done = false; L.t = 0.0f; L.u = 0.0f; L.v = 0.0f; L.tri = -1;
  L.o = L.d = L.inv = mk3(0.0f, 0.0f, 0.0f);
  WorkPool pool;
  unsigned overflow = 0, n_nodes = 0, n_tris = 0;

  for (;;) {
    const int n_out = __popc(__ballot_sync(kFull, L.item < 0 || L.done));
    if (n_out >= RefillMin<SMEM>::value) {
      if (L.item >= 0 && L.done) lane_finish(L, q, n_closest);
      const int32_t item = pool_take(pool, &RTB_CNT_FETCH(q, depth), total, L.item < 0, lane);
      if (item >= 0) {
        lane_load(L, q, item, n_closest, in_q);
        sp = 0;
        cur = s.n_tris > 0 ? s.root : RTB_REF_DONE;
        L.done = cur == RTB_REF_DONE;
        L.inv = safe_inverse(L.d);
        neg = octant_of(L.d);
      }
      if (__ballot_sync(kFull, L.item >= 0 && !L.done) == 0) {
        if (__ballot_sync(kFull, L.item >= 0) == 0 && pool.exhausted) break;
        continue;
      }
    }

    // ---- inner nodes: descend until this lane holds a leaf (or runs out of work) ----
    while (cur >= 0) {
      n_nodes++;
      cur = wide_visit<SMEM>(nodes, cur, L.o, L.inv, neg, L.t, stack, sp, overflow);
    }

    // ---- leaf ----
    if (cur != RTB_REF_DONE) {
      const int32_t code = ~cur;
      const int32_t first = code >> 3, count = (code & 7) + 1;
      bool occluded = false;
      n_tris += count;
      for (int32_t i = 0; i < count && !occluded; i++) occluded = lane_test_triangle<SMEM, ANALYTIC, true>(L, s, tri_isect, first + i);
      cur = occluded ? RTB_REF_DONE : wide_pop<SMEM>(nodes, neg, stack, sp);
    }
    if (cur == RTB_REF_DONE && L.item >= 0) { sp = 0; L.done = true; }
  }

  for (int o = 16; o > 0; o >>= 1) {
    overflow += __shfl_xor_sync(kFull, overflow, o);
    n_nodes += __shfl_xor_sync(kFull, n_nodes, o);
    n_tris += __shfl_xor_sync(kFull, n_tris, o);
  }
  if (lane == 0) {
    if (overflow) atomicAdd(&q.totals[4], (unsigned long long)overflow);
    if (n_nodes) atomicAdd(&q.totals[5], (unsigned long long)n_nodes);
    if (n_tris) atomicAdd(&q.totals[6], (unsigned long long)n_tris);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_traverse_pool: the per-lane kernel's answer to its own profile — 13 of 32 lanes live, because a warp's lanes wait for each
// other at every phase change (all descend, then all intersect) while every ray's steps are strictly sequential.  Here a warp
// keeps RTB_POOL_SLOTS = 64 rays in flight, their state in shared memory and their stacks in a slot-interleaved global scratch
// area, and REGROUPS: in the node phase the 32 lanes take 32 slots that need a node step, run until fewer than kPoolRunMin of
// them still descend, write back and pick again, until no slot needs a node step; then the same over the slots holding a leaf.
// Same node visits, same triangle tests, same per-ray order as k_traverse_lbvh — only which lane does them changes.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef RTB_POOL_MIN_BLOCKS
#define RTB_POOL_MIN_BLOCKS 6
#endif
#ifndef RTB_POOL_RUN_MIN
#define RTB_POOL_RUN_MIN 20
#endif
#ifndef RTB_POOL_REFILL
#define RTB_POOL_REFILL 16
#endif
constexpr int kPoolBlock = 128;
enum { PW_INV = 0, PW_OMN = 3, PW_OMX = 6, PW_O = 9, PW_D = 12, PW_T = 15, PW_U = 16, PW_V = 17, PW_TRI = 18, PW_ITEM = 19, PW_CUR = 20, PW_SP = 21, PW_WORDS = 22 };

// Lanes [0, n) receive the slots whose bit is set in (lo, hi), lowest slot first (at most 32); returns the lane's slot or -1.
__device__ __forceinline__ int pool_assign(unsigned lo, unsigned hi, int* sel, int lane, bool& backlog) {
  const unsigned lt = (1u << lane) - 1u;
  const int n_lo = __popc(lo), n_all = n_lo + __popc(hi);
  __syncwarp();
  if ((lo >> lane) & 1u) sel[__popc(lo & lt)] = lane;
  const int r_hi = n_lo + __popc(hi & lt);
  if (((hi >> lane) & 1u) && r_hi < 32) sel[r_hi] = lane + 32;
  __syncwarp();
  backlog = n_all > 32;
  return lane < (n_all < 32 ? n_all : 32) ? sel[lane] : -1;
}

template <bool ANALYTIC>
__global__ void __launch_bounds__(kPoolBlock, RTB_POOL_MIN_BLOCKS) k_traverse_pool(const SceneView s, const QueueView q, const int depth, const int mode,
                                                                                    float2* __restrict__ scratch) {
  __shared__ float sm_state[kPoolBlock / 32][PW_WORDS][RTB_POOL_SLOTS];
  __shared__ int sm_sel[kPoolBlock / 32][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float (*S)[RTB_POOL_SLOTS] = sm_state[wib];
  int* sel = sm_sel[wib];
  float2* const stk_base = scratch + ((size_t)blockIdx.x * (kPoolBlock / 32) + (size_t)wib) * (size_t)(RTB_STACK_LBVH * RTB_POOL_SLOTS);
  const float4* nodes = s.nodes;
  const float4* tri_isect = s.tri_isect;
  const int32_t n_closest = (mode & 1) ? RTB_CNT_RAY(q, depth) : 0;
  const int32_t n_shadow = (depth == 0 || !(mode & 2)) ? 0 : RTB_CNT_SHADOW(q, depth - 1);
  const int32_t total = n_closest + n_shadow;
  const int in_q = depth & 1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (depth == 0 && n_closest > 0) atomicAdd(&q.totals[RTB_TOT_ENTERED], (unsigned long long)n_closest);
    if (depth > 0 && n_closest > 0) atomicAdd(&q.totals[1], (unsigned long long)n_closest);
    if (n_shadow > 0) atomicAdd(&q.totals[2], (unsigned long long)n_shadow);
  }
  S[PW_ITEM][lane] = __int_as_float(-1); S[PW_ITEM][lane + 32] = __int_as_float(-1);
  S[PW_CUR][lane] = __int_as_float(RTB_REF_DONE); S[PW_CUR][lane + 32] = __int_as_float(RTB_REF_DONE);
  WorkPool pool;
  unsigned overflow = 0, n_nodes = 0, n_tris = 0;

  for (;;) {
    // ---- publish finished slots, refill free ones (each lane looks after slots lane and lane + 32) ----
    __syncwarp();
    bool is_free[2];
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const int sidx = lane + 32 * b;
      int32_t item = __float_as_int(S[PW_ITEM][sidx]);
      if (item >= 0 && __float_as_int(S[PW_CUR][sidx]) == RTB_REF_DONE) {
        Lane L;
        L.item = item; L.shadow = item >= n_closest; L.done = true;
        L.t = S[PW_T][sidx]; L.u = S[PW_U][sidx]; L.v = S[PW_V][sidx]; L.tri = __float_as_int(S[PW_TRI][sidx]);
        L.o = L.d = L.inv = mk3(0.0f, 0.0f, 0.0f);
        lane_finish(L, q, n_closest);
        item = -1;
        S[PW_ITEM][sidx] = __int_as_float(-1);
      }
      is_free[b] = item < 0;
    }
    const int n_free = __popc(__ballot_sync(kFull, is_free[0])) + __popc(__ballot_sync(kFull, is_free[1]));
    if (n_free >= RTB_POOL_REFILL && !pool.exhausted) {
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int sidx = lane + 32 * b;
        const int32_t item = pool_take(pool, &RTB_CNT_FETCH(q, depth), total, is_free[b], lane);
        if (item >= 0) {
          Lane L;
          lane_load(L, q, item, n_closest, in_q);
          const SlabRay sr = make_slab_ray(L.o, L.d);
          S[PW_INV][sidx] = sr.inv.x; S[PW_INV + 1][sidx] = sr.inv.y; S[PW_INV + 2][sidx] = sr.inv.z;
          S[PW_OMN][sidx] = sr.ood_mn.x; S[PW_OMN + 1][sidx] = sr.ood_mn.y; S[PW_OMN + 2][sidx] = sr.ood_mn.z;
          S[PW_OMX][sidx] = sr.ood_mx.x; S[PW_OMX + 1][sidx] = sr.ood_mx.y; S[PW_OMX + 2][sidx] = sr.ood_mx.z;
          S[PW_O][sidx] = L.o.x; S[PW_O + 1][sidx] = L.o.y; S[PW_O + 2][sidx] = L.o.z;
          S[PW_D][sidx] = L.d.x; S[PW_D + 1][sidx] = L.d.y; S[PW_D + 2][sidx] = L.d.z;
          S[PW_T][sidx] = L.t; S[PW_U][sidx] = 0.0f; S[PW_V][sidx] = 0.0f; S[PW_TRI][sidx] = __int_as_float(-1);
          S[PW_ITEM][sidx] = __int_as_float(item);
          S[PW_CUR][sidx] = __int_as_float(s.n_tris > 0 ? s.root : RTB_REF_DONE);
          S[PW_SP][sidx] = __int_as_float(0);
          is_free[b] = false;
        }
      }
    }
    if ((__ballot_sync(kFull, !is_free[0]) | __ballot_sync(kFull, !is_free[1])) == 0u) {
      if (pool.exhausted) break;
      continue;
    }

    // ---- node phase: until no occupied slot needs a node step ----
    for (;;) {
      __syncwarp();
      const int32_t c0 = __float_as_int(S[PW_CUR][lane]), c1 = __float_as_int(S[PW_CUR][lane + 32]);
      const unsigned lo = __ballot_sync(kFull, __float_as_int(S[PW_ITEM][lane]) >= 0 && c0 >= 0);
      const unsigned hi = __ballot_sync(kFull, __float_as_int(S[PW_ITEM][lane + 32]) >= 0 && c1 >= 0);
      if ((lo | hi) == 0u) break;
      bool backlog;
      const int my = pool_assign(lo, hi, sel, lane, backlog);
      SlabRay sr;
      sr.inv = sr.ood_mn = sr.ood_mx = mk3(0.0f, 0.0f, 0.0f);
      float t = 0.0f;
      int32_t cur = RTB_REF_DONE;
      int sp = 0;
      float2* stk = stk_base;
      if (my >= 0) {
        sr.inv = mk3(S[PW_INV][my], S[PW_INV + 1][my], S[PW_INV + 2][my]);
        sr.ood_mn = mk3(S[PW_OMN][my], S[PW_OMN + 1][my], S[PW_OMN + 2][my]);
        sr.ood_mx = mk3(S[PW_OMX][my], S[PW_OMX + 1][my], S[PW_OMX + 2][my]);
        t = S[PW_T][my];
        cur = __float_as_int(S[PW_CUR][my]);
        sp = __float_as_int(S[PW_SP][my]);
        stk = stk_base + my;
      }
      for (;;) {
        if (cur >= 0) {
          n_nodes++;
          const int32_t next = lbvh_visit<false, RTB_POOL_SLOTS>(nodes, cur, sr, t, stk, sp, overflow);
          if (next != RTB_REF_MISS) cur = next;
          else {
            cur = RTB_REF_DONE;
            while (sp > 0) {
              sp--;
              const float2 e = stk[(size_t)sp * RTB_POOL_SLOTS];
              if (!(e.x > t)) { cur = __float_as_int(e.y); break; }
            }
          }
        }
        const int n_act = __popc(__ballot_sync(kFull, cur >= 0));
        if (n_act == 0 || (backlog && n_act < RTB_POOL_RUN_MIN)) break;
      }
      if (my >= 0) { S[PW_CUR][my] = __int_as_float(cur); S[PW_SP][my] = __int_as_float(sp); }
    }

    // ---- leaf phase: until no occupied slot holds a leaf ----
    for (;;) {
      __syncwarp();
      const int32_t c0 = __float_as_int(S[PW_CUR][lane]), c1 = __float_as_int(S[PW_CUR][lane + 32]);
      const unsigned lo = __ballot_sync(kFull, __float_as_int(S[PW_ITEM][lane]) >= 0 && c0 < 0 && c0 != RTB_REF_DONE);
      const unsigned hi = __ballot_sync(kFull, __float_as_int(S[PW_ITEM][lane + 32]) >= 0 && c1 < 0 && c1 != RTB_REF_DONE);
      if ((lo | hi) == 0u) break;
      bool backlog;
      const int my = pool_assign(lo, hi, sel, lane, backlog);
      Lane L;
      L.item = -1; L.shadow = false; L.done = false; L.t = 0.0f; L.u = 0.0f; L.v = 0.0f; L.tri = -1;
      L.o = L.d = L.inv = mk3(0.0f, 0.0f, 0.0f);
      int32_t cur = RTB_REF_DONE;
      int sp = 0;
      float2* stk = stk_base;
      if (my >= 0) {
        L.o = mk3(S[PW_O][my], S[PW_O + 1][my], S[PW_O + 2][my]);
        L.d = mk3(S[PW_D][my], S[PW_D + 1][my], S[PW_D + 2][my]);
        L.t = S[PW_T][my]; L.u = S[PW_U][my]; L.v = S[PW_V][my]; L.tri = __float_as_int(S[PW_TRI][my]);
        L.item = __float_as_int(S[PW_ITEM][my]);
        L.shadow = L.item >= n_closest;
        cur = __float_as_int(S[PW_CUR][my]);
        sp = __float_as_int(S[PW_SP][my]);
        stk = stk_base + my;
      }
      for (;;) {
        if (cur < 0 && cur != RTB_REF_DONE) {
          const int32_t code = ~cur;
          const int32_t first = code >> 3, count = (code & 7) + 1;
          bool occluded = false;
          n_tris += count;
          for (int32_t i = 0; i < count && !occluded; i++) occluded = lane_test_triangle<false, ANALYTIC, true>(L, s, tri_isect, first + i);
          cur = RTB_REF_DONE;
          if (occluded) sp = 0;
          while (sp > 0) {
            sp--;
            const float2 e = stk[(size_t)sp * RTB_POOL_SLOTS];
            if (!(e.x > L.t)) { cur = __float_as_int(e.y); break; }
          }
        }
        const int n_leaf = __popc(__ballot_sync(kFull, cur < 0 && cur != RTB_REF_DONE));
        if (n_leaf == 0 || (backlog && n_leaf < RTB_POOL_RUN_MIN)) break;
      }
      if (my >= 0) {
        S[PW_T][my] = L.t; S[PW_U][my] = L.u; S[PW_V][my] = L.v; S[PW_TRI][my] = __int_as_float(L.tri);
        S[PW_CUR][my] = __int_as_float(cur); S[PW_SP][my] = __int_as_float(sp);
      }
    }
  }

  for (int o = 16; o > 0; o >>= 1) {
    overflow += __shfl_xor_sync(kFull, overflow, o);
    n_nodes += __shfl_xor_sync(kFull, n_nodes, o);
    n_tris += __shfl_xor_sync(kFull, n_tris, o);
  }
  if (lane == 0) {
    if (overflow) atomicAdd(&q.totals[4], (unsigned long long)overflow);
    if (n_nodes) atomicAdd(&q.totals[5], (unsigned long long)n_nodes);
    if (n_tris) atomicAdd(&q.totals[6], (unsigned long long)n_tris);
  }
}

// =====================================================================================================================
// Packet traversal: a warp walks the BVH ONCE for its 32 rays.  Rays that are neighbours on the screen (primary rays of an
// 8x4 tile, the shadow rays those pixels emit towards the one light) visit nearly the same nodes, so the per-lane kernels
// above fetch every node record 32 times from 32 divergent addresses (one L1 wavefront each) and idle through each other's
// steps (13 of 32 lanes live, profiles/r1e_ncu_full_k_traverse_c4.csv).  Here the node / triangle address is warp-uniform
// (one broadcast wavefront per load), the stack is one per warp in shared memory, and every lane tests its own ray against
// the shared record with the same box / triangle arithmetic as the per-lane kernels, so per-ray results are the same:
//   reference flavour: TraverseBVH's order (compute:235-264) does not depend on the ray, so the packet follows it exactly and
//     each lane sees its own subsequence of it (a lane skips a node when ITS slab test culls it, :245-246) — same tie winners;
//   LBVH flavour: near child first by majority vote; the closest hit is order-independent by closer_hit<true>.
// =====================================================================================================================
#define RTB_PSTACK 96
struct PacketStack {  // one per warp, shared memory
  int32_t ref[RTB_PSTACK];
  unsigned mask[RTB_PSTACK];   // LBVH: lanes whose ray entered the deferred child's box
  float nearest[RTB_PSTACK];   // LBVH: smallest entry distance among them
};

// L: o, d, t = bound (closest: Infinity; shadow: nextafter(distToLight)), tri = -1, shadow set.  `valid` = the lane holds a ray.
// On return L.t/u/v/tri hold the closest hit, or for a shadow ray L.tri = 0 iff an occluder was found.
template <bool ANALYTIC>
__device__ __forceinline__ void packet_traverse_lbvh(const SceneView& s, Lane& L, const bool valid, PacketStack& st, unsigned& overflow, unsigned& n_nodes,
                                                     unsigned& n_tris, unsigned& w_nodes, unsigned& w_tris) {
  const int lane = threadIdx.x & 31;
  if (s.n_tris == 0) return;
  const SlabRay sr = make_slab_ray(L.o, L.d);
  const float t_saved = L.t;
  if (!valid) L.t = -1.0f;  // every box test fails (exit <= bound < 0 <= entry)
  int sp = 0;
  int32_t cur = s.root;
  bool act = valid;
  for (;;) {
    if (cur >= 0) {
      float4 n0, n1, n2, n3;  // warp-uniform address: broadcast
      ld8<false>(&s.nodes[4 * (size_t)cur], n0, n1);
      ld8<false>(&s.nodes[4 * (size_t)cur + 2], n2, n3);
      float dl, dr;
      const bool hl = act && slab_hit(sr, mk3(n0), mk3(n1), L.t, dl);
      const bool hr = act && slab_hit(sr, mk3(n2), mk3(n3), L.t, dr);
      const unsigned ml = __ballot_sync(kFull, hl), mr = __ballot_sync(kFull, hr);
      n_nodes += act ? 1u : 0u;
      w_nodes += lane == 0 ? 1u : 0u;
      const int32_t lref = __float_as_int(n0.w), rref = __float_as_int(n1.w);
      if (ml != 0 && mr != 0) {
        const int votes_l = __popc(__ballot_sync(kFull, hl && (!hr || !(dr < dl))));
        const int votes_r = __popc(__ballot_sync(kFull, hr && (!hl || dr < dl)));
        const bool left_first = votes_l >= votes_r;
        const float d_far = left_first ? (hr ? dr : INFINITY) : (hl ? dl : INFINITY);
        const float far_min = __uint_as_float(__reduce_min_sync(kFull, __float_as_uint(d_far)));  // entries are >= 0: they order like their bits
        if (sp < RTB_PSTACK) {
          if (lane == 0) { st.ref[sp] = left_first ? rref : lref; st.mask[sp] = left_first ? mr : ml; st.nearest[sp] = far_min; }
          sp++;
          __syncwarp();
        } else if (lane == 0) overflow++;
        cur = left_first ? lref : rref;
        act = left_first ? hl : hr;
        continue;
      }
      if (ml != 0) { cur = lref; act = hl; continue; }
      if (mr != 0) { cur = rref; act = hr; continue; }
    } else {
      const int32_t code = ~cur;
      const int32_t first = code >> 3, count = (code & 7) + 1;
      w_tris += lane == 0 ? (unsigned)count : 0u;
      n_tris += act ? (unsigned)count : 0u;
      for (int32_t i = 0; i < count; i++)
        if (act && lane_test_triangle<false, ANALYTIC, true>(L, s, s.tri_isect, first + i)) { L.t = -1.0f; act = false; }  // shadow ray occluded: drops out
    }
    // next deferred child some lane still needs
    cur = RTB_REF_DONE;
    while (sp > 0) {
      sp--;
      const bool want = ((st.mask[sp] >> lane) & 1u) != 0u && !(st.nearest[sp] > L.t);
      if (__any_sync(kFull, want)) { cur = st.ref[sp]; act = want; break; }
    }
    if (cur == RTB_REF_DONE) break;
  }
  if (!valid) L.t = t_saved;
}

template <bool ANALYTIC>
__device__ __forceinline__ void packet_traverse_ref(const SceneView& s, Lane& L, const bool valid, PacketStack& st, unsigned& overflow, unsigned& n_nodes,
                                                    unsigned& n_tris, unsigned& w_nodes, unsigned& w_tris) {
  const int lane = threadIdx.x & 31;
  if (s.n_nodes == 0) return;
  Ray r; r.o = L.o; r.d = L.d; r.inv = mk3(1.0f / L.d.x, 1.0f / L.d.y, 1.0f / L.d.z);  // CreateRay :142
  int sp = 0;
  if (lane == 0) st.ref[0] = 0;
  sp = 1;
  __syncwarp();
  bool live = valid;  // a shadow ray that found its occluder stops (the per-lane kernel empties its stack)
  while (sp > 0) {
    const int32_t ni = st.ref[--sp];
    float4 lo, hi;
    ld8<false>(&s.nodes[2 * (size_t)ni], lo, hi);
    const float dst = slab_entry(r, mk3(lo), mk3(hi));
    const bool want = live && !(dst >= L.t);  // compute:246
    w_nodes += lane == 0 ? 1u : 0u;
    if (!__any_sync(kFull, want)) continue;
    n_nodes += want ? 1u : 0u;
    const int32_t count = __float_as_int(hi.w), left_or_first = __float_as_int(lo.w);
    if (count > 0) {
      w_tris += lane == 0 ? (unsigned)count : 0u;
      n_tris += want ? (unsigned)count : 0u;
      for (int32_t i = 0; i < count; i++)
        if (want && live && lane_test_triangle<false, ANALYTIC, false>(L, s, s.tri_isect, left_or_first + i)) live = false;
    } else if (sp + 2 <= RTB_PSTACK) {
      __syncwarp();  // every lane has read the entry this push may overwrite
      if (lane == 0) { st.ref[sp] = left_or_first + 1; st.ref[sp + 1] = left_or_first; }
      sp += 2;
      __syncwarp();
    } else if (lane == 0) overflow++;
  }
}

template <int BVH, bool ANALYTIC>
__device__ __forceinline__ void packet_traverse(const SceneView& s, Lane& L, const bool valid, PacketStack& st, unsigned& overflow, unsigned& n_nodes,
                                                unsigned& n_tris, unsigned& w_nodes, unsigned& w_tris) {
  if (BVH == RTB_BVH_REFERENCE) packet_traverse_ref<ANALYTIC>(s, L, valid, st, overflow, n_nodes, n_tris, w_nodes, w_tris);
  else packet_traverse_lbvh<ANALYTIC>(s, L, valid, st, overflow, n_nodes, n_tris, w_nodes, w_tris);
}

__device__ __forceinline__ void packet_counters(const QueueView& q, unsigned overflow, unsigned n_nodes, unsigned n_tris, unsigned w_nodes, unsigned w_tris) {
  for (int o = 16; o > 0; o >>= 1) {
    n_nodes += __shfl_xor_sync(kFull, n_nodes, o);
    n_tris += __shfl_xor_sync(kFull, n_tris, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (overflow) atomicAdd(&q.totals[4], (unsigned long long)overflow);
    if (n_nodes) atomicAdd(&q.totals[5], (unsigned long long)n_nodes);
    if (n_tris) atomicAdd(&q.totals[6], (unsigned long long)n_tris);
    if (w_nodes) atomicAdd(&q.totals[RTB_TOT_PACKET_NODES], (unsigned long long)w_nodes);
    if (w_tris) atomicAdd(&q.totals[RTB_TOT_PACKET_TRIS], (unsigned long long)w_tris);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_primary (K3 + K4 fused): primary rays of the chunk (CSMain compute:283-349) traced as packets, one warp per 8x4 pixel
// tile and sample.  A ray that hits nothing takes the background at once (:364-368: sampleColor = 0 + attenuation * bg with
// attenuation (1,1,1)); the hits are compacted into the depth-0 ray queue together with their hit records, so k_shade(0)
// only sees paths it has to shade.  Replaces k_raygen + k_traverse(depth 0); no ray makes a round trip through HBM before
// its first traversal.
// ---------------------------------------------------------------------------------------------------------------------
template <int BVH, bool ANALYTIC>
__global__ void __launch_bounds__(kStreamBlock) k_primary(const FrameParams f, const SceneView s, const QueueView q, const ChunkView c) {
  __shared__ PacketStack stacks[kStreamBlock / 32];
  __shared__ BlockReserve<1> reserve;
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  unsigned n_valid = 0, overflow = 0, n_nodes = 0, n_tris = 0, w_nodes = 0, w_tris = 0;
  int it = 0;
  for (int32_t base = blockIdx.x * kStreamBlock; base < c.n_slots; base += gridDim.x * kStreamBlock, it++) {  // block-uniform trip count
    const int32_t slot = base + threadIdx.x;
    Ray ray;
    int px, py, sample;
    const bool valid = primary_ray_of_slot(f, c, slot, ray, px, py, sample);
    Lane L;
    L.item = slot; L.shadow = false; L.done = false; L.t = RTB_INFINITY; L.u = 0.0f; L.v = 0.0f; L.tri = -1;
    L.o = valid ? ray.o : mk3(0.0f, 0.0f, 0.0f);
    L.d = valid ? ray.d : mk3(0.0f, 0.0f, 1.0f);
    L.inv = mk3(0.0f, 0.0f, 0.0f);
    if (__any_sync(kFull, valid)) packet_traverse<BVH, ANALYTIC>(s, L, valid, stacks[threadIdx.x >> 5], overflow, n_nodes, n_tris, w_nodes, w_tris);
    const bool found = valid && L.tri >= 0;
    if (valid) {
      n_valid++;
      const f3 bg = mk3(1.0f, 1.0f, 1.0f) * mk3(f.bg[0], f.bg[1], f.bg[2]);
      const f3 first = found ? mk3(0.0f, 0.0f, 0.0f) : mk3(0.0f, 0.0f, 0.0f) + bg;
      q.accum[slot] = make_float4(first.x, first.y, first.z, 0.0f);
    }
    const unsigned m[1] = {__ballot_sync(kFull, found)};
    int32_t* const counters[1] = {&RTB_CNT_RAY(q, 0)};
    int32_t first_at[1];
    block_reserve<1>(reserve, it, m, counters, first_at);
    if (found) {
      const int32_t at = first_at[0] + __popc(m[0] & below);
      __stcs(&q.ray_o[0][at], make_float4(ray.o.x, ray.o.y, ray.o.z, __int_as_float(slot)));
      __stcs(&q.ray_d[0][at], make_float4(ray.d.x, ray.d.y, ray.d.z, 1.0f));
      __stcs(&q.ray_a[0][at], make_float2(1.0f, 1.0f));
      __stcs(&q.hits[at], make_float4(L.t, L.u, L.v, __int_as_float(L.tri)));
    }
  }
  for (int o = 16; o > 0; o >>= 1) n_valid += __shfl_xor_sync(kFull, n_valid, o);
  if (lane == 0 && n_valid) { atomicAdd(&q.totals[0], (unsigned long long)n_valid); atomicAdd(&q.totals[RTB_TOT_ENTERED], (unsigned long long)n_valid); }
  packet_counters(q, overflow, n_nodes, n_tris, w_nodes, w_tris);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_packet: 32 consecutive entries of a queue as one packet — the closest-hit rays of `depth` (kind 0: writes their hit
// records) or the shadow rays emitted at `depth` (kind 1: adds the lit or unlit increment of :418, like lane_finish).
// Queue order follows screen order (k_primary / k_shade compact in runs of 256), so neighbours in the queue are neighbours
// on the screen as long as the surfaces they left are; api.cu uses this for the depths given by RTB_PACKET_CLOSEST /
// RTB_PACKET_SHADOW and the per-lane kernels beyond.
// ---------------------------------------------------------------------------------------------------------------------
template <int BVH, bool ANALYTIC>
__global__ void __launch_bounds__(kStreamBlock) k_packet(const SceneView s, const QueueView q, const int depth, const int kind) {
  __shared__ PacketStack stacks[kStreamBlock / 32];
  const int lane = threadIdx.x & 31;
  const int32_t n = kind == 0 ? RTB_CNT_RAY(q, depth) : RTB_CNT_SHADOW(q, depth);
  const int in_q = depth & 1;
  if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) {
    if (kind == 0) { if (depth > 0) atomicAdd(&q.totals[1], (unsigned long long)n); }
    else atomicAdd(&q.totals[2], (unsigned long long)n);
  }
  unsigned overflow = 0, n_nodes = 0, n_tris = 0, w_nodes = 0, w_tris = 0;
  const int32_t warps = (int32_t)((gridDim.x * blockDim.x) >> 5);
  for (int32_t base = ((int32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5)) * 32; base < n; base += warps * 32) {
    const int32_t idx = base + lane;
    const bool valid = idx < n;
    Lane L;
    L.item = idx; L.shadow = kind != 0; L.done = false; L.t = RTB_INFINITY; L.u = 0.0f; L.v = 0.0f; L.tri = -1;
    L.o = mk3(0.0f, 0.0f, 0.0f); L.d = mk3(0.0f, 0.0f, 1.0f); L.inv = mk3(0.0f, 0.0f, 0.0f);
    int32_t slot = 0;
    if (valid) {
      if (kind == 0) {
        const float4 o = __ldcs(&q.ray_o[in_q][idx]), d = __ldcs(&q.ray_d[in_q][idx]);
        L.o = mk3(o); L.d = mk3(d);
      } else {
        const float4 o = __ldcs(&q.sh_o[idx]), d = __ldcs(&q.sh_d[idx]);
        L.o = mk3(o); L.d = mk3(d);
        L.t = nextafterf(o.w, INFINITY);  // see lane_load
        slot = __float_as_int(d.w);
      }
    }
    packet_traverse<BVH, ANALYTIC>(s, L, valid, stacks[threadIdx.x >> 5], overflow, n_nodes, n_tris, w_nodes, w_tris);
    if (valid) {
      if (kind == 0) __stcs(&q.hits[idx], make_float4(L.t, L.u, L.v, __int_as_float(L.tri)));
      else {
        const float4 lit = __ldcs(&q.sh_lit[idx]);
        float3 inc = make_float3(lit.x, lit.y, lit.z);
        if (L.tri == 0) { const float2 un = __ldcs(&q.sh_un[idx]); inc = make_float3(lit.w, un.x, un.y); }
        const float4 prev = q.accum[slot];
        q.accum[slot] = make_float4(prev.x + inc.x, prev.y + inc.y, prev.z + inc.z, 0.0f);
      }
    }
  }
  packet_counters(q, overflow, n_nodes, n_tris, w_nodes, w_tris);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_raygen (K3): primary rays of the chunk, CSMain compute:283-349.  A warp covers an 8x4 pixel tile of one sample.  The
// scene's root box is tested here: a ray that misses it would be culled at the first traversal step anyway (compute:245-246),
// so it takes the background now (:364-368) and never enters the queue.  sampleColor starts at 0 (:356).
// ---------------------------------------------------------------------------------------------------------------------
template <int BVH>
__global__ void __launch_bounds__(kStreamBlock) k_raygen(const FrameParams f, const SceneView s, const QueueView q, const ChunkView c) {
  const int lane = threadIdx.x & 31;
  const unsigned below = (1u << lane) - 1u;
  // root box: reference BVH = node 0's own box; LBVH = union of the root record's two (padded) child boxes
  f3 rmn = mk3(0.0f, 0.0f, 0.0f), rmx = rmn;
  bool have_box = false;
  if (BVH == RTB_BVH_REFERENCE) {
    if (s.n_nodes > 0) { rmn = mk3(__ldg(&s.nodes[0])); rmx = mk3(__ldg(&s.nodes[1])); have_box = true; }
  } else if (BVH == RTB_BVH_WIDE) {
    if (s.n_tris > 0 && s.root >= 0) {  // the root record's grid: p .. p + 256 cells (a superset of the scene's box)
      const float4 h = __ldg(&s.nodes[RTB_WIDE_F4 * (size_t)s.root]);
      const unsigned hdr = __float_as_uint(h.w);
      const float cx = __uint_as_float((hdr & 0xffu) << 23) * 0.0078125f, cy = __uint_as_float((hdr & 0xff00u) << 15) * 0.0078125f, cz = __uint_as_float((hdr & 0xff0000u) << 7) * 0.0078125f;  // 256 cells = S * 2^-7
      rmn = mk3(h.x - cx * 0.00390625f, h.y - cy * 0.00390625f, h.z - cz * 0.00390625f);  // one cell of slack on either side
      rmx = mk3(h.x + cx * 1.00390625f, h.y + cy * 1.00390625f, h.z + cz * 1.00390625f);
      have_box = true;
    }
  } else if (s.n_tris > 0 && s.root >= 0) {
#if RTB_LBVH_WIDTH == 4
    const float4* rec = s.nodes + 8 * (size_t)s.root;  // unused slots repeat slot 0's box, so the union may include them
    const float4 a0 = __ldg(rec), a1 = __ldg(rec + 1), a2 = __ldg(rec + 2), b0 = __ldg(rec + 3), b1 = __ldg(rec + 4), b2 = __ldg(rec + 5);
    rmn = mk3(fminf(fminf(a0.x, a0.y), fminf(a0.z, a0.w)), fminf(fminf(a1.x, a1.y), fminf(a1.z, a1.w)), fminf(fminf(a2.x, a2.y), fminf(a2.z, a2.w)));
    rmx = mk3(fmaxf(fmaxf(b0.x, b0.y), fmaxf(b0.z, b0.w)), fmaxf(fmaxf(b1.x, b1.y), fmaxf(b1.z, b1.w)), fmaxf(fmaxf(b2.x, b2.y), fmaxf(b2.z, b2.w)));
#else
    const float4 n0 = __ldg(&s.nodes[4 * s.root]), n1 = __ldg(&s.nodes[4 * s.root + 1]);
    const float4 n2 = __ldg(&s.nodes[4 * s.root + 2]), n3 = __ldg(&s.nodes[4 * s.root + 3]);
    rmn = mk3(fminf(n0.x, n2.x), fminf(n0.y, n2.y), fminf(n0.z, n2.z));
    rmx = mk3(fmaxf(n1.x, n3.x), fmaxf(n1.y, n3.y), fmaxf(n1.z, n3.z));
#endif
    have_box = true;
  }
  const bool empty = (BVH == RTB_BVH_REFERENCE) ? s.n_nodes == 0 : s.n_tris == 0;
  unsigned n_valid = 0;
  __shared__ BlockReserve<1> reserve;
  int it = 0;
  for (int32_t base = blockIdx.x * kStreamBlock; base < c.n_slots; base += gridDim.x * kStreamBlock, it++) {  // block-uniform trip count
    const int32_t slot = base + threadIdx.x;
    Ray ray;
    int px, py, sample;
    const bool valid = primary_ray_of_slot(f, c, slot, ray, px, py, sample);
    bool survives = valid && !empty;
    if (survives && have_box) {
      if (BVH == RTB_BVH_REFERENCE) survives = !(slab_entry(ray, rmn, rmx) >= RTB_INFINITY);
      else {
        const SlabRay sr = make_slab_ray(ray.o, ray.d);
        float entry;
        survives = slab_hit(sr, rmn, rmx, RTB_INFINITY, entry);
        if (BVH == RTB_BVH_WIDE && !survives) {  // the FMA form rounds by ~1e-7 |origin / d|: decide near misses with the exact form over the slack box
          Ray exact = ray;
          survives = !(slab_entry(exact, rmn, rmx) >= RTB_INFINITY);
        }
      }
    }
    if (valid) {
      n_valid++;
      const f3 bg = mk3(1.0f, 1.0f, 1.0f) * mk3(f.bg[0], f.bg[1], f.bg[2]);  // attenuation (1,1,1) * _BackgroundColor
      const f3 first = survives ? mk3(0.0f, 0.0f, 0.0f) : mk3(0.0f, 0.0f, 0.0f) + bg;
      q.accum[slot] = make_float4(first.x, first.y, first.z, 0.0f);
    }
    const unsigned m[1] = {__ballot_sync(kFull, survives)};
    int32_t* const counters[1] = {&RTB_CNT_RAY(q, 0)};
    int32_t first[1];
    block_reserve<1>(reserve, it, m, counters, first);
    if (survives) {
      const int32_t at = first[0] + __popc(m[0] & below);
      __stcs(&q.ray_o[0][at], make_float4(ray.o.x, ray.o.y, ray.o.z, __int_as_float(slot)));
      __stcs(&q.ray_d[0][at], make_float4(ray.d.x, ray.d.y, ray.d.z, 1.0f));
      __stcs(&q.ray_a[0][at], make_float2(1.0f, 1.0f));
    }
  }
  for (int o = 16; o > 0; o >>= 1) n_valid += __shfl_xor_sync(kFull, n_valid, o);
  if (lane == 0 && n_valid) atomicAdd(&q.totals[0], (unsigned long long)n_valid);
}

// ---------------------------------------------------------------------------------------------------------------------
// Shading of one closest hit, compute:370-473: what k_shade (wavefront) and k_tail (fused) have in common.
// ---------------------------------------------------------------------------------------------------------------------
struct Shaded {
  bool emit_shadow;            // :393 — a shadow query decides between `lit` and `unlit`
  f3 sh_origin, sh_dir;        // :395-398
  float sh_dist;               // :401
  f3 lit, unlit;               // the increment of sampleColor (:418) with / without the diffuse + specular terms
  bool emit_ray;               // :424 — the path continues
  f3 start, dir, att;          // :472 and the attenuation after :440/446/453
};

template <bool ANALYTIC>
__device__ __forceinline__ void shade_hit(const FrameParams& f, const SceneView& s, const Ray& ray, f3 att, const Hit& hit, int px, int py, int sample,
                                          int depth, Shaded& o) {
  f3 pos, nrm;
  hit_surface<ANALYTIC>(s, ray, hit, pos, nrm);  // :183-187
  const Material m = fetch_material(s, __float_as_int(__ldg(&s.tri_isect[RTB_TRI_F4 * hit.tri + 1]).w));
  f3 local = mk3(0.0f, 0.0f, 0.0f);
  if (f.en_ambient == 1) local = local + m.color * m.ka;  // :379
  f3 light_pos = mk3(f.light[0], f.light[1], f.light[2]);
  if (f.soft == 1) {  // :383-388
    const f3 j = random_unit_vector(mk3((float)px + (float)sample * 9.0f, ((float)py + (float)sample * 4.0f) + (float)depth, (float)sample)) * f.light_size;
    light_pos = light_pos + j;
  }
  const f3 to_light = light_pos - pos;
  o.sh_dir = hlsl_normalize(to_light);
  const float n_dot_l = fmaxf(0.0f, dot3(nrm, o.sh_dir));
  o.unlit = (att * local) * f.light_intensity;          // :418 when the shadow test fails or is not made
  o.lit = o.unlit;
  o.sh_origin = pos;
  o.sh_dist = 0.0f;
  o.emit_shadow = f.en_diffuse == 1 && n_dot_l > 0.0f;  // :393
  if (o.emit_shadow) {
    f3 lit_local = local + (m.color * m.kd) * n_dot_l;  // :408
    if (f.en_specular == 1 && m.ks > 0.0f) {            // :409-414
      const f3 view_dir = hlsl_normalize(negate(ray.d));
      const f3 half_vec = hlsl_normalize(o.sh_dir + view_dir);
      const float k = m.ks * pow32(fmaxf(dot3(nrm, half_vec), 0.0f));
      lit_local = lit_local + mk3(k, k, k);
    }
    o.lit = (att * lit_local) * f.light_intensity;
    o.sh_origin = pos + nrm * RTB_OFFSET;  // :396
    o.sh_dist = hlsl_length(to_light);      // :401
  }
  // continuation, :420-473
  o.emit_ray = false;
  o.start = pos; o.dir = mk3(0.0f, 0.0f, 0.0f); o.att = att;
  const bool should_reflect = m.ks > 0.0f;
  const bool should_refract = (f.en_refraction == 1 && m.kr > 0.0f);
  if ((should_reflect || should_refract) && depth + 1 < f.max_depth) {
    f3 next_dir;
    if (should_refract) {
      const f3 I = hlsl_normalize(ray.d);
      f3 N = nrm;
      float eta = 1.0f / m.ior;
      if (dot3(I, N) > 0.0f) { N = negate(N); eta = m.ior; }
      const float cosi = dot3(negate(I), N);
      const float k = 1.0f - (eta * eta) * (1.0f - cosi * cosi);
      if (k >= 0.0f) {
        next_dir = eta * I + (eta * cosi - sqrtf(k)) * N;
        o.att = att * (m.color * m.kr);
        o.start = pos + next_dir * RTB_OFFSET;
      } else {  // total internal reflection
        next_dir = hlsl_reflect(I, N);
        o.att = att * (m.color * m.ks);
        o.start = pos + N * RTB_OFFSET;
      }
    } else {
      next_dir = hlsl_reflect(hlsl_normalize(ray.d), nrm);
      o.att = att * (m.color * m.ks);
      o.start = pos + nrm * RTB_OFFSET;
    }
    if (f.glossy == 1 && f.roughness > 0.0f) {  // :459-470
      const f3 j = random_unit_vector(mk3(((float)px + (float)sample * 55.0f) + (float)depth, (float)py + (float)sample * 22.0f, (float)(depth * 13))) * f.roughness;
      next_dir = hlsl_normalize(next_dir + j);
    }
    o.dir = hlsl_normalize(next_dir);  // :472
    o.emit_ray = true;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_shade: everything the reference does with a closest-hit result at one depth, for queues of at least `tail_max` rays.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef RTB_SHADE_MIN_BLOCKS
#define RTB_SHADE_MIN_BLOCKS (1024 / RTB_STREAM_BLOCK)  /* 64 registers: measured best (profiles/r1e_sweep_shade*.log) */
#endif
template <bool ANALYTIC>
__global__ void __launch_bounds__(kStreamBlock, RTB_SHADE_MIN_BLOCKS) k_shade(const FrameParams f, const SceneView s, const QueueView q, const ChunkView c, const int depth,
                                                  const int32_t tail_max) {
  const int lane = threadIdx.x & 31;
  const int32_t n = RTB_CNT_RAY(q, depth);
  if (n < tail_max) return;  // k_tail finishes these paths
  const int in_q = depth & 1, out_q = in_q ^ 1;
  const unsigned below = (1u << lane) - 1u;
  unsigned n_hits = 0;
  __shared__ BlockReserve<2> reserve;
  int it = 0;

  for (int32_t base = blockIdx.x * kStreamBlock; base < n; base += gridDim.x * kStreamBlock, it++) {  // block-uniform trip count
    const int32_t idx = base + threadIdx.x;
    const bool active = idx < n;
    int32_t slot = idx;
    int px = 0, py = 0, sample = 0;
    Ray ray;
    ray.o = ray.d = ray.inv = mk3(0.0f, 0.0f, 0.0f);
    f3 att = mk3(1.0f, 1.0f, 1.0f), prev = mk3(0.0f, 0.0f, 0.0f);
    Hit hit; hit.t = 0.0f; hit.u = 0.0f; hit.v = 0.0f; hit.tri = -1;
    if (active) {
      const float4 o = __ldcs(&q.ray_o[in_q][idx]), d = __ldcs(&q.ray_d[in_q][idx]);
      const float2 a2 = __ldcs(&q.ray_a[in_q][idx]);
      const float4 a = make_float4(d.w, a2.x, a2.y, 0.0f);
      const float4 hrec = __ldcs(&q.hits[idx]);
      slot = __float_as_int(o.w);
      ray.o = mk3(o); ray.d = mk3(d);
      att = mk3(a);
      hit.t = hrec.x; hit.u = hrec.y; hit.v = hrec.z; hit.tri = __float_as_int(hrec.w);
      prev = mk3(q.accum[slot]);
      if (f.soft == 1 || f.glossy == 1) {  // the jitter hashes are seeded with the pixel and sample (:386,462)
        int local_row;
        slot_to_pixel(f, c, slot, px, local_row, sample);
        py = band_global_row(local_row, f.band_rank, f.band_world, f.band_rows);
      }
    }
    const bool found = active && hit.tri >= 0;
    Shaded o;
    o.emit_shadow = false; o.emit_ray = false;
    if (active && !found) {  // :364-368
      const f3 sum = prev + att * mk3(f.bg[0], f.bg[1], f.bg[2]);
      q.accum[slot] = make_float4(sum.x, sum.y, sum.z, 0.0f);
    }
    if (found) {
      if (depth == 0) n_hits++;
      shade_hit<ANALYTIC>(f, s, ray, att, hit, px, py, sample, depth, o);
      if (!o.emit_shadow) {  // otherwise k_traverse adds the lit or unlit increment once the shadow query is decided
        const f3 sum = prev + o.unlit;
        q.accum[slot] = make_float4(sum.x, sum.y, sum.z, 0.0f);
      }
    }
    // queue compaction: warp ballots, one atomicAdd per BLOCK and queue
    const unsigned m_sh = __ballot_sync(kFull, o.emit_shadow);
    const unsigned m_nx = __ballot_sync(kFull, o.emit_ray);
    const unsigned masks[2] = {m_sh, m_nx};
    int32_t* const counters[2] = {&RTB_CNT_SHADOW(q, depth), &RTB_CNT_RAY(q, depth + 1)};
    int32_t first[2];
    block_reserve<2>(reserve, it, masks, counters, first);
    const int32_t b_sh = first[0], b_nx = first[1];
    if (o.emit_shadow) {
      const int32_t at = b_sh + __popc(m_sh & below);
      __stcs(&q.sh_o[at], make_float4(o.sh_origin.x, o.sh_origin.y, o.sh_origin.z, o.sh_dist));
      __stcs(&q.sh_d[at], make_float4(o.sh_dir.x, o.sh_dir.y, o.sh_dir.z, __int_as_float(slot)));
      __stcs(&q.sh_lit[at], make_float4(o.lit.x, o.lit.y, o.lit.z, o.unlit.x));
      __stcs(&q.sh_un[at], make_float2(o.unlit.y, o.unlit.z));
    }
    if (o.emit_ray) {
      const int32_t at = b_nx + __popc(m_nx & below);
      __stcs(&q.ray_o[out_q][at], make_float4(o.start.x, o.start.y, o.start.z, __int_as_float(slot)));
      __stcs(&q.ray_d[out_q][at], make_float4(o.dir.x, o.dir.y, o.dir.z, o.att.x));
      __stcs(&q.ray_a[out_q][at], make_float2(o.att.y, o.att.z));
    }
  }

  for (int o = 16; o > 0; o >>= 1) n_hits += __shfl_xor_sync(kFull, n_hits, o);
  if (lane == 0 && n_hits) atomicAdd(&q.totals[3], (unsigned long long)n_hits);
}

// ---------------------------------------------------------------------------------------------------------------------
// k_tail: once a depth's ray queue has fewer than `tail_max` entries the remaining paths no longer fill the GPU, and one
// launch per depth would cost the latency of its slowest ray every time.  k_tail instead runs each remaining path to its
// end in one thread — shade, shadow query, next closest hit, shade, ... — so slow rays of different depths overlap.  Same
// per-slot operation order as the wavefront (and as the reference's depth loop), hence the same bits.
// ---------------------------------------------------------------------------------------------------------------------
template <int BVH, bool ANALYTIC>
__global__ void __launch_bounds__(kBlock) k_tail(const FrameParams f, const SceneView s, const QueueView q, const ChunkView c, const int depth0,
                                                 const int32_t tail_max) {
  const int32_t n = RTB_CNT_RAY(q, depth0);
  if (n >= tail_max) return;  // k_shade handled this depth
  const int in_q = depth0 & 1;
  unsigned n_hits = 0, n_cont = 0, n_shadow = 0, overflow = 0;
  for (int32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const float4 o4 = __ldcs(&q.ray_o[in_q][idx]), d4 = __ldcs(&q.ray_d[in_q][idx]);
    const float2 a2 = __ldcs(&q.ray_a[in_q][idx]);
    const float4 a4 = make_float4(d4.w, a2.x, a2.y, 0.0f);
    const float4 hrec = __ldcs(&q.hits[idx]);
    const int32_t slot = __float_as_int(o4.w);
    Ray ray = make_ray(mk3(o4), mk3(d4));
    f3 att = mk3(a4);
    Hit hit; hit.t = hrec.x; hit.u = hrec.y; hit.v = hrec.z; hit.tri = __float_as_int(hrec.w);
    f3 acc = mk3(q.accum[slot]);
    int px = 0, py = 0, sample = 0;
    if (f.soft == 1 || f.glossy == 1) {
      int local_row;
      slot_to_pixel(f, c, slot, px, local_row, sample);
      py = band_global_row(local_row, f.band_rank, f.band_world, f.band_rows);
    }
    for (int depth = depth0;; depth++) {
      if (hit.tri < 0) { acc = acc + att * mk3(f.bg[0], f.bg[1], f.bg[2]); break; }  // :364-368
      if (depth == 0) n_hits++;
      Shaded o;
      shade_hit<ANALYTIC>(f, s, ray, att, hit, px, py, sample, depth, o);
      if (o.emit_shadow) {
        n_shadow++;
        Ray sr; sr.o = o.sh_origin; sr.d = o.sh_dir; sr.inv = mk3(1.0f / o.sh_dir.x, 1.0f / o.sh_dir.y, 1.0f / o.sh_dir.z);  // :395-398
        Hit sh;
        const bool occluded = traverse<BVH, true, ANALYTIC>(s, sr, o.sh_dist, sh, overflow);
        acc = acc + (occluded ? o.unlit : o.lit);  // :406-418
      } else {
        acc = acc + o.unlit;
      }
      if (!o.emit_ray) break;
      n_cont++;
      ray = make_ray(o.start, o.dir);
      att = o.att;
      traverse<BVH, false, ANALYTIC>(s, ray, 0.0f, hit, overflow);
    }
    q.accum[slot] = make_float4(acc.x, acc.y, acc.z, 0.0f);
  }
  for (int o = 16; o > 0; o >>= 1) {
    n_hits += __shfl_xor_sync(kFull, n_hits, o);
    n_cont += __shfl_xor_sync(kFull, n_cont, o);
    n_shadow += __shfl_xor_sync(kFull, n_shadow, o);
    overflow += __shfl_xor_sync(kFull, overflow, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n_hits) atomicAdd(&q.totals[3], (unsigned long long)n_hits);
    if (n_cont) atomicAdd(&q.totals[1], (unsigned long long)n_cont);
    if (n_shadow) atomicAdd(&q.totals[2], (unsigned long long)n_shadow);
    if (overflow) atomicAdd(&q.totals[4], (unsigned long long)overflow);
  }
}

// Output row of a local row: the full frame (RTB_OUT_FRAME) or this rank's packed rows (RTB_OUT_COMPACT).
__device__ __forceinline__ size_t out_index(const FrameParams& f, int local_row, int px) {
  const int row = f.out_compact ? local_row : band_global_row(local_row, f.band_rank, f.band_world, f.band_rows);
  return (size_t)row * (size_t)f.width + (size_t)px;
}

__global__ void __launch_bounds__(kBlock) k_resolve(const FrameParams f, const QueueView q, const ChunkView c, uchar4* __restrict__ dst) {
  const int n_px = c.rows * f.width;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
    const int r = i / f.width, px = i - r * f.width;
    f3 sum = mk3(0.0f, 0.0f, 0.0f);
    for (int sidx = 0; sidx < f.spp; sidx++) sum = sum + mk3(__ldcs(&q.accum[pixel_to_slot(f, c, px, r, sidx)]));  // :475
    const float ns = (float)f.spp;
    const f3 fin = mk3(sum.x / ns, sum.y / ns, sum.z / ns);  // :478
    dst[out_index(f, c.row0 + r, px)] = make_uchar4((unsigned char)quantize_unorm8(fin.x, f.srgb), (unsigned char)quantize_unorm8(fin.y, f.srgb),
                                                     (unsigned char)quantize_unorm8(fin.z, f.srgb), 255);
  }
}

template <int BVH, bool ANALYTIC>
__global__ void __launch_bounds__(kBlock) k_debug(const FrameParams f, const SceneView s, const ChunkView c, uchar4* __restrict__ dst) {
  const int n_px = c.rows * f.width;
  unsigned overflow = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
    const int r = i / f.width, px = i - r * f.width;
    const int py = band_global_row(c.row0 + r, f.band_rank, f.band_world, f.band_rows);
    const Ray ray = generate_ray(f, px, py, -2);  // :486-489
    Hit h;
    const bool found = traverse<BVH, false, ANALYTIC>(s, ray, 0.0f, h, overflow);
    f3 fin;
    if (f.debug == 1) { const float g = h.t / 100.0f; fin = found ? mk3(g, g, g) : mk3(1.0f, 0.0f, 0.0f); }
    else if (f.debug == 2) {
      fin = mk3(0.0f, 0.0f, 1.0f);
      if (found) { f3 pos, nrm; hit_surface<ANALYTIC>(s, ray, h, pos, nrm); fin = nrm * 0.5f + mk3(0.5f, 0.5f, 0.5f); }
    }
    else fin = found ? mk3(0.0f, 1.0f, 0.0f) : mk3(0.2f, 0.2f, 0.2f);
    dst[out_index(f, c.row0 + r, px)] = make_uchar4((unsigned char)quantize_unorm8(fin.x, f.srgb), (unsigned char)quantize_unorm8(fin.y, f.srgb),
                                                     (unsigned char)quantize_unorm8(fin.z, f.srgb), 255);
  }
}

template <int BVH, bool ANALYTIC>
__global__ void __launch_bounds__(kBlock) k_aux(const FrameParams f, const SceneView s, int32_t* __restrict__ prim, float* __restrict__ t_out, int32_t* __restrict__ mat) {
  // 8x4 tiles keep a warp's rays coherent; i enumerates tile-major
  const int tiles_x = (f.width + 7) >> 3;
  const int n_px = tiles_x * ((f.height + 3) >> 2) * 32;
  unsigned overflow = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
    const int tile = i >> 5, lane = i & 31;
    const int px = (tile % tiles_x) * 8 + (lane & 7), py = (tile / tiles_x) * 4 + (lane >> 3);
    if (px >= f.width || py >= f.height) continue;
    const Ray ray = generate_ray(f, px, py, -1);
    Hit h;
    const bool found = traverse<BVH, false, ANALYTIC>(s, ray, 0.0f, h, overflow);
    const size_t at = (size_t)py * (size_t)f.width + (size_t)px;
    if (prim) prim[at] = found ? __float_as_int(__ldg(&s.tri_isect[RTB_TRI_F4 * h.tri]).w) : -1;
    if (t_out) t_out[at] = h.t;
    if (mat) mat[at] = found ? __float_as_int(__ldg(&s.tri_isect[RTB_TRI_F4 * h.tri + 1]).w) : -1;
  }
}

template <typename K>
int blocks_per_sm(K kernel) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kTravBlock, 0) != cudaSuccess || n < 1) n = 1;
  return n;
}

}  // namespace

int traverse_blocks_per_sm(int bvh) {
  if (bvh == RTB_BVH_WIDE) return blocks_per_sm(k_traverse_wide<false, false>);
  return bvh == RTB_BVH_REFERENCE ? blocks_per_sm(k_traverse_ref<false, false>) : blocks_per_sm(k_traverse_lbvh<false, false>);
}

void launch_raygen(int bvh, const FrameParams& f, const SceneView& s, const QueueView& q, const ChunkView& c, int grid, cudaStream_t st) {
  if (bvh == RTB_BVH_REFERENCE) k_raygen<RTB_BVH_REFERENCE><<<grid * kBlock / kStreamBlock, kStreamBlock, 0, st>>>(f, s, q, c);
  else if (bvh == RTB_BVH_WIDE) k_raygen<RTB_BVH_WIDE><<<grid * kBlock / kStreamBlock, kStreamBlock, 0, st>>>(f, s, q, c);
  else k_raygen<RTB_BVH_LBVH><<<grid * kBlock / kStreamBlock, kStreamBlock, 0, st>>>(f, s, q, c);
}

int node_record_f4(int bvh) { return bvh == RTB_BVH_REFERENCE ? 2 : (bvh == RTB_BVH_WIDE ? RTB_WIDE_F4 : lbvh_node_f4); }

size_t traverse_smem_bytes(int bvh, const SceneView& s) {
  return ((size_t)s.n_nodes * node_record_f4(bvh) + (size_t)s.n_tris * RTB_TRI_F4) * sizeof(float4);
}

cudaError_t traverse_enable_smem(int bvh, size_t bytes) {
  if (bvh == RTB_BVH_REFERENCE) return cudaFuncSetAttribute(k_traverse_ref<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (bvh == RTB_BVH_WIDE) return cudaFuncSetAttribute(k_traverse_wide<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return cudaFuncSetAttribute(k_traverse_lbvh<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Scenes with analytic primitives (s.n_prims > 0) run the ANALYTIC instantiations; everything else keeps the leaner
// triangle-only code.  The shared-memory variant exists for triangle-only scenes.
void launch_traverse(int bvh, const SceneView& s, const QueueView& q, int depth, int mode, int grid, size_t smem_bytes, cudaStream_t st) {
  const bool ref = bvh == RTB_BVH_REFERENCE;
  if (bvh == RTB_BVH_WIDE) {
    if (s.n_prims > 0) k_traverse_wide<false, true><<<grid, kTravBlock, 0, st>>>(s, q, depth, mode);
    else if (smem_bytes > 0) k_traverse_wide<true, false><<<grid, kBlockSmem, smem_bytes, st>>>(s, q, depth, mode);
    else k_traverse_wide<false, false><<<grid, kTravBlock, 0, st>>>(s, q, depth, mode);
    return;
  }
  if (s.n_prims > 0) {
    if (ref) k_traverse_ref<false, true><<<grid, kTravBlock, 0, st>>>(s, q, depth, mode);
    else k_traverse_lbvh<false, true><<<grid, kTravBlock, 0, st>>>(s, q, depth, mode);
  } else if (smem_bytes > 0) {  // small scene: one 1024-thread block per SM works out of its shared-memory copy
    if (ref) k_traverse_ref<true, false><<<grid, kBlockSmem, smem_bytes, st>>>(s, q, depth, mode);
    else k_traverse_lbvh<true, false><<<grid, kBlockSmem, smem_bytes, st>>>(s, q, depth, mode);
  } else {
    if (ref) k_traverse_ref<false, false><<<grid, kTravBlock, 0, st>>>(s, q, depth, mode);
    else k_traverse_lbvh<false, false><<<grid, kTravBlock, 0, st>>>(s, q, depth, mode);
  }
}

void launch_primary(int bvh, const FrameParams& f, const SceneView& s, const QueueView& q, const ChunkView& c, int grid, cudaStream_t st) {
  const bool ref = bvh == RTB_BVH_REFERENCE;
  if (s.n_prims > 0) {
    if (ref) k_primary<RTB_BVH_REFERENCE, true><<<grid, kStreamBlock, 0, st>>>(f, s, q, c);
    else k_primary<RTB_BVH_LBVH, true><<<grid, kStreamBlock, 0, st>>>(f, s, q, c);
  } else {
    if (ref) k_primary<RTB_BVH_REFERENCE, false><<<grid, kStreamBlock, 0, st>>>(f, s, q, c);
    else k_primary<RTB_BVH_LBVH, false><<<grid, kStreamBlock, 0, st>>>(f, s, q, c);
  }
}

void launch_packet(int bvh, const SceneView& s, const QueueView& q, int depth, int kind, int grid, cudaStream_t st) {
  const bool ref = bvh == RTB_BVH_REFERENCE;
  if (s.n_prims > 0) {
    if (ref) k_packet<RTB_BVH_REFERENCE, true><<<grid, kStreamBlock, 0, st>>>(s, q, depth, kind);
    else k_packet<RTB_BVH_LBVH, true><<<grid, kStreamBlock, 0, st>>>(s, q, depth, kind);
  } else {
    if (ref) k_packet<RTB_BVH_REFERENCE, false><<<grid, kStreamBlock, 0, st>>>(s, q, depth, kind);
    else k_packet<RTB_BVH_LBVH, false><<<grid, kStreamBlock, 0, st>>>(s, q, depth, kind);
  }
}

int stream_block_threads() { return kStreamBlock; }

// k_traverse_pool (binary LBVH records, scene in global memory): persistent grid, slot-interleaved stack scratch per warp.
int pool_blocks_per_sm() {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_traverse_pool<false>, kPoolBlock, 0) != cudaSuccess || n < 1) n = 1;
  return n;
}
size_t pool_scratch_bytes(int grid) { return (size_t)grid * (kPoolBlock / 32) * (size_t)(RTB_STACK_LBVH * RTB_POOL_SLOTS) * sizeof(float2); }
void launch_traverse_pool(const SceneView& s, const QueueView& q, int depth, int mode, int grid, void* scratch, cudaStream_t st) {
  if (s.n_prims > 0) k_traverse_pool<true><<<grid, kPoolBlock, 0, st>>>(s, q, depth, mode, (float2*)scratch);
  else k_traverse_pool<false><<<grid, kPoolBlock, 0, st>>>(s, q, depth, mode, (float2*)scratch);
}

void launch_shade(const FrameParams& f, const SceneView& s, const QueueView& q, const ChunkView& c, int depth, int32_t tail_max, int grid,
                  cudaStream_t st) {
  if (s.n_prims > 0) k_shade<true><<<grid * kBlock / kStreamBlock, kStreamBlock, 0, st>>>(f, s, q, c, depth, tail_max);
  else k_shade<false><<<grid * kBlock / kStreamBlock, kStreamBlock, 0, st>>>(f, s, q, c, depth, tail_max);
}

void launch_tail(int bvh, const FrameParams& f, const SceneView& s, const QueueView& q, const ChunkView& c, int depth, int32_t tail_max, int grid,
                 cudaStream_t st) {
  if (s.n_prims > 0) {
    if (bvh == RTB_BVH_REFERENCE) k_tail<RTB_BVH_REFERENCE, true><<<grid, kBlock, 0, st>>>(f, s, q, c, depth, tail_max);
    else if (bvh == RTB_BVH_WIDE) k_tail<RTB_BVH_WIDE, true><<<grid, kBlock, 0, st>>>(f, s, q, c, depth, tail_max);
    else k_tail<RTB_BVH_LBVH, true><<<grid, kBlock, 0, st>>>(f, s, q, c, depth, tail_max);
  } else {
    if (bvh == RTB_BVH_REFERENCE) k_tail<RTB_BVH_REFERENCE, false><<<grid, kBlock, 0, st>>>(f, s, q, c, depth, tail_max);
    else if (bvh == RTB_BVH_WIDE) k_tail<RTB_BVH_WIDE, false><<<grid, kBlock, 0, st>>>(f, s, q, c, depth, tail_max);
    else k_tail<RTB_BVH_LBVH, false><<<grid, kBlock, 0, st>>>(f, s, q, c, depth, tail_max);
  }
}

void launch_resolve(const FrameParams& f, const QueueView& q, const ChunkView& c, void* dst, int grid, cudaStream_t st) {
  k_resolve<<<grid, kBlock, 0, st>>>(f, q, c, (uchar4*)dst);
}

void launch_debug(int bvh, const FrameParams& f, const SceneView& s, const ChunkView& c, void* dst, int grid, cudaStream_t st) {
  if (s.n_prims > 0) {
    if (bvh == RTB_BVH_REFERENCE) k_debug<RTB_BVH_REFERENCE, true><<<grid, kBlock, 0, st>>>(f, s, c, (uchar4*)dst);
    else if (bvh == RTB_BVH_WIDE) k_debug<RTB_BVH_WIDE, true><<<grid, kBlock, 0, st>>>(f, s, c, (uchar4*)dst);
    else k_debug<RTB_BVH_LBVH, true><<<grid, kBlock, 0, st>>>(f, s, c, (uchar4*)dst);
  } else {
    if (bvh == RTB_BVH_REFERENCE) k_debug<RTB_BVH_REFERENCE, false><<<grid, kBlock, 0, st>>>(f, s, c, (uchar4*)dst);
    else if (bvh == RTB_BVH_WIDE) k_debug<RTB_BVH_WIDE, false><<<grid, kBlock, 0, st>>>(f, s, c, (uchar4*)dst);
    else k_debug<RTB_BVH_LBVH, false><<<grid, kBlock, 0, st>>>(f, s, c, (uchar4*)dst);
  }
}

void launch_aux(int bvh, const FrameParams& f, const SceneView& s, int32_t* prim, float* t, int32_t* mat, int grid, cudaStream_t st) {
  if (s.n_prims > 0) {
    if (bvh == RTB_BVH_REFERENCE) k_aux<RTB_BVH_REFERENCE, true><<<grid, kBlock, 0, st>>>(f, s, prim, t, mat);
    else if (bvh == RTB_BVH_WIDE) k_aux<RTB_BVH_WIDE, true><<<grid, kBlock, 0, st>>>(f, s, prim, t, mat);
    else k_aux<RTB_BVH_LBVH, true><<<grid, kBlock, 0, st>>>(f, s, prim, t, mat);
  } else {
    if (bvh == RTB_BVH_REFERENCE) k_aux<RTB_BVH_REFERENCE, false><<<grid, kBlock, 0, st>>>(f, s, prim, t, mat);
    else if (bvh == RTB_BVH_WIDE) k_aux<RTB_BVH_WIDE, false><<<grid, kBlock, 0, st>>>(f, s, prim, t, mat);
    else k_aux<RTB_BVH_LBVH, false><<<grid, kBlock, 0, st>>>(f, s, prim, t, mat);
  }
}

}  // namespace rtb
