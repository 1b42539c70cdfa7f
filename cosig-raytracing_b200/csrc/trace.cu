// trace.cu — the wavefront kernels of the per-pixel path (replaces CSMain, Assets/Shaders/BVHRayTracing.compute:273-511).
//
//   k_trace_shade<BVH, PRIMARY>   one depth of the reference's `for depth` loop (:360-473) for every live path:
//                                 [PRIMARY: ray generation :283-349] -> closest hit :362 -> miss/background :364-368 ->
//                                 shading :370-418 (the shadow test is deferred to k_shadow through the shadow queue) ->
//                                 continuation :420-473, compacted into the next depth's ray queue with one
//                                 ballot + one atomicAdd per warp.
//   k_shadow<BVH>                 the shadow query :395-406 as an any-hit traversal; adds the lit or unlit increment of
//                                 :418 to the slot's sampleColor.
//   k_resolve                     :475-478,510: average the spp slots of a pixel, saturate, UNORM8, store (possibly into
//                                 a peer GPU's frame: the NVLink gather is this store).
//   k_debug<BVH>, k_aux<BVH>      debug views :484-508 and the primary-hit maps used by the parity tests.
//
// All kernels are persistent: a fixed grid (a multiple of the SM count) whose warps claim 32 queue entries at a time with
// one atomicAdd, so queue sizes never travel to the host and a frame is a fixed launch sequence.
#include "kernels.hpp"
#include "trace.cuh"

namespace rtb {

namespace {

constexpr int kBlock = 256;

__device__ __forceinline__ int32_t warp_claim(int32_t* counter, int lane) {
  int32_t base = 0;
  if (lane == 0) base = atomicAdd(counter, 32);
  return __shfl_sync(0xffffffffu, base, 0);
}

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

template <int BVH, bool PRIMARY>
__global__ void __launch_bounds__(kBlock) k_trace_shade(const FrameParams f, const SceneView s, const QueueView q, const ChunkView c, const int depth) {
  const int lane = threadIdx.x & 31;
  const int32_t n = PRIMARY ? c.n_slots : RTB_CNT_RAY(q, depth);
  const int in_q = depth & 1, out_q = in_q ^ 1;
  unsigned n_rays = 0, n_hits = 0, overflow = 0;
  if (!PRIMARY && blockIdx.x == 0 && threadIdx.x == 0 && n > 0) atomicAdd(&q.totals[1], (unsigned long long)n);

  for (;;) {
    const int32_t base = warp_claim(&RTB_CNT_FETCH_RAY(q, depth), lane);
    if (base >= n) break;
    const int32_t idx = base + lane;
    bool active = idx < n;
    int32_t slot = idx;
    int px = 0, py = 0, sample = 0;
    Ray ray;
    f3 att = mk3(1.0f, 1.0f, 1.0f);
    if (PRIMARY) {
      int local_row;
      active = active && slot_to_pixel(f, c, slot, px, local_row, sample);
      py = band_global_row(local_row, f.band_rank, f.band_world, f.band_rows);
      if (active) ray = generate_ray(f, px, py, sample);
    } else if (active) {
      const float4 o = __ldcs(&q.ray_o[in_q][idx]), d = __ldcs(&q.ray_d[in_q][idx]), a = __ldcs(&q.ray_att[in_q][idx]);
      slot = __float_as_int(o.w);
      ray = make_ray(mk3(o), mk3(d));
      att = mk3(a);
      if (f.soft == 1 || f.glossy == 1) {  // the jitter hashes are seeded with the pixel and sample (:386,462)
        int local_row;
        slot_to_pixel(f, c, slot, px, local_row, sample);
        py = band_global_row(local_row, f.band_rank, f.band_world, f.band_rows);
      }
    }

    bool emit_shadow = false, emit_ray = false;
    float4 sh_o, sh_d, sh_lit, sh_unlit, nx_o, nx_d, nx_att;
    if (active) {
      n_rays++;
      Hit hit;
      const bool found = traverse<BVH, false>(s, ray, 0.0f, hit, overflow);
      f3 prev = mk3(0.0f, 0.0f, 0.0f);
      if (!PRIMARY) prev = mk3(q.accum[slot]);
      if (!found) {  // :364-368
        const f3 sum = prev + att * mk3(f.bg[0], f.bg[1], f.bg[2]);
        q.accum[slot] = make_float4(sum.x, sum.y, sum.z, 0.0f);
      } else {
        if (PRIMARY) n_hits++;
        const f3 pos = ray.o + hit.t * ray.d;  // :183
        const f3 nrm = hit_normal(s, hit);
        const Material m = fetch_material(s, __float_as_int(__ldg(&s.tri_isect[3 * hit.tri + 1]).w));
        f3 local = mk3(0.0f, 0.0f, 0.0f);
        if (f.en_ambient == 1) local = local + m.color * m.ka;  // :379
        f3 light_pos = mk3(f.light[0], f.light[1], f.light[2]);
        if (f.soft == 1) {  // :383-388
          const f3 j = random_unit_vector(mk3((float)px + (float)sample * 9.0f, ((float)py + (float)sample * 4.0f) + (float)depth, (float)sample)) * f.light_size;
          light_pos = light_pos + j;
        }
        const f3 to_light = light_pos - pos;
        const f3 light_dir = hlsl_normalize(to_light);
        const float n_dot_l = fmaxf(0.0f, dot3(nrm, light_dir));
        const f3 unlit = (att * local) * f.light_intensity;  // :418 when the shadow test fails or is not made
        if (f.en_diffuse == 1 && n_dot_l > 0.0f) {  // :393-416
          f3 lit_local = local + (m.color * m.kd) * n_dot_l;
          if (f.en_specular == 1 && m.ks > 0.0f) {
            const f3 view_dir = hlsl_normalize(negate(ray.d));
            const f3 half_vec = hlsl_normalize(light_dir + view_dir);
            const float k = m.ks * pow32(fmaxf(dot3(nrm, half_vec), 0.0f));
            lit_local = lit_local + mk3(k, k, k);
          }
          const f3 lit = (att * lit_local) * f.light_intensity;
          const f3 so = pos + nrm * RTB_OFFSET;
          emit_shadow = true;
          sh_o = make_float4(so.x, so.y, so.z, hlsl_length(to_light));
          sh_d = make_float4(light_dir.x, light_dir.y, light_dir.z, __int_as_float(slot));
          sh_lit = make_float4(lit.x, lit.y, lit.z, 0.0f);
          sh_unlit = make_float4(unlit.x, unlit.y, unlit.z, 0.0f);
          if (PRIMARY) q.accum[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        } else {
          const f3 sum = prev + unlit;
          q.accum[slot] = make_float4(sum.x, sum.y, sum.z, 0.0f);
        }
        // continuation, :420-473
        const bool should_reflect = m.ks > 0.0f;
        const bool should_refract = (f.en_refraction == 1 && m.kr > 0.0f);
        if ((should_reflect || should_refract) && depth + 1 < f.max_depth) {
          f3 next_dir, start = pos;
          if (should_refract) {
            const f3 I = hlsl_normalize(ray.d);
            f3 N = nrm;
            float eta = 1.0f / m.ior;
            if (dot3(I, N) > 0.0f) { N = negate(N); eta = m.ior; }
            const float cosi = dot3(negate(I), N);
            const float k = 1.0f - (eta * eta) * (1.0f - cosi * cosi);
            if (k >= 0.0f) {
              next_dir = eta * I + (eta * cosi - sqrtf(k)) * N;
              att = att * (m.color * m.kr);
              start = start + next_dir * RTB_OFFSET;
            } else {  // total internal reflection
              next_dir = hlsl_reflect(I, N);
              att = att * (m.color * m.ks);
              start = start + N * RTB_OFFSET;
            }
          } else {
            next_dir = hlsl_reflect(hlsl_normalize(ray.d), nrm);
            att = att * (m.color * m.ks);
            start = start + nrm * RTB_OFFSET;
          }
          if (f.glossy == 1 && f.roughness > 0.0f) {  // :459-470
            const f3 j = random_unit_vector(mk3(((float)px + (float)sample * 55.0f) + (float)depth, (float)py + (float)sample * 22.0f, (float)(depth * 13))) * f.roughness;
            next_dir = hlsl_normalize(next_dir + j);
          }
          const f3 nd = hlsl_normalize(next_dir);  // :472
          emit_ray = true;
          nx_o = make_float4(start.x, start.y, start.z, __int_as_float(slot));
          nx_d = make_float4(nd.x, nd.y, nd.z, 0.0f);
          nx_att = make_float4(att.x, att.y, att.z, 0.0f);
        }
      }
    }

    // queue compaction: one atomicAdd per warp and queue
    const unsigned m_sh = __ballot_sync(0xffffffffu, emit_shadow);
    const unsigned m_nx = __ballot_sync(0xffffffffu, emit_ray);
    int32_t b_sh = 0, b_nx = 0;
    if (lane == 0) {
      if (m_sh) b_sh = atomicAdd(&RTB_CNT_SHADOW(q, depth), __popc(m_sh));
      if (m_nx) b_nx = atomicAdd(&RTB_CNT_RAY(q, depth + 1), __popc(m_nx));
    }
    b_sh = __shfl_sync(0xffffffffu, b_sh, 0);
    b_nx = __shfl_sync(0xffffffffu, b_nx, 0);
    const unsigned below = (1u << lane) - 1u;
    if (emit_shadow) {
      const int32_t at = b_sh + __popc(m_sh & below);
      __stcs(&q.sh_o[at], sh_o); __stcs(&q.sh_d[at], sh_d); __stcs(&q.sh_lit[at], sh_lit); __stcs(&q.sh_unlit[at], sh_unlit);
    }
    if (emit_ray) {
      const int32_t at = b_nx + __popc(m_nx & below);
      __stcs(&q.ray_o[out_q][at], nx_o); __stcs(&q.ray_d[out_q][at], nx_d); __stcs(&q.ray_att[out_q][at], nx_att);
    }
  }

  // per-warp totals
  for (int o = 16; o > 0; o >>= 1) {
    n_rays += __shfl_xor_sync(0xffffffffu, n_rays, o);
    n_hits += __shfl_xor_sync(0xffffffffu, n_hits, o);
    overflow += __shfl_xor_sync(0xffffffffu, overflow, o);
  }
  if (lane == 0) {
    if (PRIMARY && n_rays) atomicAdd(&q.totals[0], (unsigned long long)n_rays);
    if (PRIMARY && n_hits) atomicAdd(&q.totals[3], (unsigned long long)n_hits);
    if (overflow) atomicAdd(&q.totals[4], (unsigned long long)overflow);
  }
}

template <int BVH>
__global__ void __launch_bounds__(kBlock) k_shadow(const FrameParams f, const SceneView s, const QueueView q, const int depth) {
  const int lane = threadIdx.x & 31;
  const int32_t n = RTB_CNT_SHADOW(q, depth);
  unsigned overflow = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) atomicAdd(&q.totals[2], (unsigned long long)n);
  for (;;) {
    const int32_t base = warp_claim(&RTB_CNT_FETCH_SHADOW(q, depth), lane);
    if (base >= n) break;
    const int32_t idx = base + lane;
    if (idx < n) {
      const float4 o = __ldcs(&q.sh_o[idx]), d = __ldcs(&q.sh_d[idx]);
      Ray ray; ray.o = mk3(o); ray.d = mk3(d); ray.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);  // :395-398
      Hit hit;
      const bool shadowed = traverse<BVH, true>(s, ray, o.w, hit, overflow);  // lit <=> !hit || t > distToLight, :406
      const float4 inc = shadowed ? __ldcs(&q.sh_unlit[idx]) : __ldcs(&q.sh_lit[idx]);
      const int32_t slot = __float_as_int(d.w);
      const float4 prev = q.accum[slot];
      q.accum[slot] = make_float4(prev.x + inc.x, prev.y + inc.y, prev.z + inc.z, 0.0f);
    }
  }
  for (int o = 16; o > 0; o >>= 1) overflow += __shfl_xor_sync(0xffffffffu, overflow, o);
  if (lane == 0 && overflow) atomicAdd(&q.totals[4], (unsigned long long)overflow);
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

template <int BVH>
__global__ void __launch_bounds__(kBlock) k_debug(const FrameParams f, const SceneView s, const ChunkView c, uchar4* __restrict__ dst) {
  const int n_px = c.rows * f.width;
  unsigned overflow = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
    const int r = i / f.width, px = i - r * f.width;
    const int py = band_global_row(c.row0 + r, f.band_rank, f.band_world, f.band_rows);
    const Ray ray = generate_ray(f, px, py, -2);  // :486-489
    Hit h;
    const bool found = traverse<BVH, false>(s, ray, 0.0f, h, overflow);
    f3 fin;
    if (f.debug == 1) { const float g = h.t / 100.0f; fin = found ? mk3(g, g, g) : mk3(1.0f, 0.0f, 0.0f); }
    else if (f.debug == 2) fin = found ? hit_normal(s, h) * 0.5f + mk3(0.5f, 0.5f, 0.5f) : mk3(0.0f, 0.0f, 1.0f);
    else fin = found ? mk3(0.0f, 1.0f, 0.0f) : mk3(0.2f, 0.2f, 0.2f);
    dst[out_index(f, c.row0 + r, px)] = make_uchar4((unsigned char)quantize_unorm8(fin.x, f.srgb), (unsigned char)quantize_unorm8(fin.y, f.srgb),
                                                     (unsigned char)quantize_unorm8(fin.z, f.srgb), 255);
  }
}

template <int BVH>
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
    const bool found = traverse<BVH, false>(s, ray, 0.0f, h, overflow);
    const size_t at = (size_t)py * (size_t)f.width + (size_t)px;
    if (prim) prim[at] = found ? __float_as_int(__ldg(&s.tri_isect[3 * h.tri]).w) : -1;
    if (t_out) t_out[at] = h.t;
    if (mat) mat[at] = found ? __float_as_int(__ldg(&s.tri_isect[3 * h.tri + 1]).w) : -1;
  }
}

template <typename K>
int blocks_per_sm(K kernel) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kBlock, 0) != cudaSuccess || n < 1) n = 1;
  return n;
}

}  // namespace

int trace_blocks_per_sm(int bvh, bool primary) {
  if (bvh == RTB_BVH_REFERENCE) return primary ? blocks_per_sm(k_trace_shade<RTB_BVH_REFERENCE, true>) : blocks_per_sm(k_trace_shade<RTB_BVH_REFERENCE, false>);
  return primary ? blocks_per_sm(k_trace_shade<RTB_BVH_LBVH, true>) : blocks_per_sm(k_trace_shade<RTB_BVH_LBVH, false>);
}
int shadow_blocks_per_sm(int bvh) {
  return bvh == RTB_BVH_REFERENCE ? blocks_per_sm(k_shadow<RTB_BVH_REFERENCE>) : blocks_per_sm(k_shadow<RTB_BVH_LBVH>);
}

void launch_trace_shade(int bvh, bool primary, const FrameParams& f, const SceneView& s, const QueueView& q, const ChunkView& c, int depth,
                        int grid, cudaStream_t st) {
  if (bvh == RTB_BVH_REFERENCE) {
    if (primary) k_trace_shade<RTB_BVH_REFERENCE, true><<<grid, kBlock, 0, st>>>(f, s, q, c, depth);
    else k_trace_shade<RTB_BVH_REFERENCE, false><<<grid, kBlock, 0, st>>>(f, s, q, c, depth);
  } else {
    if (primary) k_trace_shade<RTB_BVH_LBVH, true><<<grid, kBlock, 0, st>>>(f, s, q, c, depth);
    else k_trace_shade<RTB_BVH_LBVH, false><<<grid, kBlock, 0, st>>>(f, s, q, c, depth);
  }
}

void launch_shadow(int bvh, const FrameParams& f, const SceneView& s, const QueueView& q, int depth, int grid, cudaStream_t st) {
  if (bvh == RTB_BVH_REFERENCE) k_shadow<RTB_BVH_REFERENCE><<<grid, kBlock, 0, st>>>(f, s, q, depth);
  else k_shadow<RTB_BVH_LBVH><<<grid, kBlock, 0, st>>>(f, s, q, depth);
}

void launch_resolve(const FrameParams& f, const QueueView& q, const ChunkView& c, void* dst, int grid, cudaStream_t st) {
  k_resolve<<<grid, kBlock, 0, st>>>(f, q, c, (uchar4*)dst);
}

void launch_debug(int bvh, const FrameParams& f, const SceneView& s, const ChunkView& c, void* dst, int grid, cudaStream_t st) {
  if (bvh == RTB_BVH_REFERENCE) k_debug<RTB_BVH_REFERENCE><<<grid, kBlock, 0, st>>>(f, s, c, (uchar4*)dst);
  else k_debug<RTB_BVH_LBVH><<<grid, kBlock, 0, st>>>(f, s, c, (uchar4*)dst);
}

void launch_aux(int bvh, const FrameParams& f, const SceneView& s, int32_t* prim, float* t, int32_t* mat, int grid, cudaStream_t st) {
  if (bvh == RTB_BVH_REFERENCE) k_aux<RTB_BVH_REFERENCE><<<grid, kBlock, 0, st>>>(f, s, prim, t, mat);
  else k_aux<RTB_BVH_LBVH><<<grid, kBlock, 0, st>>>(f, s, prim, t, mat);
}

}  // namespace rtb
