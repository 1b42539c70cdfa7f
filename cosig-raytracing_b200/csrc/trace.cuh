// trace.cuh — device functions of the per-pixel path: ray generation, BVH traversal (reference-shape and LBVH),
// Möller–Trumbore, shading and continuation.  Every function cites the lines of Assets/Shaders/BVHRayTracing.compute it
// restates; operation order follows that file token by token because the arithmetic spec forbids re-association.
#pragma once
#include "rtb_device.cuh"

namespace rtb {

struct Ray { f3 o, d, inv; };
// CreateRay, compute:137-144
__device__ __forceinline__ Ray make_ray(f3 o, f3 d) {
  Ray r; r.o = o; r.d = d; r.inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); return r;
}

struct Hit { float t, u, v; int32_t tri; };  // tri: leaf-order triangle index, -1 = miss

// ---------------------------------------------------------------------------------------------------------------------
// Hashes (compute:108-131).  cos/sin of RandomUnitVector use a fixed FP32 polynomial so that CPU checker and GPU agree
// bit for bit; everything else is + - * floor sqrt.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void hash22(float px, float py, float& ox, float& oy) {
  float a = hlsl_frac(px * .1031f), b = hlsl_frac(py * .1030f), c = hlsl_frac(px * .0973f);
  const float d = dot3(mk3(a, b, c), mk3(b + 33.33f, c + 33.33f, a + 33.33f));
  a = a + d; b = b + d; c = c + d;
  ox = hlsl_frac((a + b) * c);
  oy = hlsl_frac((a + c) * b);
}
__device__ __forceinline__ f3 hash33(f3 p) {
  p = mk3(hlsl_frac(p.x * .1031f), hlsl_frac(p.y * .1030f), hlsl_frac(p.z * .0973f));
  const float d = dot3(p, mk3(p.y + 33.33f, p.x + 33.33f, p.z + 33.33f));
  p = mk3(p.x + d, p.y + d, p.z + d);
  return mk3(hlsl_frac((p.x + p.y) * p.z), hlsl_frac((p.x + p.x) * p.y), hlsl_frac((p.y + p.x) * p.x));
}
__device__ __forceinline__ void det_sincos(float a, float& s_out, float& c_out) {
  const int k = (int)floorf(a * 0.63661975f);
  const float r = a - (float)k * 1.5707964f;
  const float r2 = r * r;
  const float s = r * (1.0f + r2 * (-1.6666667e-1f + r2 * (8.3333338e-3f + r2 * (-1.9841270e-4f + r2 * (2.7557319e-6f + r2 * -2.5052108e-8f)))));
  const float c = 1.0f + r2 * (-0.5f + r2 * (4.1666668e-2f + r2 * (-1.3888889e-3f + r2 * (2.4801587e-5f + r2 * (-2.7557319e-7f + r2 * 2.0876757e-9f)))));
  switch (k & 3) {
    case 0: s_out = s; c_out = c; break;
    case 1: s_out = c; c_out = -s; break;
    case 2: s_out = -s; c_out = -c; break;
    default: s_out = -c; c_out = s; break;
  }
}
__device__ __forceinline__ f3 random_unit_vector(f3 seed) {
  const f3 h = hash33(seed);
  const float z = h.z * 2.0f - 1.0f;
  const float a = h.x * 6.2831853f;
  const float r = sqrtf(1.0f - z * z);
  float s, c;
  det_sincos(a, s, c);
  return mk3(r * c, r * s, z);
}

// ---------------------------------------------------------------------------------------------------------------------
// Ray generation, CSMain compute:283-349.  sample >= 0: AA sample (jitter + motion blur apply); sample == -1: pixel-centre
// ray under the current projection (rtb_render_aux); sample == -2: the always-perspective centre ray of the debug views
// (:486-489).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ Ray generate_ray(const FrameParams& f, int px, int py, int sample) {
  const float width = (float)f.width, height = (float)f.height;
  const float aspect = width / height;
  const float half_height = f.cam_dist * f.tan_half;
  const float plane_height = 2.0f * half_height;
  const float plane_width = plane_height * aspect;
  float ox = 0.5f, oy = 0.5f;
  if (sample >= 0 && f.spp > 1) {
    const int gy = sample / f.grid_w, gx = sample % f.grid_w;
    float jx, jy;
    hash22((float)px + (float)sample * 13.0f, (float)py + (float)sample * 7.0f, jx, jy);
    ox = ((float)gx + jx) / (float)f.grid_w;
    oy = ((float)gy + jy) / (float)f.grid_h;
  }
  f3 oc, dc;
  if (f.ortho == 1 && sample >= -1) {
    const float ohh = f.ortho_size, ohw = ohh * aspect;
    const float ou = ((((float)px + ox) / width - 0.5f) * 2.0f) * ohw;
    const float ov = ((((float)py + oy) / height - 0.5f) * 2.0f) * ohh;
    oc = mk3(ou, ov, f.cam_dist);
    dc = mk3(0.0f, 0.0f, -1.0f);
  } else {
    const float u = (((float)px + ox) / width - 0.5f) * plane_width;
    const float v = (((float)py + oy) / height - 0.5f) * plane_height;
    oc = mk3(0.0f, 0.0f, f.cam_dist);
    dc = hlsl_normalize(mk3(u, v, 0.0f) - oc);
  }
  const float* M = f.cam;
  f3 o = mk3(((M[0] * oc.x + M[1] * oc.y) + M[2] * oc.z) + M[3] * 1.0f,
             ((M[4] * oc.x + M[5] * oc.y) + M[6] * oc.z) + M[7] * 1.0f,
             ((M[8] * oc.x + M[9] * oc.y) + M[10] * oc.z) + M[11] * 1.0f);
  const f3 d = hlsl_normalize(mk3((M[0] * dc.x + M[1] * dc.y) + M[2] * dc.z, (M[4] * dc.x + M[5] * dc.y) + M[6] * dc.z,
                                  (M[8] * dc.x + M[9] * dc.y) + M[10] * dc.z));
  if (f.blur == 1 && sample >= 0) {  // :342-349
    const f3 r = random_unit_vector(mk3((float)px + (float)sample, (float)py, (float)sample));
    const f3 shake = ((r - mk3(0.5f, 0.5f, 0.5f)) * 0.2f) * f.shutter;
    o = o + shake;
  }
  return make_ray(o, d);
}

// ---------------------------------------------------------------------------------------------------------------------
// Intersection primitives
// ---------------------------------------------------------------------------------------------------------------------
// IntersectAABB, compute:199-216
__device__ __forceinline__ float slab_entry(const Ray& r, f3 mn, f3 mx) {
  const f3 t0 = (mn - r.o) * r.inv, t1 = (mx - r.o) * r.inv;
  const float dst_a = fmaxf(fmaxf(fminf(t0.x, t1.x), fminf(t0.y, t1.y)), fminf(t0.z, t1.z));
  const float dst_b = fminf(fminf(fmaxf(t0.x, t1.x), fmaxf(t0.y, t1.y)), fmaxf(t0.z, t1.z));
  if (dst_a > dst_b || dst_b < 0.0f) return RTB_INFINITY;
  return dst_a;
}

// Slab test of the LBVH traversal: same semantics as slab_entry, evaluated as fma(bound, inv, -origin*inv).  The rounding
// differs from IntersectAABB's by a few ulp of |origin*inv|; LBVH boxes are padded outward at build time by more than
// that (lbvh.cu: k_emit), so every box the exact test would enter is still entered.  It never decides t, u, v or ids.
__device__ __forceinline__ float slab_entry_fma(f3 inv, f3 ood, f3 mn, f3 mx) {
  const float t0x = __fmaf_rn(mn.x, inv.x, -ood.x), t1x = __fmaf_rn(mx.x, inv.x, -ood.x);
  const float t0y = __fmaf_rn(mn.y, inv.y, -ood.y), t1y = __fmaf_rn(mx.y, inv.y, -ood.y);
  const float t0z = __fmaf_rn(mn.z, inv.z, -ood.z), t1z = __fmaf_rn(mx.z, inv.z, -ood.z);
  const float dst_a = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fminf(t0z, t1z));
  const float dst_b = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
  if (dst_a > dst_b || dst_b < 0.0f) return RTB_INFINITY;
  return dst_a;
}
// The form the persistent LBVH traversal uses: entry = max(slab entries, 0), exit = min(slab exits, bound); the box is opened
// iff entry <= exit.  Compared with "skip when entry >= bound" (compute:246) this also opens a box whose entry equals the
// bound exactly — a superset, so no hit can be lost, and a triangle there cannot pass the strict "t < bound" test.
__device__ __forceinline__ bool slab_hit_fma(f3 inv, f3 ood, f3 mn, f3 mx, float bound, float& entry) {
  const float t0x = __fmaf_rn(mn.x, inv.x, -ood.x), t1x = __fmaf_rn(mx.x, inv.x, -ood.x);
  const float t0y = __fmaf_rn(mn.y, inv.y, -ood.y), t1y = __fmaf_rn(mx.y, inv.y, -ood.y);
  const float t0z = __fmaf_rn(mn.z, inv.z, -ood.z), t1z = __fmaf_rn(mx.z, inv.z, -ood.z);
  entry = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
  const float exit = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), bound));
  return entry <= exit;
}

// Reciprocal direction for the FMA slab test: finite even for axis-parallel rays (a +-inf reciprocal would turn
// fma(bound, inf, -origin*inf) into inf - inf = NaN), so such rays see slabs at +-1e18 * (bound - origin).
__device__ __forceinline__ float safe_rcp(float d) { return fabsf(d) > 1e-18f ? 1.0f / d : copysignf(1e18f, d); }
__device__ __forceinline__ f3 safe_inverse(f3 d) { return mk3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z)); }

// The ray constants of the LBVH box test.  fma(bound, 1/d, -o/d) differs from the reference's (bound - o) * (1/d) by the rounding
// of o/d — about 2^-24 |o| in space — and the triangle test itself places hits with an uncertainty of that order (o - v0 is
// rounded), so whether a hit survives the box tests must not depend on it.  The build pads boxes for the scene's own scale
// (lbvh.cu: k_emit); what grows with the CAMERA's distance is handled here, per ray and at no cost per node: the min planes and
// the max planes get their own o/d, shifted by e = 2^-21 max|o| |1/d| towards "enters earlier, leaves later" (for d > 0 the min
// plane is the near one; for d < 0 the max plane).
struct SlabRay { f3 inv, ood_mn, ood_mx; };
__device__ __forceinline__ SlabRay make_slab_ray(f3 o, f3 d) {
  SlabRay r;
  r.inv = safe_inverse(d);
  const f3 ood = o * r.inv;
  const float reach = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z)) * 4.76837158203125e-7f;  // 2^-21 max|o|
  const f3 e = mk3(reach * r.inv.x, reach * r.inv.y, reach * r.inv.z);  // signed like 1/d
  r.ood_mn = ood + e;
  r.ood_mx = ood - e;
  return r;
}
__device__ __forceinline__ bool slab_hit(const SlabRay& r, f3 mn, f3 mx, float bound, float& entry) {
  const float t0x = __fmaf_rn(mn.x, r.inv.x, -r.ood_mn.x), t1x = __fmaf_rn(mx.x, r.inv.x, -r.ood_mx.x);
  const float t0y = __fmaf_rn(mn.y, r.inv.y, -r.ood_mn.y), t1y = __fmaf_rn(mx.y, r.inv.y, -r.ood_mx.y);
  const float t0z = __fmaf_rn(mn.z, r.inv.z, -r.ood_mn.z), t1z = __fmaf_rn(mx.z, r.inv.z, -r.ood_mx.z);
  entry = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
  const float exit = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), bound));
  return entry <= exit;
}


// ---------------------------------------------------------------------------------------------------------------------
// Analytic primitives (RTB_PRIM_ANALYTIC): the reference's SphereInstance / BoxInstance, Assets/Services/BVH/HittableObjects.cs
// (never instantiated there; semantics taken from it): the ray goes to object space with its direction re-normalised
// (:49-54, :153-155), meets the unit sphere (quadratic, :82-107) or the unit cube (slabs with face tracking, :180-223),
// the hit point returns to world space and t = |pWS - origin| (:60-62).  P = 6 float4: objectToWorld rows, worldToObject rows.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ f3 mul_point_rows(const float4 r0, const float4 r1, const float4 r2, f3 v) {  // MultiplyPoint3x4
  return mk3(((r0.x * v.x + r0.y * v.y) + r0.z * v.z) + r0.w, ((r1.x * v.x + r1.y * v.y) + r1.z * v.z) + r1.w,
             ((r2.x * v.x + r2.y * v.y) + r2.z * v.z) + r2.w);
}
__device__ __forceinline__ f3 mul_vector_rows(const float4 r0, const float4 r1, const float4 r2, f3 v) {  // MultiplyVector
  return mk3((r0.x * v.x + r0.y * v.y) + r0.z * v.z, (r1.x * v.x + r1.y * v.y) + r1.z * v.z, (r2.x * v.x + r2.y * v.y) + r2.z * v.z);
}
__device__ __forceinline__ bool intersect_unit_sphere(f3 o, f3 d, float& t) {
  const float a = dot3(d, d);
  const float b = 2.0f * dot3(o, d);
  const float c = dot3(o, o) - 1.0f;
  const float disc = b * b - (4.0f * a) * c;
  if (disc < 0.0f) return false;
  const float s = sqrtf(disc);
  const float t0 = (-b - s) / (2.0f * a), t1 = (-b + s) / (2.0f * a);
  t = (t0 > 1e-3f) ? t0 : t1;
  return t > 1e-3f;
}
// face: 0 none, 1 -x, 2 +x, 3 -y, 4 +y, 5 -z, 6 +z
__device__ __forceinline__ bool intersect_unit_box(f3 o, f3 d, float& t, int& face) {
  float tmin = -1e20f, tmax = 1e20f;
  int nmin = 0, nmax = 0;
  const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
  for (int axis = 0; axis < 3; axis++) {
    const float inv_d = fabsf(dd[axis]) > 1e-8f ? 1.0f / dd[axis] : INFINITY;
    float t1 = (-0.5f - oo[axis]) * inv_d, t2 = (0.5f - oo[axis]) * inv_d;
    int n1 = 1 + 2 * axis, n2 = 2 + 2 * axis;
    if (t1 > t2) { const float tt = t1; t1 = t2; t2 = tt; const int nn = n1; n1 = n2; n2 = nn; }
    if (t1 > tmin) { tmin = t1; nmin = n1; }
    if (t2 < tmax) { tmax = t2; nmax = n2; }
    if (tmin > tmax) return false;
    if (tmax < 1e-3f) return false;
  }
  t = tmin >= 1e-3f ? tmin : tmax;
  face = (t == tmin) ? nmin : nmax;
  return t >= 1e-3f;
}
// Hit test of one analytic primitive.  Returns true with world t, object-space t and the box face when it is hit with
// t_world > Epsilon; the caller applies its upper bound.
__device__ __forceinline__ bool intersect_analytic(const float4* __restrict__ P, int kind, f3 ro, f3 rd, float& t_world, float& t_os, int& face) {
  const float4 w0 = __ldg(&P[3]), w1 = __ldg(&P[4]), w2 = __ldg(&P[5]);
  const f3 o = mul_point_rows(w0, w1, w2, ro);
  const f3 d = unity_normalized(mul_vector_rows(w0, w1, w2, rd));
  face = 0;
  if (kind == 1 ? !intersect_unit_sphere(o, d, t_os) : !intersect_unit_box(o, d, t_os, face)) return false;
  const f3 p_os = o + t_os * d;
  const f3 p_ws = mul_point_rows(__ldg(&P[0]), __ldg(&P[1]), __ldg(&P[2]), p_os);
  const f3 dv = p_ws - ro;
  t_world = sqrtf((dv.x * dv.x + dv.y * dv.y) + dv.z * dv.z);
  return t_world > RTB_EPSILON;
}
// World position and normal of an analytic hit: rec.positionWS = pWS, rec.normalWS = (worldToObject^T nOS).normalized (:65-71, :166-170)
__device__ __forceinline__ void analytic_surface(const float4* __restrict__ P, int kind, f3 ro, f3 rd, float t_os, int face, f3& pos, f3& nrm) {
  const float4 w0 = __ldg(&P[3]), w1 = __ldg(&P[4]), w2 = __ldg(&P[5]);
  const f3 o = mul_point_rows(w0, w1, w2, ro);
  const f3 d = unity_normalized(mul_vector_rows(w0, w1, w2, rd));
  const f3 p_os = o + t_os * d;
  pos = mul_point_rows(__ldg(&P[0]), __ldg(&P[1]), __ldg(&P[2]), p_os);
  f3 n_os;
  if (kind == 1) n_os = unity_normalized(p_os);
  else n_os = mk3(face == 1 ? -1.0f : (face == 2 ? 1.0f : 0.0f), face == 3 ? -1.0f : (face == 4 ? 1.0f : 0.0f), face == 5 ? -1.0f : (face == 6 ? 1.0f : 0.0f));
  nrm = unity_normalized(mk3((w0.x * n_os.x + w1.x * n_os.y) + w2.x * n_os.z, (w0.y * n_os.x + w1.y * n_os.y) + w2.y * n_os.z,
                             (w0.z * n_os.x + w1.z * n_os.y) + w2.z * n_os.z));
}

// IntersectTriangle, compute:153-190, on the stored (v0, e1 = v1-v0, e2 = v2-v0).  Returns true and t,u,v when the
// triangle is hit with t > Epsilon; the caller applies its own upper bound (closest: t < best.t; shadow: t <= dist).
__device__ __forceinline__ bool moller_trumbore(const Ray& r, f3 v0, f3 e1, f3 e2, float& t, float& u, float& v) {
  const f3 pvec = cross3(r.d, e2);
  const float det = dot3(e1, pvec);
  if (fabsf(det) < RTB_EPSILON) return false;
  const float inv_det = 1.0f / det;
  const f3 tvec = r.o - v0;
  u = dot3(tvec, pvec) * inv_det;
  if (u < 0.0f || u > 1.0f) return false;
  const f3 qvec = cross3(tvec, e1);
  v = dot3(r.d, qvec) * inv_det;
  if (v < 0.0f || u + v > 1.0f) return false;
  t = dot3(e2, qvec) * inv_det;
  return t > RTB_EPSILON;
}

// Which of two hits a closest-hit query keeps.  Reference flavour (TIE = false): compute:179, strictly closer, so the first
// triangle tested wins among equal t — the reference's traversal order decides.  LBVH flavour (TIE = true): among equal t the
// smaller leaf-order index wins, which makes the result independent of the order triangles are tested in: the per-lane
// ordered traversal, the per-thread one and the packet traversal (trace.cu) all return the same (t, triangle).
template <bool TIE>
__device__ __forceinline__ bool closer_hit(float t, int32_t tri, float best_t, int32_t best_tri) {
  return t < best_t || (TIE && t == best_t && tri < best_tri);
}

// For an analytic primitive the hit record holds (t_world, t_object, face code) in (t, u, v).
template <bool ANALYTIC, bool TIE>
__device__ __forceinline__ void test_triangle_closest(const SceneView& s, const Ray& r, int32_t tri, Hit& best) {
  const float4 a = __ldg(&s.tri_isect[RTB_TRI_F4 * tri]), b = __ldg(&s.tri_isect[RTB_TRI_F4 * tri + 1]), c = __ldg(&s.tri_isect[RTB_TRI_F4 * tri + 2]);
  float t, u, v;
  if (ANALYTIC && __float_as_int(c.w) != 0) {
    int face;
    if (intersect_analytic(&s.prims[6 * __float_as_int(a.x)], __float_as_int(c.w), r.o, r.d, t, u, face) && closer_hit<TIE>(t, tri, best.t, best.tri)) {
      best.t = t; best.u = u; best.v = (float)face; best.tri = tri;
    }
    return;
  }
  if (moller_trumbore(r, mk3(a), mk3(b), mk3(c), t, u, v) && closer_hit<TIE>(t, tri, best.t, best.tri)) { best.t = t; best.u = u; best.v = v; best.tri = tri; }
}
template <bool ANALYTIC>
__device__ __forceinline__ bool test_triangle_any(const SceneView& s, const Ray& r, int32_t tri, float t_limit) {
  const float4 a = __ldg(&s.tri_isect[RTB_TRI_F4 * tri]), b = __ldg(&s.tri_isect[RTB_TRI_F4 * tri + 1]), c = __ldg(&s.tri_isect[RTB_TRI_F4 * tri + 2]);
  float t, u, v;
  if (ANALYTIC && __float_as_int(c.w) != 0) {
    int face;
    return intersect_analytic(&s.prims[6 * __float_as_int(a.x)], __float_as_int(c.w), r.o, r.d, t, u, face) && t <= t_limit;
  }
  return moller_trumbore(r, mk3(a), mk3(b), mk3(c), t, u, v) && t <= t_limit;
}

// ---------------------------------------------------------------------------------------------------------------------
// Reference traversal, TraverseBVH compute:225-267: LIFO stack, left child first, no distance ordering, a node is culled
// when its slab entry distance >= the best t so far.  ANY = shadow query: stops at the first triangle with
// Epsilon < t <= t_limit, which decides "lit" exactly like the reference's closest-hit test (:403-406): lit <=> no such hit.
// ---------------------------------------------------------------------------------------------------------------------
template <bool ANY, bool ANALYTIC>
__device__ __forceinline__ bool traverse_reference(const SceneView& s, const Ray& r, float t_limit, Hit& best, unsigned& overflow) {
  best.t = RTB_INFINITY; best.u = 0.0f; best.v = 0.0f; best.tri = -1;
  if (s.n_nodes == 0) return false;
  int32_t stack[RTB_STACK_REF];
  int sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    const int32_t ni = stack[--sp];
    const float4 lo = __ldg(&s.nodes[2 * ni]), hi = __ldg(&s.nodes[2 * ni + 1]);
    const float dst = slab_entry(r, mk3(lo), mk3(hi));
    if (ANY ? (dst > t_limit) : (dst >= best.t)) continue;
    const int32_t count = __float_as_int(hi.w), left_or_first = __float_as_int(lo.w);
    if (count > 0) {
      for (int32_t i = 0; i < count; i++) {
        if (ANY) { if (test_triangle_any<ANALYTIC>(s, r, left_or_first + i, t_limit)) return true; }
        else test_triangle_closest<ANALYTIC, false>(s, r, left_or_first + i, best);
      }
    } else if (sp + 2 <= RTB_STACK_REF) {
      stack[sp++] = left_or_first + 1;
      stack[sp++] = left_or_first;
    } else {
      overflow++;
    }
  }
  return best.tri >= 0;
}

// Scene data of the traversal kernels comes either from global memory through the read-only path, or — small scenes —
// from the copy the block staged in shared memory (SMEM).
template <bool SMEM>
__device__ __forceinline__ float4 ld4(const float4* p) { return SMEM ? *p : __ldg(p); }
// Two consecutive float4 (32-byte aligned) in ONE 256-bit load through the read-only path (sm_100: LDG.E.ENL2.256.CONSTANT).
// A divergent warp pays L1 wavefronts per instruction and distinct line, so fetching a 64-byte node record with two
// instructions instead of four halves the L1 data-pipe work of a node visit (RTB_LD256=0 keeps the 128-bit loads).
#ifndef RTB_LD256
#define RTB_LD256 1
#endif
template <bool SMEM>
__device__ __forceinline__ void ld8(const float4* p, float4& a, float4& b) {
#if RTB_LD256
  if (!SMEM) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
    return;
  }
#endif
  a = ld4<SMEM>(p); b = ld4<SMEM>(p + 1);
}

// ---------------------------------------------------------------------------------------------------------------------
// One LBVH node visit, shared by the persistent kernel and the per-thread traversals (so both order children the same
// way and pick the same winner among equal t): tests the record's children against the ray, returns the nearest child
// that can still matter and pushes the others, farthest first, with their entry distances.  RTB_REF_MISS: none was hit.
// ---------------------------------------------------------------------------------------------------------------------
#if RTB_LBVH_WIDTH == 4
template <bool SMEM>
__device__ __forceinline__ int32_t lbvh_visit(const float4* nodes, int32_t cur, const SlabRay& sr, float bound, float2* stack,
                                              int& sp, unsigned& overflow) {
  const float4* rec = nodes + 8 * (size_t)cur;
  const float4 mnx = ld4<SMEM>(rec), mny = ld4<SMEM>(rec + 1), mnz = ld4<SMEM>(rec + 2);
  const float4 mxx = ld4<SMEM>(rec + 3), mxy = ld4<SMEM>(rec + 4), mxz = ld4<SMEM>(rec + 5);
  const float4 rf = ld4<SMEM>(rec + 6);
  float e0, e1, e2, e3;
  const bool h0 = slab_hit(sr, mk3(mnx.x, mny.x, mnz.x), mk3(mxx.x, mxy.x, mxz.x), bound, e0);
  const bool h1 = slab_hit(sr, mk3(mnx.y, mny.y, mnz.y), mk3(mxx.y, mxy.y, mxz.y), bound, e1);
  const bool h2 = slab_hit(sr, mk3(mnx.z, mny.z, mnz.z), mk3(mxx.z, mxy.z, mxz.z), bound, e2);
  const bool h3 = slab_hit(sr, mk3(mnx.w, mny.w, mnz.w), mk3(mxx.w, mxy.w, mxz.w), bound, e3);
  int32_t r0 = __float_as_int(rf.x), r1 = __float_as_int(rf.y), r2 = __float_as_int(rf.z), r3 = __float_as_int(rf.w);
  // misses (and unused slots) sort to the end
  if (!h0) e0 = INFINITY;
  if (!h1 || r1 == RTB_REF_DONE) e1 = INFINITY;
  if (!h2 || r2 == RTB_REF_DONE) e2 = INFINITY;
  if (!h3 || r3 == RTB_REF_DONE) e3 = INFINITY;
  const int n_hit = (e0 < INFINITY) + (e1 < INFINITY) + (e2 < INFINITY) + (e3 < INFINITY);
  if (n_hit == 0) return RTB_REF_MISS;
#define RTB_CE(ea, ra, eb, rb) { if (eb < ea) { const float te = ea; ea = eb; eb = te; const int32_t tr = ra; ra = rb; rb = tr; } }
  RTB_CE(e0, r0, e1, r1) RTB_CE(e2, r2, e3, r3) RTB_CE(e0, r0, e2, r2) RTB_CE(e1, r1, e3, r3) RTB_CE(e1, r1, e2, r2)
#undef RTB_CE
  if (sp + n_hit - 1 > RTB_STACK_LBVH) { overflow++; return r0; }
  if (n_hit > 3) { stack[sp] = make_float2(e3, __int_as_float(r3)); sp++; }
  if (n_hit > 2) { stack[sp] = make_float2(e2, __int_as_float(r2)); sp++; }
  if (n_hit > 1) { stack[sp] = make_float2(e1, __int_as_float(r1)); sp++; }
  return r0;
}
#else
// Software prefetch (build-time experiments, tools/build_variants.sh): RTB_PREFETCH_PUSH 1 / 2 = when a child is deferred, ask L1 / L2
// for its record, so that the pop which visits it later does not pay the full miss; RTB_PREFETCH_LEAF 1 = a lane that has
// reached a leaf asks for its triangle records while the other lanes of the warp still descend.
#ifndef RTB_PREFETCH_PUSH
#define RTB_PREFETCH_PUSH 0
#endif
#ifndef RTB_PREFETCH_LEAF
#define RTB_PREFETCH_LEAF 0
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_leaf(const float4* tri_isect, int32_t leaf_ref) {
  const int32_t code = ~leaf_ref;
  const int32_t first = code >> 3, count = (code & 7) + 1;
  prefetch_l1(&tri_isect[RTB_TRI_F4 * (size_t)first]);
  if (count > 2) prefetch_l1(&tri_isect[RTB_TRI_F4 * (size_t)(first + count) - 1]);
}

// STRIDE: distance between consecutive stack entries (1: a thread-private array; RTB_POOL_SLOTS: the slot-interleaved scratch of
// k_traverse_pool).
template <bool SMEM, int STRIDE = 1>
__device__ __forceinline__ int32_t lbvh_visit(const float4* nodes, int32_t cur, const SlabRay& sr, float bound, float2* stack,
                                              int& sp, unsigned& overflow) {
  float4 n0, n1, n2, n3;
  const float4* rec = nodes + 4 * (size_t)(uint32_t)cur;  // one widening multiply-add (4 * cur in 32 bits costs a shift first)
  ld8<SMEM>(rec, n0, n1);
  ld8<SMEM>(rec + 2, n2, n3);
  float dl, dr;
  const bool hl = slab_hit(sr, mk3(n0), mk3(n1), bound, dl);
  const bool hr = slab_hit(sr, mk3(n2), mk3(n3), bound, dr);
  const int32_t lref = __float_as_int(n0.w), rref = __float_as_int(n1.w);
  if (hl && hr) {
    const bool left_first = !(dr < dl);
    if (sp < RTB_STACK_LBVH) { stack[(size_t)sp * STRIDE] = make_float2(left_first ? dr : dl, __int_as_float(left_first ? rref : lref)); sp++; }
    else overflow++;
#if RTB_PREFETCH_PUSH
    if (!SMEM) {
      const int32_t far = left_first ? rref : lref;
      if (far >= 0) { if (RTB_PREFETCH_PUSH == 1) prefetch_l1(&nodes[4 * (size_t)far]); else prefetch_l2(&nodes[4 * (size_t)far]); }
    }
#endif
    return left_first ? lref : rref;
  }
  if (hl) return lref;
  if (hr) return rref;
  return RTB_REF_MISS;
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// LBVH traversal: 64-byte nodes holding both children's boxes, near child first, deferred child kept with its entry
// distance so it can be dropped without a fetch once a closer hit is known.  Triangle arithmetic is the same as above,
// so t/u/v of a given (ray, triangle) pair are bit-identical in both modes; the box test is the FMA form over padded
// boxes.  (This per-thread form serves the aux / debug kernels; the wavefront uses the persistent form in trace.cu.)
// Node layout (4 x float4): (lmin, left_ref) (lmax, right_ref) (rmin, -) (rmax, -).
// ---------------------------------------------------------------------------------------------------------------------
template <bool ANY, bool ANALYTIC>
__device__ __forceinline__ bool traverse_lbvh(const SceneView& s, const Ray& r, float t_limit, Hit& best, unsigned& overflow) {
  best.t = RTB_INFINITY; best.u = 0.0f; best.v = 0.0f; best.tri = -1;
  if (s.n_tris == 0) return false;
  float2 stack[RTB_STACK_LBVH];  // deferred children: (entry distance, node / leaf reference) in one 8-byte local-memory access
  int sp = 0;
  int32_t cur = s.root;
  const SlabRay sr = make_slab_ray(r.o, r.d);
  for (;;) {
    if (cur >= 0) {
      // same box test, bound and child order as k_traverse_lbvh
      const float bound = ANY ? nextafterf(t_limit, INFINITY) : best.t;
      const int32_t next = lbvh_visit<false>(s.nodes, cur, sr, bound, stack, sp, overflow);
      if (next != RTB_REF_MISS) { cur = next; continue; }
    } else {
      const int32_t code = ~cur;
      const int32_t first = code >> 3, count = (code & 7) + 1;
      for (int32_t i = 0; i < count; i++) {
        if (ANY) { if (test_triangle_any<ANALYTIC>(s, r, first + i, t_limit)) return true; }
        else test_triangle_closest<ANALYTIC, true>(s, r, first + i, best);
      }
    }
    // pop the next deferred child that can still matter
    for (;;) {
      if (sp == 0) return best.tri >= 0;
      sp--;
      const float2 e = stack[sp];
      if (ANY ? !(e.x > t_limit) : !(e.x > best.t)) { cur = __float_as_int(e.y); break; }  // '>': a triangle tying with the best hit may sit exactly at the entry
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// 8-wide quantised LBVH (RTB_BVH_WIDE: RTB_BVH_LBVH scenes when RTB_WIDE=1; lbvh.cu: k_wide_level builds it from the binary radix
// tree).  One 96-byte record = three 32-byte sectors = three 256-bit loads decides eight children, so a ray makes about a
// third of the dependent node fetches of the two-box binary records and half the L1 wavefronts.  Record (24 words):
//   w0-2  p = min corner of the node's box          w3  bytes: ESx, ESy, ESz (biased exponents of S = 2^15 * cell), valid mask
//   w4-7  child references of slots 0-3             w8-15  lo.x[8] lo.y[8] lo.z[8] hi.x[8]   (one byte per slot)
//   w16-19 hi.y[8] hi.z[8]                          w20-23 child references of slots 4-7
// A child box is p + q * cell per axis with q rounded outward to 8 bits.  The byte goes into mantissa bits 8-15 of 1.0f by ONE
// byte-permute (f = 1 + q * 2^-15, no integer->float conversion), and a slab plane is t = f * A + B with A = S / d and
// B = (p - o) / d - A.  The rounding of that form is at most 2^-22 (|(p - o) / d| + |A|); the near planes are moved earlier and
// the far planes later by 2^-20 of the same sum (1/32 of a cell), so every box the exact test would enter is entered whatever
// the distance between camera and scene (no build-time padding, nothing to assume about the camera).  The boxes decide only
// which nodes are opened; t, u, v and ids come from the unchanged triangle test.
// Slot s sits on the + side of x / y / z where bit 2 / 1 / 0 of s is set (greedy assignment at build time), so visiting the hit
// slots in increasing order of (s XOR the ray's negative-direction bits) goes roughly front to back without any distances.
// ---------------------------------------------------------------------------------------------------------------------
#define RTB_STACK_WIDE 48  /* deferred sibling groups: at most one per level of the wide tree */

__device__ __forceinline__ unsigned octant_of(f3 d) { return (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u); }
// m'[k] = m[k ^ neg] for an 8-bit mask
__device__ __forceinline__ unsigned octant_permute(unsigned m, unsigned neg) {
  if (neg & 1u) m = ((m & 0x55u) << 1) | ((m >> 1) & 0x55u);
  if (neg & 2u) m = ((m & 0x33u) << 2) | ((m >> 2) & 0x33u);
  if (neg & 4u) m = ((m & 0x0fu) << 4) | (m >> 4);
  return m;
}
__device__ __forceinline__ float quant_to_float(unsigned word, unsigned selector) { return __uint_as_float(__byte_perm(word, 0x3F800000u, selector)); }

// Tests the eight child boxes of the record at `rec` against the ray; returns the hit mask over slots and the slots' references.
template <bool SMEM>
__device__ __forceinline__ unsigned wide_test(const float4* rec, f3 o, f3 inv, unsigned neg, float bound, int32_t (&refs)[8]) {
  float4 a0, a1, b0, b1, c0, c1;
  ld8<SMEM>(rec, a0, a1);
  ld8<SMEM>(rec + 2, b0, b1);
  ld8<SMEM>(rec + 4, c0, c1);
  refs[0] = __float_as_int(a1.x); refs[1] = __float_as_int(a1.y); refs[2] = __float_as_int(a1.z); refs[3] = __float_as_int(a1.w);
  refs[4] = __float_as_int(c1.x); refs[5] = __float_as_int(c1.y); refs[6] = __float_as_int(c1.z); refs[7] = __float_as_int(c1.w);
  const unsigned hdr = __float_as_uint(a0.w);
  const float ax = inv.x * __uint_as_float((hdr & 0xffu) << 23), ay = inv.y * __uint_as_float((hdr & 0xff00u) << 15), az = inv.z * __uint_as_float((hdr & 0xff0000u) << 7);
  const float ox = (a0.x - o.x) * inv.x, oy = (a0.y - o.y) * inv.y, oz = (a0.z - o.z) * inv.z;
  const float ex = (fabsf(ox) + fabsf(ax)) * 9.5367431640625e-7f, ey = (fabsf(oy) + fabsf(ay)) * 9.5367431640625e-7f, ez = (fabsf(oz) + fabsf(az)) * 9.5367431640625e-7f;
  const float bx = ox - ax, by = oy - ay, bz = oz - az;
  const float bnx = bx - ex, bny = by - ey, bnz = bz - ez, bfx = bx + ex, bfy = by + ey, bfz = bz + ez;
  // per axis: the planes a ray moving in + direction meets first are the lo planes
  const bool px = (neg & 4u) == 0u, py = (neg & 2u) == 0u, pz = (neg & 1u) == 0u;
  const unsigned lox[2] = {__float_as_uint(b0.x), __float_as_uint(b0.y)}, loy[2] = {__float_as_uint(b0.z), __float_as_uint(b0.w)};
  const unsigned loz[2] = {__float_as_uint(b1.x), __float_as_uint(b1.y)}, hix[2] = {__float_as_uint(b1.z), __float_as_uint(b1.w)};
  const unsigned hiy[2] = {__float_as_uint(c0.x), __float_as_uint(c0.y)}, hiz[2] = {__float_as_uint(c0.z), __float_as_uint(c0.w)};
  const unsigned nx[2] = {px ? lox[0] : hix[0], px ? lox[1] : hix[1]}, fx[2] = {px ? hix[0] : lox[0], px ? hix[1] : lox[1]};
  const unsigned ny[2] = {py ? loy[0] : hiy[0], py ? loy[1] : hiy[1]}, fy[2] = {py ? hiy[0] : loy[0], py ? hiy[1] : loy[1]};
  const unsigned nz[2] = {pz ? loz[0] : hiz[0], pz ? loz[1] : hiz[1]}, fz[2] = {pz ? hiz[0] : loz[0], pz ? hiz[1] : loz[1]};
  unsigned mask = 0u;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int w = c >> 2;
    const unsigned sel = 0x7604u | ((unsigned)(c & 3) << 4);
    const float tnx = __fmaf_rn(quant_to_float(nx[w], sel), ax, bnx), tny = __fmaf_rn(quant_to_float(ny[w], sel), ay, bny), tnz = __fmaf_rn(quant_to_float(nz[w], sel), az, bnz);
    const float tfx = __fmaf_rn(quant_to_float(fx[w], sel), ax, bfx), tfy = __fmaf_rn(quant_to_float(fy[w], sel), ay, bfy), tfz = __fmaf_rn(quant_to_float(fz[w], sel), az, bfz);
    const float entry = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.0f));
    const float exit = fminf(fminf(tfx, tfy), fminf(tfz, bound));
    if (entry <= exit) mask |= 1u << c;
  }
  return mask & (hdr >> 24);
}

__device__ __forceinline__ int32_t select_ref(const int32_t (&refs)[8], unsigned slot) {
  const int32_t a = (slot & 1u) ? refs[1] : refs[0], b = (slot & 1u) ? refs[3] : refs[2], c = (slot & 1u) ? refs[5] : refs[4], d = (slot & 1u) ? refs[7] : refs[6];
  const int32_t ab = (slot & 2u) ? b : a, cd = (slot & 2u) ? d : c;
  return (slot & 4u) ? cd : ab;
}
template <bool SMEM>
__device__ __forceinline__ int32_t load_ref(const float4* nodes, int32_t node, unsigned slot) {
  const int32_t* w = (const int32_t*)(nodes + RTB_WIDE_F4 * (size_t)node);
  const int32_t* p = &w[(slot < 4u ? 4 : 16) + (int)slot];
  return SMEM ? *p : __ldg(p);
}

// One step of the wide traversal: visits node `cur`, returns the nearest child still to be looked at (a node >= 0 or a leaf
// reference < 0) and defers its hit siblings as ONE stack entry (node, octant-ordered mask); RTB_REF_DONE when the ray is done.
template <bool SMEM>
__device__ __forceinline__ int32_t wide_pop(const float4* nodes, unsigned neg, uint2* stack, int& sp) {
  if (sp == 0) return RTB_REF_DONE;
  const uint2 e = stack[sp - 1];
  const unsigned k = (unsigned)__ffs((int)e.y) - 1u, rest = e.y & (e.y - 1u);
  if (rest) stack[sp - 1].y = rest; else sp--;
  return load_ref<SMEM>(nodes, (int32_t)e.x, k ^ neg);
}
template <bool SMEM>
__device__ __forceinline__ int32_t wide_visit(const float4* nodes, int32_t cur, f3 o, f3 inv, unsigned neg, float bound, uint2* stack, int& sp, unsigned& overflow) {
  int32_t refs[8];
  unsigned m = octant_permute(wide_test<SMEM>(nodes + RTB_WIDE_F4 * (size_t)cur, o, inv, neg, bound, refs), neg);
  if (m == 0u) return wide_pop<SMEM>(nodes, neg, stack, sp);
  const unsigned k = (unsigned)__ffs((int)m) - 1u;
  m &= m - 1u;
  if (m) {
    if (sp < RTB_STACK_WIDE) { stack[sp] = make_uint2((unsigned)cur, m); sp++; }
    else overflow++;
  }
  return select_ref(refs, k ^ neg);
}

// Per-thread closest-hit / any-hit query over the wide tree (k_tail, k_aux, k_debug; the wavefront uses the persistent form).
template <bool ANY, bool ANALYTIC>
__device__ __forceinline__ bool traverse_wide(const SceneView& s, const Ray& r, float t_limit, Hit& best, unsigned& overflow) {
  best.t = RTB_INFINITY; best.u = 0.0f; best.v = 0.0f; best.tri = -1;
  if (s.n_tris == 0) return false;
  uint2 stack[RTB_STACK_WIDE];
  int sp = 0;
  int32_t cur = s.root;
  const f3 inv = safe_inverse(r.d);
  const unsigned neg = octant_of(r.d);
  const float any_bound = nextafterf(t_limit, INFINITY);
  while (cur != RTB_REF_DONE) {
    if (cur >= 0) { cur = wide_visit<false>(s.nodes, cur, r.o, inv, neg, ANY ? any_bound : best.t, stack, sp, overflow); continue; }
    const int32_t code = ~cur;
    const int32_t first = code >> 3, count = (code & 7) + 1;
    for (int32_t i = 0; i < count; i++) {
      if (ANY) { if (test_triangle_any<ANALYTIC>(s, r, first + i, t_limit)) return true; }
      else test_triangle_closest<ANALYTIC, true>(s, r, first + i, best);
    }
    cur = wide_pop<false>(s.nodes, neg, stack, sp);
  }
  return best.tri >= 0;
}

template <int BVH, bool ANY, bool ANALYTIC>
__device__ __forceinline__ bool traverse(const SceneView& s, const Ray& r, float t_limit, Hit& best, unsigned& overflow) {
  if (BVH == RTB_BVH_REFERENCE) return traverse_reference<ANY, ANALYTIC>(s, r, t_limit, best, overflow);
  if (BVH == RTB_BVH_WIDE) return traverse_wide<ANY, ANALYTIC>(s, r, t_limit, best, overflow);
  return traverse_lbvh<ANY, ANALYTIC>(s, r, t_limit, best, overflow);
}

// Position and shading normal of a closest hit: triangle (compute:183-187) or analytic primitive.
template <bool ANALYTIC>
__device__ __forceinline__ void hit_surface(const SceneView& s, const Ray& ray, const Hit& h, f3& pos, f3& nrm);

// Interpolated normal of a hit, compute:186-187
__device__ __forceinline__ f3 hit_normal(const SceneView& s, const Hit& h) {
  const float4 n0 = __ldg(&s.tri_shade[3 * h.tri]), n1 = __ldg(&s.tri_shade[3 * h.tri + 1]), n2 = __ldg(&s.tri_shade[3 * h.tri + 2]);
  const float w = 1.0f - h.u - h.v;
  return hlsl_normalize((w * mk3(n0) + h.u * mk3(n1)) + h.v * mk3(n2));
}

template <bool ANALYTIC>
__device__ __forceinline__ void hit_surface(const SceneView& s, const Ray& ray, const Hit& h, f3& pos, f3& nrm) {
  if (ANALYTIC) {
    const int kind = __float_as_int(__ldg(&s.tri_isect[RTB_TRI_F4 * h.tri + 2]).w);
    if (kind != 0) {
      const int idx = __float_as_int(__ldg(&s.tri_isect[RTB_TRI_F4 * h.tri]).x);
      analytic_surface(&s.prims[6 * idx], kind, ray.o, ray.d, h.u, (int)h.v, pos, nrm);
      return;
    }
  }
  pos = ray.o + h.t * ray.d;
  nrm = hit_normal(s, h);
}

struct Material { f3 color; float ka, kd, ks, kr, ior; };
// compute:371-376; an index outside the material buffer takes the same defaults
__device__ __forceinline__ Material fetch_material(const SceneView& s, int32_t index) {
  Material m;
  m.color = mk3(1.0f, 1.0f, 1.0f); m.ka = 0.1f; m.kd = 0.7f; m.ks = 0.0f; m.kr = 0.0f; m.ior = 1.0f;
  if (index >= 0 && index < s.n_mats) {
    const float4 a = __ldg(&s.materials[2 * index]), b = __ldg(&s.materials[2 * index + 1]);
    m.color = mk3(a); m.ka = a.w; m.kd = b.x; m.ks = b.y; m.kr = b.z; m.ior = b.w;
  }
  return m;
}

// byte = floor(saturate(c) * 255 + 0.5); NaN -> 0.  srgb != 0 applies the sRGB OETF first.
__device__ __forceinline__ unsigned quantize_unorm8(float c, int srgb) {
  if (!(c == c)) c = 0.0f;
  c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
  if (srgb) c = c <= 0.0031308f ? 12.92f * c : 1.055f * powf(c, 1.0f / 2.4f) - 0.055f;
  return (unsigned)(int)floorf(c * 255.0f + 0.5f);
}

}  // namespace rtb
