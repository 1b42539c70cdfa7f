// flatten.cu — K1, scene upload on the device: object-space flattening of meshes, boxes and spheres into triangle arrays
// (replaces SceneGeometryConverter.ExtractTriangles, Assets/Services/SceneGeometryConverter.cs:18-264, which appends to a
// C# list on the CPU), and the gather into BVH leaf order (BVHBuilder.Flatten's triangle re-emission,
// Assets/Services/BVH/BVHBuilder.cs:224-227).  One thread per emitted triangle; the composite matrices, inverse-transpose
// normal matrices and the 402-vertex unit-sphere table come from the host (they need libm), everything here is
// + - * / sqrt in the reference's operation order, so the arrays are bit-identical to the CPU restatement's.
#include "kernels.hpp"

namespace rtb {
namespace {

constexpr int kBlock = 256;

// Matrix4x4.MultiplyPoint3x4 / MultiplyVector
__device__ __forceinline__ f3 mul_point(const float* m, f3 v) {
  return mk3(((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3], ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7],
             ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11]);
}
__device__ __forceinline__ f3 mul_vector3(const float* n, f3 v) {
  return mk3((n[0] * v.x + n[1] * v.y) + n[2] * v.z, (n[3] * v.x + n[4] * v.y) + n[5] * v.z, (n[6] * v.x + n[7] * v.y) + n[8] * v.z);
}

// AddCube's 12 triangles as corner indices, SceneGeometryConverter.cs:139-154
__constant__ unsigned char kCubeIndex[12][3] = {{0, 2, 1}, {0, 3, 2}, {5, 7, 6}, {5, 4, 7}, {3, 6, 2}, {3, 7, 6},
                                               {4, 1, 5}, {4, 0, 1}, {4, 3, 7}, {4, 0, 3}, {1, 6, 2}, {1, 5, 6}};

__global__ void __launch_bounds__(kBlock) k_flatten(const float* __restrict__ tri_in, const FlattenObject* __restrict__ objs, int n_objs,
                                                    const float* __restrict__ sphere, int32_t n_out, float4* __restrict__ raw,
                                                    float4* __restrict__ nrm) {
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) {
    // owning object: last entry with out_first <= i
    int lo = 0, hi = n_objs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (objs[mid].out_first <= i) lo = mid; else hi = mid - 1;
    }
    const FlattenObject& ob = objs[lo];
    const int32_t k = i - ob.out_first;
    f3 a, b, c, na, nb, nc;
    int32_t material = ob.material;
    if (ob.kind == OBJ_MESH) {  // :23-34 + CreateGPUTriangle :56-60
      const float* t = tri_in + (size_t)(ob.src_first + k) * 10;
      material = __float_as_int(t[0]);
      a = mul_point(ob.m, mk3(t[1], t[2], t[3]));
      b = mul_point(ob.m, mk3(t[4], t[5], t[6]));
      c = mul_point(ob.m, mk3(t[7], t[8], t[9]));
      na = nb = nc = unity_normalized(cross3(b - a, c - a));
    } else if (ob.kind == OBJ_BOX) {  // AddCube :120-155
      f3 v[3];
      for (int j = 0; j < 3; j++) {
        const int ci = kCubeIndex[k][j];
        const f3 corner = mk3((ci == 1 || ci == 2 || ci == 5 || ci == 6) ? 0.5f : -0.5f, (ci == 2 || ci == 3 || ci == 6 || ci == 7) ? 0.5f : -0.5f,
                              ci >= 4 ? 0.5f : -0.5f);
        v[j] = mul_point(ob.m, corner);
      }
      a = v[0]; b = v[1]; c = v[2];
      na = nb = nc = unity_normalized(cross3(b - a, c - a));
    } else if (ob.kind == OBJ_BOX_ANALYTIC || ob.kind == OBJ_SPHERE_ANALYTIC) {
      // Analytic primitive: its entry is the world-space bounding box of the 8 transformed corners of the unit cube
      // (+-0.5) or of the unit sphere's cube (+-1), SphereInstance / BoxInstance constructors, HittableObjects.cs:22-39, 129-145.
      // Rows (bmin, bmax, bmin) let every builder treat it like a triangle whose three "vertices" span that box.
      const float h = ob.kind == OBJ_BOX_ANALYTIC ? 0.5f : 1.0f;
      f3 bmin = mk3(INFINITY, INFINITY, INFINITY), bmax = mk3(-INFINITY, -INFINITY, -INFINITY);
      for (int ci = 0; ci < 8; ci++) {
        const f3 p = mul_point(ob.m, mk3((ci & 1) ? h : -h, (ci & 2) ? h : -h, (ci & 4) ? h : -h));
        bmin = mk3(fminf(bmin.x, p.x), fminf(bmin.y, p.y), fminf(bmin.z, p.z));
        bmax = mk3(fmaxf(bmax.x, p.x), fmaxf(bmax.y, p.y), fmaxf(bmax.z, p.z));
      }
      const f3 ctr = (bmin + bmax) * 0.5f;
      raw[3 * (size_t)i] = make_float4(bmin.x, bmin.y, bmin.z, ctr.x);
      raw[3 * (size_t)i + 1] = make_float4(bmax.x, bmax.y, bmax.z, ctr.y);
      raw[3 * (size_t)i + 2] = make_float4(bmin.x, bmin.y, bmin.z, ctr.z);
      nrm[3 * (size_t)i] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(material));
      nrm[3 * (size_t)i + 1] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(ob.kind == OBJ_SPHERE_ANALYTIC ? 1 : 2));
      nrm[3 * (size_t)i + 2] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(ob.src_first));
      continue;
    } else {  // AddSphere :192-229 + AddSmoothTri :245-264
      int ia, ib, ic;
      const int last = kSphereVerts - 1;
      if (k < 24) { ia = 0; ib = k + 2; ic = k + 1; }
      else if (k < 24 + 720) {
        const int kk = k - 24, lat = kk / 48, rem = kk - lat * 48, lon = rem >> 1;
        const int cur = lon + lat * 25 + 1, next = cur + 1, below = cur + 25, below_next = below + 1;
        if ((rem & 1) == 0) { ia = cur; ib = below; ic = next; } else { ia = next; ib = below; ic = below_next; }
      } else { const int lon = k - 744; ia = last; ib = last - 25 + lon; ic = last - 25 + lon + 1; }
      const f3 pa = mk3(sphere[3 * ia], sphere[3 * ia + 1], sphere[3 * ia + 2]);
      const f3 pb = mk3(sphere[3 * ib], sphere[3 * ib + 1], sphere[3 * ib + 2]);
      const f3 pc = mk3(sphere[3 * ic], sphere[3 * ic + 1], sphere[3 * ic + 2]);
      a = mul_point(ob.m, pa); b = mul_point(ob.m, pb); c = mul_point(ob.m, pc);
      na = unity_normalized(mul_vector3(ob.nm, unity_normalized(pa)));
      nb = unity_normalized(mul_vector3(ob.nm, unity_normalized(pb)));
      nc = unity_normalized(mul_vector3(ob.nm, unity_normalized(pc)));
    }
    const f3 sum = (a + b) + c;  // center = (v0 + v1 + v2) / 3.0f, :74
    raw[3 * (size_t)i] = make_float4(a.x, a.y, a.z, sum.x / 3.0f);
    raw[3 * (size_t)i + 1] = make_float4(b.x, b.y, b.z, sum.y / 3.0f);
    raw[3 * (size_t)i + 2] = make_float4(c.x, c.y, c.z, sum.z / 3.0f);
    nrm[3 * (size_t)i] = make_float4(na.x, na.y, na.z, __int_as_float(material));
    nrm[3 * (size_t)i + 1] = make_float4(nb.x, nb.y, nb.z, 0.0f);
    nrm[3 * (size_t)i + 2] = make_float4(nc.x, nc.y, nc.z, 0.0f);
  }
}

__global__ void __launch_bounds__(kBlock) k_pack(const float4* __restrict__ raw, const float4* __restrict__ nrm, const int32_t* __restrict__ perm,
                                                 int32_t n, float4* __restrict__ isect, float4* __restrict__ shade) {
  for (int32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int32_t src = perm[j];
    const float4 a = raw[3 * (size_t)src], b = raw[3 * (size_t)src + 1], c = raw[3 * (size_t)src + 2];
    const float4 n0 = nrm[3 * (size_t)src], n1 = nrm[3 * (size_t)src + 1], n2 = nrm[3 * (size_t)src + 2];
    if (RTB_TRI_F4 > 3) isect[RTB_TRI_F4 * (size_t)j + 3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (__float_as_int(n1.w) != 0) {  // analytic primitive: (table index, -, -, prim_id) (-, -, -, material) (-, -, -, kind)
      isect[RTB_TRI_F4 * (size_t)j] = make_float4(n2.w, 0.0f, 0.0f, __int_as_float(src));
      isect[RTB_TRI_F4 * (size_t)j + 1] = make_float4(0.0f, 0.0f, 0.0f, n0.w);
      isect[RTB_TRI_F4 * (size_t)j + 2] = make_float4(0.0f, 0.0f, 0.0f, n1.w);
      shade[3 * (size_t)j] = shade[3 * (size_t)j + 1] = shade[3 * (size_t)j + 2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      continue;
    }
    // edges exactly as IntersectTriangle forms them (BVHRayTracing.compute:155-156): v1 - v0, v2 - v0
    isect[RTB_TRI_F4 * (size_t)j] = make_float4(a.x, a.y, a.z, __int_as_float(src));
    isect[RTB_TRI_F4 * (size_t)j + 1] = make_float4(b.x - a.x, b.y - a.y, b.z - a.z, n0.w);
    isect[RTB_TRI_F4 * (size_t)j + 2] = make_float4(c.x - a.x, c.y - a.y, c.z - a.z, 0.0f);
    shade[3 * (size_t)j] = make_float4(n0.x, n0.y, n0.z, 0.0f);
    shade[3 * (size_t)j + 1] = make_float4(n1.x, n1.y, n1.z, 0.0f);
    shade[3 * (size_t)j + 2] = make_float4(n2.x, n2.y, n2.z, 0.0f);
  }
}

inline int grid_for(int64_t n) {
  const int64_t g = (n + kBlock - 1) / kBlock;
  return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}

}  // namespace

void launch_flatten(const float* tri_in, const FlattenObject* objs, int n_objs, const float* sphere_table, int32_t n_out, float4* raw, float4* nrm,
                    cudaStream_t st) {
  if (n_out <= 0) return;
  k_flatten<<<grid_for(n_out), kBlock, 0, st>>>(tri_in, objs, n_objs, sphere_table, n_out, raw, nrm);
}

void launch_pack(const float4* raw, const float4* nrm, const int32_t* perm, int32_t n, float4* isect, float4* shade, cudaStream_t st) {
  if (n <= 0) return;
  k_pack<<<grid_for(n), kBlock, 0, st>>>(raw, nrm, perm, n, isect, shade);
}

}  // namespace rtb
