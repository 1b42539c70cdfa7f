// tools/bvh_study.cpp — OFFLINE PLANNING TOOL (not product, not oracle): how many node records and triangle tests does the
// library's traversal scheme cost per ray under different BVH builders?  Built and driven by tools/bvh_study.py.
//
//   bvh_study <triangles.bin> <rays.bin>
//     triangles.bin: int64 n, then n x 9 float32 (v0 v1 v2)
//     rays.bin:      int64 m, then m x 7 float32 (origin, direction, t_max; t_max <= 0: closest-hit ray, > 0: any-hit up to t_max)
//
// Builders (all binary trees over two-box records like csrc/lbvh.cu emits, leaves of <= 4 triangles):
//   lbvh    63-bit Morton codes of centroids over the bounding CUBE, split at the highest differing bit (what lbvh.cu builds)
//   sah     top-down binned SAH (16 bins per axis, all three axes), the usual quality yardstick
//   median  spatial median of the longest axis (what the reference's BVHBuilder.cs does)
// Traversal = the ordered scheme of trace.cuh (near child first, deferred child dropped when its entry >= best t; any-hit rays
// stop at the first triangle in range).  Counts per ray: records visited (one record = both child boxes), triangles tested.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>

struct V3 { float x, y, z; };
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct Box {
  V3 lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
  void add(V3 p) { lo = {std::min(lo.x, p.x), std::min(lo.y, p.y), std::min(lo.z, p.z)}; hi = {std::max(hi.x, p.x), std::max(hi.y, p.y), std::max(hi.z, p.z)}; }
  void add(const Box& b) { add(b.lo); add(b.hi); }
  float area() const { const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z; return dx < 0 ? 0.0f : 2.0f * (dx * dy + dy * dz + dz * dx); }
};
struct Tri { V3 v0, v1, v2; };
struct Node { Box box[2]; int32_t child[2]; };  // child >= 0: node; < 0: leaf ~(first << 3 | count - 1)

static std::vector<Tri> tris;
static std::vector<Box> tbox;
static std::vector<V3> cen;
static std::vector<int32_t> order;  // leaf order -> triangle
static std::vector<Node> nodes;

static Box range_box(int b, int e) { Box r; for (int i = b; i < e; i++) r.add(tbox[order[i]]); return r; }
static int32_t leaf_ref(int first, int count) { return ~((first << 3) | (count - 1)); }

enum Builder { LBVH, SAH, MEDIAN, LBVH_SCHED };
static std::vector<uint64_t> morton;  // per position in `order` (LBVH only)

static inline uint64_t spread21(uint64_t v) {
  v &= 0x1fffff;
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}

// returns child reference for range [b, e)
static int32_t build(Builder how, int b, int e, int bit) {
  const int n = e - b;
  if (n <= 4) return leaf_ref(b, n);
  int mid = -1;
  if (how == LBVH || how == LBVH_SCHED) {
    while (bit >= 0) {  // first position whose code has `bit` set
      const uint64_t mask = 1ULL << bit;
      if ((morton[b] & mask) == (morton[e - 1] & mask)) { bit--; continue; }
      mid = (int)(std::partition_point(morton.begin() + b, morton.begin() + e, [&](uint64_t c) { return (c & mask) == 0; }) - morton.begin());
      bit--;
      break;
    }
    if (mid < 0) mid = (b + e) / 2;  // identical codes
  } else {
    Box cb;
    for (int i = b; i < e; i++) cb.add(cen[order[i]]);
    const float ext[3] = {cb.hi.x - cb.lo.x, cb.hi.y - cb.lo.y, cb.hi.z - cb.lo.z};
    auto coord = [&](int t, int ax) { const V3& c = cen[t]; return ax == 0 ? c.x : (ax == 1 ? c.y : c.z); };
    const float lo3[3] = {cb.lo.x, cb.lo.y, cb.lo.z};
    int best_ax = -1;
    float best_pos = 0.0f;
    if (how == MEDIAN) {
      best_ax = ext[1] > ext[0] ? 1 : 0;
      if (ext[2] > ext[best_ax]) best_ax = 2;
      best_pos = lo3[best_ax] + 0.5f * ext[best_ax];
    } else {
      constexpr int B = 16;
      float best_cost = 1e30f;
      for (int ax = 0; ax < 3; ax++) {
        if (!(ext[ax] > 0.0f)) continue;
        Box bb[B]; int cnt[B] = {0};
        const float scale = B / ext[ax];
        for (int i = b; i < e; i++) {
          int k = (int)((coord(order[i], ax) - lo3[ax]) * scale);
          k = k < 0 ? 0 : (k >= B ? B - 1 : k);
          bb[k].add(tbox[order[i]]); cnt[k]++;
        }
        float right_area[B]; int right_cnt[B];
        Box acc; int c = 0;
        for (int k = B - 1; k > 0; k--) { acc.add(bb[k]); c += cnt[k]; right_area[k] = acc.area(); right_cnt[k] = c; }
        Box left; int lc = 0;
        for (int k = 0; k < B - 1; k++) {
          left.add(bb[k]); lc += cnt[k];
          if (lc == 0 || right_cnt[k + 1] == 0) continue;
          const float cost = left.area() * lc + right_area[k + 1] * right_cnt[k + 1];
          if (cost < best_cost) { best_cost = cost; best_ax = ax; best_pos = lo3[ax] + (k + 1) / scale; }
        }
      }
    }
    if (best_ax >= 0) mid = (int)(std::partition(order.begin() + b, order.begin() + e, [&](int t) { return coord(t, best_ax) < best_pos; }) - order.begin());
    if (mid <= b || mid >= e) {  // degenerate: split by count along the longest axis
      int ax = ext[1] > ext[0] ? 1 : 0;
      if (ext[2] > ext[ax]) ax = 2;
      mid = (b + e) / 2;
      std::nth_element(order.begin() + b, order.begin() + mid, order.begin() + e, [&](int x, int y) { return coord(x, ax) < coord(y, ax); });
    }
  }
  const int32_t me = (int32_t)nodes.size();
  nodes.push_back(Node());
  const int32_t l = build(how, b, mid, bit), r = build(how, mid, e, bit);
  nodes[me].child[0] = l; nodes[me].child[1] = r;
  nodes[me].box[0] = range_box(b, mid); nodes[me].box[1] = range_box(mid, e);
  return me;
}

static inline bool slab(const Box& bx, V3 o, V3 inv, float bound, float& entry) {
  const float t0x = (bx.lo.x - o.x) * inv.x, t1x = (bx.hi.x - o.x) * inv.x;
  const float t0y = (bx.lo.y - o.y) * inv.y, t1y = (bx.hi.y - o.y) * inv.y;
  const float t0z = (bx.lo.z - o.z) * inv.z, t1z = (bx.hi.z - o.z) * inv.z;
  entry = std::max(std::max(std::min(t0x, t1x), std::min(t0y, t1y)), std::max(std::min(t0z, t1z), 0.0f));
  const float exit = std::min(std::min(std::max(t0x, t1x), std::max(t0y, t1y)), std::min(std::max(t0z, t1z), bound));
  return entry <= exit;
}
static inline bool hit_tri(const Tri& t, V3 o, V3 d, float& tt) {
  const V3 e1 = t.v1 - t.v0, e2 = t.v2 - t.v0, p = cross(d, e2);
  const float det = dot(e1, p);
  if (std::fabs(det) < 1e-4f) return false;
  const float inv = 1.0f / det;
  const V3 tv = o - t.v0;
  const float u = dot(tv, p) * inv;
  if (u < 0 || u > 1) return false;
  const V3 q = cross(tv, e1);
  const float v = dot(d, q) * inv;
  if (v < 0 || u + v > 1) return false;
  tt = dot(e2, q) * inv;
  return tt > 1e-4f;
}

struct Counts { double records = 0, tris = 0, hits = 0; long long rays = 0; int max_records = 0; };
static std::vector<int> per_ray;  // node records + triangle tests of every ray (steps of the while-while loop)

static void trace(int32_t root, const float* rays, long long m, Counts& c) {
#pragma omp parallel
  {
    Counts local;
#pragma omp for schedule(dynamic, 1024)
    for (long long r = 0; r < m; r++) {
      const float* q = rays + 7 * r;
      const V3 o{q[0], q[1], q[2]}, d{q[3], q[4], q[5]};
      const bool any = q[6] > 0.0f;
      float best = any ? std::nextafter(q[6], 1e30f) : 3.4e38f;
      auto rcp = [](float x) { return std::fabs(x) > 1e-18f ? 1.0f / x : std::copysign(1e18f, x); };
      const V3 inv{rcp(d.x), rcp(d.y), rcp(d.z)};
      int32_t stack_ref[128]; float stack_d[128]; int sp = 0;
      int32_t cur = root;
      int recs = 0; bool found = false;
      for (;;) {
        if (cur >= 0) {
          recs++;
          const Node& n = nodes[cur];
          float dl, dr;
          const bool hl = slab(n.box[0], o, inv, best, dl), hr = slab(n.box[1], o, inv, best, dr);
          if (hl && hr) {
            const bool left_first = !(dr < dl);
            stack_ref[sp] = left_first ? n.child[1] : n.child[0]; stack_d[sp] = left_first ? dr : dl; sp++;
            cur = left_first ? n.child[0] : n.child[1];
            continue;
          }
          if (hl) { cur = n.child[0]; continue; }
          if (hr) { cur = n.child[1]; continue; }
        } else {
          const int code = ~cur, first = code >> 3, count = (code & 7) + 1;
          bool stop = false;
          for (int i = 0; i < count && !stop; i++) {
            local.tris++;
            float tt;
            if (hit_tri(tris[order[first + i]], o, d, tt) && tt < best) { found = true; if (any) stop = true; else best = tt; }
          }
          if (stop) break;
        }
        cur = INT32_MIN;
        while (sp > 0) { sp--; if (!(stack_d[sp] >= best)) { cur = stack_ref[sp]; break; } }
        if (cur == INT32_MIN) break;
      }
      per_ray[(size_t)r] = recs;
      local.records += recs; local.rays++; local.hits += found; local.max_records = std::max(local.max_records, recs);
    }
#pragma omp critical
    { c.records += local.records; c.tris += local.tris; c.hits += local.hits; c.rays += local.rays; c.max_records = std::max(c.max_records, local.max_records); }
  }
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: bvh_study triangles.bin rays.bin\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  int64_t n = 0;
  if (!f || std::fread(&n, 8, 1, f) != 1) return 1;
  tris.resize((size_t)n);
  if (std::fread(tris.data(), sizeof(Tri), (size_t)n, f) != (size_t)n) return 1;
  std::fclose(f);
  f = std::fopen(argv[2], "rb");
  int64_t m = 0;
  if (!f || std::fread(&m, 8, 1, f) != 1) return 1;
  std::vector<float> rays((size_t)m * 7);
  if (std::fread(rays.data(), 28, (size_t)m, f) != (size_t)m) return 1;
  std::fclose(f);
  tbox.resize((size_t)n); cen.resize((size_t)n);
  Box all;
  for (int64_t i = 0; i < n; i++) {
    Box b; b.add(tris[i].v0); b.add(tris[i].v1); b.add(tris[i].v2);
    tbox[i] = b;
    cen[i] = {(tris[i].v0.x + tris[i].v1.x + tris[i].v2.x) / 3.0f, (tris[i].v0.y + tris[i].v1.y + tris[i].v2.y) / 3.0f, (tris[i].v0.z + tris[i].v1.z + tris[i].v2.z) / 3.0f};
    all.add(cen[i]);
  }
  const char* names[4] = {"lbvh", "sah", "median", "lbvh-sched"};
  for (int how = 0; how < 4; how++) {
    order.resize((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    nodes.clear();
    if (how == LBVH) {
      const float side = std::max(std::max(all.hi.x - all.lo.x, all.hi.y - all.lo.y), all.hi.z - all.lo.z);
      const double scale = side > 0 ? 2097151.0 / side : 0.0;
      std::vector<uint64_t> code((size_t)n);
      for (int64_t i = 0; i < n; i++)
        code[i] = spread21((uint64_t)((cen[i].x - all.lo.x) * scale)) << 2 | spread21((uint64_t)((cen[i].y - all.lo.y) * scale)) << 1 | spread21((uint64_t)((cen[i].z - all.lo.z) * scale));
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return code[a] < code[b]; });
      morton.resize((size_t)n);
      for (int64_t i = 0; i < n; i++) morton[i] = code[order[i]];
    }
    if (how == LBVH_SCHED) {
      // generalised Morton code: every key bit halves the axis whose remaining extent is largest (the reference builder's
      // longest-axis rule as a bit schedule), each axis quantised over its OWN extent
      float ext[3] = {all.hi.x - all.lo.x, all.hi.y - all.lo.y, all.hi.z - all.lo.z};
      const float full[3] = {ext[0], ext[1], ext[2]};
      int used[3] = {0, 0, 0}, sched[63];
      for (int b = 0; b < 63; b++) {
        int ax = -1;
        for (int a = 0; a < 3; a++) if (used[a] < 21 && (ax < 0 || ext[a] > ext[ax])) ax = a;
        sched[b] = ax; used[ax]++; ext[ax] *= 0.5f;
      }
      std::vector<uint64_t> code((size_t)n);
      for (int64_t i = 0; i < n; i++) {
        const float c[3] = {cen[i].x - all.lo.x, cen[i].y - all.lo.y, cen[i].z - all.lo.z};
        uint32_t q[3]; int taken[3] = {0, 0, 0};
        for (int a = 0; a < 3; a++) q[a] = full[a] > 0 ? (uint32_t)std::min(2097151.0, (double)c[a] / full[a] * 2097151.0) : 0;
        uint64_t k = 0;
        for (int b = 0; b < 63; b++) { const int a = sched[b]; k = (k << 1) | ((q[a] >> (20 - taken[a])) & 1u); taken[a]++; }
        code[i] = k;
      }
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return code[a] < code[b]; });
      morton.resize((size_t)n);
      for (int64_t i = 0; i < n; i++) morton[i] = code[order[i]];
    }
    const int32_t root = build((Builder)how, 0, (int)n, 62);
    double sah = 0.0;  // surface-area cost of the tree relative to the root box
    Box rb = range_box(0, (int)n);
    for (const Node& nd : nodes) sah += nd.box[0].area() + nd.box[1].area();
    Counts c;
    per_ray.assign((size_t)m, 0);
    trace(root, rays.data(), m, c);
    // lock-step bound: 32 consecutive rays share a warp; without refill the warp runs as long as its longest ray
    double sum_max = 0.0; long long groups = 0;
    for (long long g = 0; g + 32 <= m; g += 32) { int mx = 0; for (int k = 0; k < 32; k++) mx = std::max(mx, per_ray[(size_t)(g + k)]); sum_max += mx; groups++; }
    const double lockstep = groups ? (c.records / c.rays) / (sum_max / groups) : 0.0;
    std::printf("%-6s nodes %zu  sum(child area)/root area %.1f | rays %lld: records/ray %.2f, triangles/ray %.2f, hit fraction %.3f, longest ray %d records, mean/max over 32 consecutive rays %.2f\n",
                names[how], nodes.size(), sah / rb.area(), c.rays, c.records / c.rays, c.tris / c.rays, c.hits / c.rays, c.max_records, lockstep);
  }
  return 0;
}
