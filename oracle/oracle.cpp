// oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference renderer of mpoboas/cosig-raytracing, used only as the checker by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under cosig-raytracing_b200/ may
// link, import or execute it.
//
// PARITY UNPINNED: the reference ships no tests, golden images or expected ids and cannot be executed
// here (Unity C# + HLSL, no toolchain).  This file restates, function by function, what the reference's sources do; Unity's
// closed-source maths (Matrix4x4 / Quaternion / HLSL intrinsics) is fixed to the definitions in SURVEY.md Appendix D.
// Arithmetic spec: IEEE binary32, every + - * / sqrt individually rounded (build with -ffp-contract=off, no fast-math),
// evaluation order exactly as written below.  Paths cited are relative to the reference repository root.
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC oracle/oracle.cpp -o oracle/liboracle.so

#include "../include/rtb.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------------------------------
// FP32 vector helpers (HLSL intrinsics as fixed by SURVEY App. D)
// ---------------------------------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
static inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
static inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float fmin_(float a, float b) { return (a < b || b != b) ? a : b; }  // returns the non-NaN operand
static inline float fmax_(float a, float b) { return (a > b || b != b) ? a : b; }
static inline V3 hlsl_normalize(V3 v) { float r = 1.0f / sqrtf(dot(v, v)); return v * r; }
static inline float hlsl_length(V3 v) { return sqrtf(dot(v, v)); }
static inline V3 hlsl_reflect(V3 i, V3 n) { float k = 2.0f * dot(n, i); return i - k * n; }
static inline float fracf_(float x) { return x - floorf(x); }
static inline float pow32(float x) { x = x * x; x = x * x; x = x * x; x = x * x; x = x * x; return x; }

static const float kInfinity = 3.402823466e+38f;  // BVHRayTracing.compute:101 (FLT_MAX, not IEEE inf)
static const float kEpsilon = 1e-4f;              // BVHRayTracing.compute:102
static const float kOffset = 1e-4f * 100.0f;      // "Epsilon * 100.0", :396,442,447,454

// Deterministic sin/cos for RandomUnitVector: quadrant reduction + Taylor/Horner in FP32.  a in [0, 2pi].
static inline void det_sincos(float a, float* s_out, float* c_out) {
  int k = (int)floorf(a * 0.63661975f);
  float r = a - (float)k * 1.5707964f;
  float r2 = r * r;
  float s = r * (1.0f + r2 * (-1.6666667e-1f + r2 * (8.3333338e-3f + r2 * (-1.9841270e-4f + r2 * (2.7557319e-6f + r2 * -2.5052108e-8f)))));
  float c = 1.0f + r2 * (-0.5f + r2 * (4.1666668e-2f + r2 * (-1.3888889e-3f + r2 * (2.4801587e-5f + r2 * (-2.7557319e-7f + r2 * 2.0876757e-9f)))));
  switch (k & 3) {
    case 0: *s_out = s; *c_out = c; break;
    case 1: *s_out = c; *c_out = -s; break;
    case 2: *s_out = -s; *c_out = -c; break;
    default: *s_out = -c; *c_out = s; break;
  }
}

// BVHRayTracing.compute:108-113
static inline void Hash22(float px, float py, float* ox, float* oy) {
  float a = fracf_(px * .1031f), b = fracf_(py * .1030f), c = fracf_(px * .0973f);
  float d = dot(v3(a, b, c), v3(b + 33.33f, c + 33.33f, a + 33.33f));
  a = a + d; b = b + d; c = c + d;
  *ox = fracf_((a + b) * c);
  *oy = fracf_((a + c) * b);
}
// BVHRayTracing.compute:116-121
static inline V3 Hash33(V3 p) {
  p = v3(fracf_(p.x * .1031f), fracf_(p.y * .1030f), fracf_(p.z * .0973f));
  float d = dot(p, v3(p.y + 33.33f, p.x + 33.33f, p.z + 33.33f));
  p = v3(p.x + d, p.y + d, p.z + d);
  return v3(fracf_((p.x + p.y) * p.z), fracf_((p.x + p.x) * p.y), fracf_((p.y + p.x) * p.x));
}
// BVHRayTracing.compute:124-131
static inline V3 RandomUnitVector(V3 seed) {
  V3 h = Hash33(seed);
  float z = h.z * 2.0f - 1.0f;
  float a = h.x * 6.2831853f;
  float r = sqrtf(1.0f - z * z);
  float s, c;
  det_sincos(a, &s, &c);
  return v3(r * c, r * s, z);
}

// ---------------------------------------------------------------------------------------------------------------------
// Unity maths (closed source; definitions of SURVEY App. D).  Row-major storage m[r][c], column vectors.
// ---------------------------------------------------------------------------------------------------------------------
struct M4 { float m[4][4]; };
static M4 m4_identity() { M4 r; memset(&r, 0, sizeof r); r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1.0f; return r; }
static M4 m4_mul(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      r.m[i][j] = ((a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j]) + a.m[i][3] * b.m[3][j];
  return r;
}
static M4 m4_translate(float x, float y, float z) { M4 r = m4_identity(); r.m[0][3] = x; r.m[1][3] = y; r.m[2][3] = z; return r; }
static M4 m4_scale(float x, float y, float z) { M4 r = m4_identity(); r.m[0][0] = x; r.m[1][1] = y; r.m[2][2] = z; return r; }
struct Quat { float x, y, z, w; };
static const float kDeg2Rad = 0.0174532924f;  // Mathf.Deg2Rad
static Quat quat_angle_axis(float deg, float ax, float ay, float az) {
  float half = deg * kDeg2Rad * 0.5f;
  float s = (float)sin((double)half), c = (float)cos((double)half);
  return Quat{ax * s, ay * s, az * s, c};
}
static Quat quat_mul(Quat l, Quat r) {
  return Quat{l.w * r.x + l.x * r.w + l.y * r.z - l.z * r.y, l.w * r.y + l.y * r.w + l.z * r.x - l.x * r.z,
              l.w * r.z + l.z * r.w + l.x * r.y - l.y * r.x, l.w * r.w - l.x * r.x - l.y * r.y - l.z * r.z};
}
static M4 m4_rotate(Quat q) {
  float x = q.x * 2.0f, y = q.y * 2.0f, z = q.z * 2.0f;
  float xx = q.x * x, yy = q.y * y, zz = q.z * z, xy = q.x * y, xz = q.x * z, yz = q.y * z, wx = q.w * x, wy = q.w * y, wz = q.w * z;
  M4 r = m4_identity();
  r.m[0][0] = 1.0f - (yy + zz); r.m[1][0] = xy + wz;          r.m[2][0] = xz - wy;
  r.m[0][1] = xy - wz;          r.m[1][1] = 1.0f - (xx + zz); r.m[2][1] = yz + wx;
  r.m[0][2] = xz + wy;          r.m[1][2] = yz - wx;          r.m[2][2] = 1.0f - (xx + yy);
  return r;
}
// Quaternion.Euler(x,y,z): rotate about Z, then X, then Y  (q = qY * qX * qZ)
static Quat quat_euler(float xd, float yd, float zd) {
  Quat qx = quat_angle_axis(xd, 1, 0, 0), qy = quat_angle_axis(yd, 0, 1, 0), qz = quat_angle_axis(zd, 0, 0, 1);
  return quat_mul(quat_mul(qy, qx), qz);
}
// Matrix4x4.inverse: adjugate / determinant evaluated in double, rounded once to FP32; singular -> zero matrix.
static M4 m4_inverse(const M4& a) {
  double m[16];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i * 4 + j] = (double)a.m[i][j];
  double inv[16];
  inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
  M4 r;
  if (det == 0.0) { memset(&r, 0, sizeof r); return r; }
  double id = 1.0 / det;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = (float)(inv[i * 4 + j] * id);
  return r;
}
static M4 m4_transpose(const M4& a) { M4 r; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = a.m[j][i]; return r; }
static inline V3 mul_point3x4(const M4& m, V3 v) {
  return v3(((m.m[0][0] * v.x + m.m[0][1] * v.y) + m.m[0][2] * v.z) + m.m[0][3],
            ((m.m[1][0] * v.x + m.m[1][1] * v.y) + m.m[1][2] * v.z) + m.m[1][3],
            ((m.m[2][0] * v.x + m.m[2][1] * v.y) + m.m[2][2] * v.z) + m.m[2][3]);
}
static inline V3 mul_vector(const M4& m, V3 v) {
  return v3((m.m[0][0] * v.x + m.m[0][1] * v.y) + m.m[0][2] * v.z, (m.m[1][0] * v.x + m.m[1][1] * v.y) + m.m[1][2] * v.z,
            (m.m[2][0] * v.x + m.m[2][1] * v.y) + m.m[2][2] * v.z);
}
// Vector3.normalized
static inline V3 unity_normalized(V3 v) {
  float mag = sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
  if (mag > 1e-5f) return v3(v.x / mag, v.y / mag, v.z / mag);
  return v3(0, 0, 0);
}

// ---------------------------------------------------------------------------------------------------------------------
// Scene text loader — Assets/Services/SceneService.cs:26-334
// ---------------------------------------------------------------------------------------------------------------------
struct OwnedScene {
  rtb_scene_desc d;
  std::vector<int32_t> xoff; std::vector<rtb_xform_elem> xel;
  std::vector<int32_t> lxf; std::vector<float> lrgb;
  std::vector<rtb_material> mats; std::vector<rtb_mesh> meshes; std::vector<rtb_triangle> tris;
  std::vector<rtb_prim> spheres, boxes;
  void finish() {
    d.n_xforms = (int32_t)xoff.size() - 1; d.xform_offsets = xoff.data(); d.xform_elems = xel.data();
    d.n_lights = (int32_t)lxf.size(); d.light_xforms = lxf.data(); d.light_rgb = lrgb.data();
    d.n_materials = (int32_t)mats.size(); d.materials = mats.data();
    d.n_meshes = (int32_t)meshes.size(); d.meshes = meshes.data();
    d.n_triangles = (int64_t)tris.size(); d.triangles = tris.data();
    d.n_spheres = (int32_t)spheres.size(); d.spheres = spheres.data();
    d.n_boxes = (int32_t)boxes.size(); d.boxes = boxes.data();
  }
  void copy_from(const rtb_scene_desc& s) {
    d = s;
    xoff.assign(s.xform_offsets, s.xform_offsets + s.n_xforms + 1);
    if (s.n_xforms == 0 && xoff.empty()) xoff.push_back(0);
    xel.assign(s.xform_elems, s.xform_elems + (s.n_xforms ? s.xform_offsets[s.n_xforms] : 0));
    lxf.assign(s.light_xforms, s.light_xforms + s.n_lights);
    lrgb.assign(s.light_rgb, s.light_rgb + (s.light_rgb ? 3 * s.n_lights : 0));
    mats.assign(s.materials, s.materials + s.n_materials);
    meshes.assign(s.meshes, s.meshes + s.n_meshes);
    tris.assign(s.triangles, s.triangles + s.n_triangles);
    spheres.assign(s.spheres, s.spheres + s.n_spheres);
    boxes.assign(s.boxes, s.boxes + s.n_boxes);
    finish();
  }
};

struct ParseError { std::string msg; };

// Clean(): strip "//" comment, trim (SceneService.cs:258-267)
static std::string clean_line(const std::string& in) {
  std::string s = in;
  size_t c = s.find("//");
  if (c != std::string::npos) s.erase(c);
  size_t b = 0, e = s.size();
  while (b < e && isspace((unsigned char)s[b])) b++;
  while (e > b && isspace((unsigned char)s[e - 1])) e--;
  return s.substr(b, e - b);
}
static bool is_segment(const std::string& line, const char* name) {  // :272-275 (ordinal, ignore case)
  size_t n = strlen(name);
  if (line.size() != n) return false;
  for (size_t i = 0; i < n; i++) if (tolower((unsigned char)line[i]) != tolower((unsigned char)name[i])) return false;
  return true;
}
static double parse_double(const std::string& tok) {  // :308-311
  const char* p = tok.c_str();
  char* end = nullptr;
  double v = strtod(p, &end);
  while (end && *end && isspace((unsigned char)*end)) end++;
  if (end == p || (end && *end)) throw ParseError{"bad number '" + tok + "'"};
  return v;
}
static std::vector<std::string> split_ws(const std::string& s) {  // Split(' ', '\t', RemoveEmptyEntries)
  std::vector<std::string> out;
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) i++;
    size_t j = i;
    while (j < s.size() && s[j] != ' ' && s[j] != '\t') j++;
    if (j > i) out.push_back(s.substr(i, j - i));
    i = j;
  }
  return out;
}
static std::vector<double> parse_floats(const std::string& line) {  // :316-323
  std::vector<double> v;
  for (auto& t : split_ws(line)) v.push_back(parse_double(t));
  return v;
}
struct Lines {
  std::vector<std::string> l;
  size_t i = 0;
  const std::string& at(size_t k) const { if (k >= l.size()) throw ParseError{"unexpected end of file"}; return l[k]; }
  void expect_brace(const char*) {  // :280-301: skip blank lines, then consume one line whatever it is
    while (i < l.size() && clean_line(l[i]).empty()) i++;
    i++;
  }
};
static double need(const std::vector<double>& v, size_t k) { if (k >= v.size()) throw ParseError{"too few numbers on line"}; return v[k]; }

static void parse_scene_text(const std::string& text, OwnedScene& sc) {
  Lines L;
  {
    size_t p = 0;
    while (p <= text.size()) {  // File.ReadAllLines: split on \n, \r\n, \r
      size_t q = text.find_first_of("\r\n", p);
      if (q == std::string::npos) { if (p < text.size()) L.l.push_back(text.substr(p)); break; }
      L.l.push_back(text.substr(p, q - p));
      p = (text[q] == '\r' && q + 1 < text.size() && text[q + 1] == '\n') ? q + 2 : q + 1;
    }
  }
  memset(&sc.d, 0, sizeof sc.d);
  sc.xoff.assign(1, 0);
  size_t& i = L.i;
  while (i < L.l.size()) {
    std::string line = clean_line(L.l[i]);
    i++;
    if (line.empty()) continue;
    if (is_segment(line, "Image")) {  // :45-64
      L.expect_brace("{");
      auto res = parse_floats(clean_line(L.at(i++)));
      auto bg = parse_floats(clean_line(L.at(i++)));
      L.expect_brace("}");
      sc.d.has_image = 1;
      sc.d.image_w = (int)need(res, 0); sc.d.image_h = (int)need(res, 1);
      sc.d.bg[0] = (float)need(bg, 0); sc.d.bg[1] = (float)need(bg, 1); sc.d.bg[2] = (float)need(bg, 2);
    } else if (is_segment(line, "Transformation")) {  // :65-116
      L.expect_brace("{");
      while (i < L.l.size()) {
        std::string inner = clean_line(L.l[i]);
        if (inner == "}") { i++; break; }
        if (inner.empty()) { i++; continue; }
        auto tok = split_ws(inner);
        if (tok.empty()) { i++; continue; }
        auto arg = [&](size_t k) { if (k >= tok.size()) throw ParseError{"transform element needs more arguments"}; return (float)parse_double(tok[k]); };
        rtb_xform_elem e; memset(&e, 0, sizeof e);
        bool ok = true;
        if (tok[0] == "T") { e.type = RTB_XF_T; e.x = arg(1); e.y = arg(2); e.z = arg(3); }
        else if (tok[0] == "S") { e.type = RTB_XF_S; e.x = arg(1); e.y = arg(2); e.z = arg(3); }
        else if (tok[0] == "Rx") { e.type = RTB_XF_RX; e.angle_deg = arg(1); }
        else if (tok[0] == "Ry") { e.type = RTB_XF_RY; e.angle_deg = arg(1); }
        else if (tok[0] == "Rz") { e.type = RTB_XF_RZ; e.angle_deg = arg(1); }
        else ok = false;
        if (ok) sc.xel.push_back(e);
        i++;
      }
      sc.xoff.push_back((int32_t)sc.xel.size());
    } else if (is_segment(line, "Camera")) {  // :117-138
      L.expect_brace("{");
      int t = (int)parse_double(clean_line(L.at(i++)));
      double dist = parse_double(clean_line(L.at(i++)));
      double fov = parse_double(clean_line(L.at(i++)));
      L.expect_brace("}");
      sc.d.has_camera = 1; sc.d.cam_xform = t; sc.d.cam_distance = (float)dist; sc.d.cam_vfov_deg = (float)fov;
    } else if (is_segment(line, "Light")) {  // :139-157
      L.expect_brace("{");
      int t = (int)parse_double(clean_line(L.at(i++)));
      auto rgb = parse_floats(clean_line(L.at(i++)));
      L.expect_brace("}");
      sc.lxf.push_back(t);
      for (int k = 0; k < 3; k++) sc.lrgb.push_back((float)need(rgb, k));
    } else if (is_segment(line, "Material")) {  // :158-180
      L.expect_brace("{");
      auto col = parse_floats(clean_line(L.at(i++)));
      auto co = parse_floats(clean_line(L.at(i++)));
      L.expect_brace("}");
      rtb_material m{(float)need(col, 0), (float)need(col, 1), (float)need(col, 2), (float)need(co, 0), (float)need(co, 1),
                     (float)need(co, 2), (float)need(co, 3), (float)need(co, 4)};
      sc.mats.push_back(m);
    } else if (is_segment(line, "Triangles")) {  // :181-210
      L.expect_brace("{");
      rtb_mesh mesh; memset(&mesh, 0, sizeof mesh);
      mesh.xform = (int)parse_double(clean_line(L.at(i++)));
      mesh.first_tri = (int64_t)sc.tris.size();
      while (i < L.l.size()) {
        std::string inner = clean_line(L.l[i]);
        if (inner == "}") { i++; break; }
        if (inner.empty()) { i++; continue; }
        rtb_triangle t;
        t.material = (int)parse_double(inner);
        float* dst[3] = {t.v0, t.v1, t.v2};
        for (int k = 0; k < 3; k++) {
          auto v = parse_floats(clean_line(L.at(i + 1 + k)));
          for (int c = 0; c < 3; c++) dst[k][c] = (float)need(v, c);
        }
        sc.tris.push_back(t);
        i += 4;
      }
      mesh.n_tris = (int64_t)sc.tris.size() - mesh.first_tri;
      sc.meshes.push_back(mesh);
    } else if (is_segment(line, "Sphere") || is_segment(line, "Box")) {  // :211-238
      bool sph = is_segment(line, "Sphere");
      L.expect_brace("{");
      int t = (int)parse_double(clean_line(L.at(i++)));
      int m = (int)parse_double(clean_line(L.at(i++)));
      L.expect_brace("}");
      (sph ? sc.spheres : sc.boxes).push_back(rtb_prim{t, m});
    }
  }
  sc.finish();
}

// ---------------------------------------------------------------------------------------------------------------------
// Geometry flattening — Assets/Services/SceneGeometryConverter.cs
// ---------------------------------------------------------------------------------------------------------------------
struct Tri { V3 v0, v1, v2, n0, n1, n2, center; int material; };  // GPUTriangle, BVHBuilder.cs:10-21

// BuildMatrix, SceneGeometryConverter.cs:83-114 == BuildComposite, RayTracer.cs:410-437
static M4 build_matrix(const rtb_scene_desc& s, int index) {
  if (index < 0 || index >= s.n_xforms) return m4_identity();
  M4 M = m4_identity();
  for (int k = s.xform_offsets[index]; k < s.xform_offsets[index + 1]; k++) {
    const rtb_xform_elem& e = s.xform_elems[k];
    M4 t = m4_identity();
    switch (e.type) {
      case RTB_XF_T: t = m4_translate(e.x, e.y, e.z); break;
      case RTB_XF_S: t = m4_scale(e.x, e.y, e.z); break;
      case RTB_XF_RX: t = m4_rotate(quat_angle_axis(e.angle_deg, 1, 0, 0)); break;
      case RTB_XF_RY: t = m4_rotate(quat_angle_axis(e.angle_deg, 0, 1, 0)); break;
      case RTB_XF_RZ: t = m4_rotate(quat_angle_axis(e.angle_deg, 0, 0, 1)); break;
    }
    M = m4_mul(M, t);
  }
  return M;
}
// CreateGPUTriangleWithNormals :66-77
static Tri make_tri_n(V3 a, V3 b, V3 c, V3 na, V3 nb, V3 nc, int mat) {
  Tri t; t.v0 = a; t.v1 = b; t.v2 = c; t.n0 = na; t.n1 = nb; t.n2 = nc;
  V3 s = (a + b) + c;
  t.center = v3(s.x / 3.0f, s.y / 3.0f, s.z / 3.0f);
  t.material = mat;
  return t;
}
// CreateGPUTriangle :56-60
static Tri make_tri(V3 a, V3 b, V3 c, int mat) {
  V3 fn = unity_normalized(cross(b - a, c - a));
  return make_tri_n(a, b, c, fn, fn, fn, mat);
}
// AddCube :120-155
static void add_cube(std::vector<Tri>& out, const M4& m, int mat) {
  V3 v[8] = {v3(-0.5f, -0.5f, -0.5f), v3(0.5f, -0.5f, -0.5f), v3(0.5f, 0.5f, -0.5f), v3(-0.5f, 0.5f, -0.5f),
             v3(-0.5f, -0.5f, 0.5f),  v3(0.5f, -0.5f, 0.5f),  v3(0.5f, 0.5f, 0.5f),  v3(-0.5f, 0.5f, 0.5f)};
  for (auto& p : v) p = mul_point3x4(m, p);
  static const int idx[12][3] = {{0, 2, 1}, {0, 3, 2}, {5, 7, 6}, {5, 4, 7}, {3, 6, 2}, {3, 7, 6},
                                 {4, 1, 5}, {4, 0, 1}, {4, 3, 7}, {4, 0, 3}, {1, 6, 2}, {1, 5, 6}};
  for (auto& f : idx) out.push_back(make_tri(v[f[0]], v[f[1]], v[f[2]], mat));
}
// AddSmoothTri :245-264
static void add_smooth_tri(std::vector<Tri>& out, const M4& m, const M4& normalMat, V3 a, V3 b, V3 c, int mat) {
  V3 na = unity_normalized(a), nb = unity_normalized(b), nc = unity_normalized(c);
  V3 va = mul_point3x4(m, a), vb = mul_point3x4(m, b), vc = mul_point3x4(m, c);
  V3 tna = unity_normalized(mul_vector(normalMat, na));
  V3 tnb = unity_normalized(mul_vector(normalMat, nb));
  V3 tnc = unity_normalized(mul_vector(normalMat, nc));
  out.push_back(make_tri_n(va, vb, vc, tna, tnb, tnc, mat));
}
// AddSphere :161-230
static void add_sphere(std::vector<Tri>& out, const M4& m, int mat) {
  const int nbLong = 24, nbLat = 16;
  std::vector<V3> sv((nbLong + 1) * nbLat + 2);
  const float pi = 3.14159274f, twopi = pi * 2.0f;
  sv[0] = v3(0, 1, 0);
  for (int lat = 0; lat < nbLat; lat++) {
    float a1 = pi * (float)(lat + 1) / (float)(nbLat + 1);
    float sin1 = (float)sin((double)a1), cos1 = (float)cos((double)a1);
    for (int lon = 0; lon <= nbLong; lon++) {
      float a2 = twopi * (float)(lon == nbLong ? 0 : lon) / (float)nbLong;
      float sin2 = (float)sin((double)a2), cos2 = (float)cos((double)a2);
      sv[lon + lat * (nbLong + 1) + 1] = v3(sin1 * cos2, cos1, sin1 * sin2) * 1.0f;
    }
  }
  sv[sv.size() - 1] = v3(0, -1, 0);
  M4 normalMat = m4_transpose(m4_inverse(m));  // :258 (recomputed per triangle in the reference; same value)
  for (int lon = 0; lon < nbLong; lon++) add_smooth_tri(out, m, normalMat, sv[0], sv[lon + 2], sv[lon + 1], mat);
  for (int lat = 0; lat < nbLat - 1; lat++)
    for (int lon = 0; lon < nbLong; lon++) {
      int cur = lon + lat * (nbLong + 1) + 1, next = cur + 1, below = cur + (nbLong + 1), belowNext = below + 1;
      add_smooth_tri(out, m, normalMat, sv[cur], sv[below], sv[next], mat);
      add_smooth_tri(out, m, normalMat, sv[next], sv[below], sv[belowNext], mat);
    }
  int last = (int)sv.size() - 1;
  for (int lon = 0; lon < nbLong; lon++)
    add_smooth_tri(out, m, normalMat, sv[last], sv[last - (nbLong + 1) + lon], sv[last - (nbLong + 1) + lon + 1], mat);
}
// ExtractTriangles :18-51 — meshes, then boxes, then spheres
static std::vector<Tri> extract_triangles(const rtb_scene_desc& s, bool meshes_only = false) {
  std::vector<Tri> out;
  out.reserve((size_t)s.n_triangles + 12 * (size_t)s.n_boxes + 768 * (size_t)s.n_spheres);
  for (int mi = 0; mi < s.n_meshes; mi++) {
    M4 m = build_matrix(s, s.meshes[mi].xform);
    for (int64_t k = 0; k < s.meshes[mi].n_tris; k++) {
      const rtb_triangle& t = s.triangles[s.meshes[mi].first_tri + k];
      V3 a = mul_point3x4(m, v3(t.v0[0], t.v0[1], t.v0[2]));
      V3 b = mul_point3x4(m, v3(t.v1[0], t.v1[1], t.v1[2]));
      V3 c = mul_point3x4(m, v3(t.v2[0], t.v2[1], t.v2[2]));
      out.push_back(make_tri(a, b, c, t.material));
    }
  }
  if (meshes_only) return out;  // analytic mode: boxes and spheres stay primitives (struct Analytic below)
  for (int i = 0; i < s.n_boxes; i++) add_cube(out, build_matrix(s, s.boxes[i].xform), s.boxes[i].material);
  for (int i = 0; i < s.n_spheres; i++) add_sphere(out, build_matrix(s, s.spheres[i].xform), s.spheres[i].material);
  return out;
}

// ---------------------------------------------------------------------------------------------------------------------
// BVH build — Assets/Services/BVH/BVHBuilder.cs:76-238, AABB.cs:23-39,72
// ---------------------------------------------------------------------------------------------------------------------
struct GpuNode { V3 mn; int leftOrFirst; V3 mx; int count; };  // GPUBVHNode, BVHBuilder.cs:27-34
struct BNode { V3 mn, mx; int left = -1, right = -1; int start = 0, count = 0; };

struct Builder {
  const std::vector<Tri>& tris;
  std::vector<int> idx;
  std::vector<BNode> pool;
  explicit Builder(const std::vector<Tri>& t) : tris(t) { idx.resize(t.size()); for (size_t i = 0; i < t.size(); i++) idx[i] = (int)i; }
  static float comp(V3 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
  int partition(int start, int count, int axis, float pivot) {  // :160-183
    int i = start, j = start + count - 1;
    while (i <= j) {
      float c = comp(tris[idx[i]].center, axis);
      if (c < pivot) i++;
      else { std::swap(idx[i], idx[j]); j--; }
    }
    return i;
  }
  int build(int start, int count) {  // BuildRecursive :103-153
    int me = (int)pool.size();
    pool.emplace_back();
    V3 mn = v3(INFINITY, INFINITY, INFINITY), mx = v3(-INFINITY, -INFINITY, -INFINITY);  // AABB.Empty
    for (int k = 0; k < count; k++) {
      const Tri& t = tris[idx[start + k]];
      const V3 p[3] = {t.v0, t.v1, t.v2};
      for (auto& v : p) {  // Encapsulate: Vector3.Min / Max (Mathf.Min per component)
        mn = v3(v.x < mn.x ? v.x : mn.x, v.y < mn.y ? v.y : mn.y, v.z < mn.z ? v.z : mn.z);
        mx = v3(v.x > mx.x ? v.x : mx.x, v.y > mx.y ? v.y : mx.y, v.z > mx.z ? v.z : mx.z);
      }
    }
    pool[me].mn = mn; pool[me].mx = mx; pool[me].start = start; pool[me].count = count;
    if (count <= 4) return me;
    V3 size = mx - mn;
    int axis = 0;
    if (size.y > size.x) axis = 1;
    if (size.z > comp(size, axis)) axis = 2;
    V3 center = (mn + mx) * 0.5f;
    float split = comp(center, axis);
    int mid = partition(start, count, axis, split);
    if (mid == start || mid == start + count) return me;
    int l = build(start, mid - start);
    int r = build(mid, (start + count) - mid);
    pool[me].left = l; pool[me].right = r; pool[me].count = 0;
    return me;
  }
};

// Analytic primitive mode (RTB_PRIM_ANALYTIC): spheres and boxes keep the semantics of the reference's (never instantiated)
// SphereInstance / BoxInstance, Assets/Services/BVH/HittableObjects.cs:6-224 — unit sphere / unit cube in object space, ray
// taken to object space with the direction re-normalised, world t = |pWS - origin|, normal = normalize((M^-1)^T n).
struct Analytic { int kind; /* 1 sphere, 2 box */ M4 M, W; int material; };
struct AnalyticHit { float tOS; int face; };  // what shading needs besides t: object-space t; box face 1..6 = -x +x -y +y -z +z

struct Scene {
  OwnedScene owned;
  bool analytic = false;
  std::vector<Analytic> prims;      // analytic mode: boxes then spheres (emission order continues after the mesh triangles)
  std::vector<Tri> tris_emit;       // emission order (prim_id = index here)
  std::vector<Tri> tris;            // BVH leaf order (what the reference uploads)
  std::vector<int> orig;            // tris[i] == tris_emit[orig[i]]
  std::vector<GpuNode> nodes;
  std::vector<rtb_material> mats;   // SetupMaterialBuffer, RayTracer.cs:455-499
  int max_leaf = 0;
  // Checker-only (see LeafAccel below): big leaves get a private sub-tree; node index -> entry of leaf_accel, or -1
  std::vector<int> accel_of_node;
  std::vector<struct LeafAccel> leaf_accel;
  // Checker-only (orc_set_exact_closest): TraverseBVH culls a node only when its slab entry lies clearly BEYOND the best t so far,
  // which makes it return the exact closest hit (what a scan of all triangles returns) instead of the reference's answer.
  bool exact_closest = false;
};

// ---------------------------------------------------------------------------------------------------------------------
// CHECKER-ONLY ACCELERATOR — not part of the reference and not a change of its semantics.  BVHBuilder.cs makes a leaf of ANY
// size when its partition fails (:142-145: eval_scene has a 578-triangle leaf, the C3 sphere grid two leaves of 98 310), and
// TraverseBVH tests every triangle of a leaf in order (:251-256), which makes the restatement too slow to check full C3
// frames.  A leaf with more than g_leaf_accel_min triangles therefore gets a private balanced sub-tree over ITS triangles,
// and "test all of the leaf's triangles in order, keep the first of the closest" (the strict '<' of :179) is evaluated as
// "closest t below the incoming bound; among equal t the smallest leaf position" — the same triangle, the same t/u/v bits,
// because every (ray, triangle) test is the unchanged IntersectTriangle.  The sub-tree's boxes are padded far beyond any
// rounding of its slab test (1e-3 + 1e-5 |coordinate|), so it never skips a triangle IntersectTriangle would accept; the
// reference tree above the leaf, its node order and its culling are untouched.  Counters report the leaf's full triangle
// count, as the linear scan would.  tests/test_host_cpu.py compares accelerated and plain renders bit for bit.
// ---------------------------------------------------------------------------------------------------------------------
static int g_leaf_accel_min = 32;  // 0 = never accelerate
struct SubNode { V3 mn, mx; int left = -1, right = -1, first = 0, count = 0; };
struct LeafAccel {
  std::vector<SubNode> nodes;
  std::vector<int> order;  // leaf-order triangle indices, grouped by sub-tree leaf
};

static int build_sub(LeafAccel& a, const std::vector<Tri>& tris, int first, int count) {
  const int me = (int)a.nodes.size();
  a.nodes.emplace_back();
  V3 mn = v3(INFINITY, INFINITY, INFINITY), mx = v3(-INFINITY, -INFINITY, -INFINITY), cmn = mn, cmx = mx;
  for (int k = 0; k < count; k++) {
    const Tri& t = tris[a.order[first + k]];
    const V3 p[3] = {t.v0, t.v1, t.v2};
    for (auto& v : p) {
      mn = v3(std::min(mn.x, v.x), std::min(mn.y, v.y), std::min(mn.z, v.z));
      mx = v3(std::max(mx.x, v.x), std::max(mx.y, v.y), std::max(mx.z, v.z));
    }
    cmn = v3(std::min(cmn.x, t.center.x), std::min(cmn.y, t.center.y), std::min(cmn.z, t.center.z));
    cmx = v3(std::max(cmx.x, t.center.x), std::max(cmx.y, t.center.y), std::max(cmx.z, t.center.z));
  }
  auto pad = [](float v) { return 1e-3f + 1e-5f * fabsf(v); };
  a.nodes[me].mn = v3(mn.x - pad(mn.x), mn.y - pad(mn.y), mn.z - pad(mn.z));
  a.nodes[me].mx = v3(mx.x + pad(mx.x), mx.y + pad(mx.y), mx.z + pad(mx.z));
  a.nodes[me].first = first; a.nodes[me].count = count;
  if (count <= 4) return me;
  const V3 ext = cmx - cmn;
  int axis = 0;
  if (ext.y > ext.x) axis = 1;
  if (ext.z > (axis == 0 ? ext.x : ext.y)) axis = 2;
  auto key = [&](int i) { const V3& c = tris[i].center; return axis == 0 ? c.x : (axis == 1 ? c.y : c.z); };
  const int half = count / 2;  // object median: always balanced, whatever the geometry
  std::nth_element(a.order.begin() + first, a.order.begin() + first + half, a.order.begin() + first + count,
                   [&](int x, int y) { const float kx = key(x), ky = key(y); return kx < ky || (kx == ky && x < y); });
  const int l = build_sub(a, tris, first, half);
  const int r = build_sub(a, tris, first + half, count - half);
  a.nodes[me].left = l; a.nodes[me].right = r; a.nodes[me].count = 0;
  return me;
}

static void build_scene(Scene& sc) {
  const rtb_scene_desc& d = sc.owned.d;
  sc.tris_emit = extract_triangles(d, sc.analytic);
  sc.prims.clear();
  if (sc.analytic) {
    for (int i = 0; i < d.n_boxes; i++) { M4 M = build_matrix(d, d.boxes[i].xform); sc.prims.push_back(Analytic{2, M, m4_inverse(M), d.boxes[i].material}); }
    for (int i = 0; i < d.n_spheres; i++) { M4 M = build_matrix(d, d.spheres[i].xform); sc.prims.push_back(Analytic{1, M, m4_inverse(M), d.spheres[i].material}); }
  }
  sc.nodes.clear(); sc.tris.clear(); sc.orig.clear(); sc.max_leaf = 0;
  if (!sc.tris_emit.empty()) {
    Builder b(sc.tris_emit);
    int root = b.build(0, (int)sc.tris_emit.size());
    // Flatten :189-238 — BFS, sibling pairs adjacent, triangles re-emitted in leaf visit order
    std::queue<std::pair<int, int>> q;
    sc.nodes.emplace_back();
    q.push({root, 0});
    while (!q.empty()) {
      auto [n, at] = q.front();
      q.pop();
      const BNode& bn = b.pool[n];
      GpuNode g; g.mn = bn.mn; g.mx = bn.mx;
      if (bn.count > 0) {
        g.count = bn.count; g.leftOrFirst = (int)sc.tris.size();
        for (int k = 0; k < bn.count; k++) { sc.tris.push_back(sc.tris_emit[b.idx[bn.start + k]]); sc.orig.push_back(b.idx[bn.start + k]); }
        sc.max_leaf = std::max(sc.max_leaf, bn.count);
      } else {
        g.count = 0;
        int l = (int)sc.nodes.size();
        sc.nodes.emplace_back(); sc.nodes.emplace_back();
        g.leftOrFirst = l;
        q.push({bn.left, l}); q.push({bn.right, l + 1});
      }
      sc.nodes[at] = g;
    }
  }
  sc.accel_of_node.assign(sc.nodes.size(), -1);
  sc.leaf_accel.clear();
  if (g_leaf_accel_min > 0)
    for (size_t ni = 0; ni < sc.nodes.size(); ni++) {
      const GpuNode& n = sc.nodes[ni];
      if (n.count <= g_leaf_accel_min) continue;
      sc.accel_of_node[ni] = (int)sc.leaf_accel.size();
      sc.leaf_accel.emplace_back();
      LeafAccel& a = sc.leaf_accel.back();
      a.order.resize((size_t)n.count);
      for (int k = 0; k < n.count; k++) a.order[(size_t)k] = n.leftOrFirst + k;
      build_sub(a, sc.tris, 0, n.count);
    }
  sc.mats.clear();
  if (d.n_materials == 0) sc.mats.push_back(rtb_material{1, 1, 1, 0.1f, 0.7f, 0, 0, 1.0f});
  else sc.mats.assign(d.materials, d.materials + d.n_materials);
}

// ---------------------------------------------------------------------------------------------------------------------
// Kernel — Assets/Shaders/BVHRayTracing.compute
// ---------------------------------------------------------------------------------------------------------------------
struct Ray { V3 o, d, inv; };
static inline Ray CreateRay(V3 o, V3 d) { return Ray{o, d, v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z)}; }  // :137-144
// HitRecord :22-29 (position/normal derived on demand).  tri >= 0: triangle (leaf order), u,v barycentrics.
// prim >= 0: analytic primitive, u = object-space t, v = box face code.
struct Hit { bool hit; float t; int tri; float u, v; int prim = -1; };

struct Counters {
  int64_t rays_primary = 0, rays_continuation = 0, rays_shadow = 0, nodes_visited = 0, tris_tested = 0, closest_hits = 0, primary_hits = 0;
  int64_t nodes_visited_shadow = 0, tris_tested_shadow = 0;  // the share of nodes_visited / tris_tested spent in shadow queries
  int max_stack = 0;
};

// IntersectAABB :199-216
static inline float IntersectAABB(const Ray& r, V3 mn, V3 mx) {
  V3 t0 = (mn - r.o) * r.inv, t1 = (mx - r.o) * r.inv;
  V3 tmin = v3(fmin_(t0.x, t1.x), fmin_(t0.y, t1.y), fmin_(t0.z, t1.z));
  V3 tmax = v3(fmax_(t0.x, t1.x), fmax_(t0.y, t1.y), fmax_(t0.z, t1.z));
  float dstA = fmax_(fmax_(tmin.x, tmin.y), tmin.z);
  float dstB = fmin_(fmin_(tmax.x, tmax.y), tmax.z);
  if (dstA > dstB || dstB < 0) return kInfinity;
  return dstA;
}
// IntersectTriangle :153-190
static inline void IntersectTriangle(const Ray& r, const Tri& tri, int index, Hit& best) {
  V3 e1 = tri.v1 - tri.v0, e2 = tri.v2 - tri.v0;
  V3 pvec = cross(r.d, e2);
  float det = dot(e1, pvec);
  if (fabsf(det) < kEpsilon) return;
  float invDet = 1.0f / det;
  V3 tvec = r.o - tri.v0;
  float u = dot(tvec, pvec) * invDet;
  if (u < 0.0f || u > 1.0f) return;
  V3 qvec = cross(tvec, e1);
  float v = dot(r.d, qvec) * invDet;
  if (v < 0.0f || u + v > 1.0f) return;
  float t = dot(e2, qvec) * invDet;
  if (t > kEpsilon && t < best.t) { best.hit = true; best.t = t; best.tri = index; best.u = u; best.v = v; }
}
// Checker-only: the leaf loop of :251-256 over a big leaf, through its sub-tree (see LeafAccel).  Same result as
// `for (i = first; i < first + count; i++) IntersectTriangle(r, tris[i], i, hit)`.
static void IntersectLeafAccelerated(const Scene& sc, const LeafAccel& a, const Ray& r, Hit& hit) {
  bool from_leaf = false;  // the current best came from this leaf (only then does the leaf position break ties)
  int stack[128];
  int sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    const SubNode& n = a.nodes[(size_t)stack[--sp]];
    const float dst = IntersectAABB(r, n.mn, n.mx);
    if (dst >= kInfinity || dst > hit.t) continue;  // '>' : a triangle tying with the best may sit exactly at the entry
    if (n.count > 0) {
      for (int k = 0; k < n.count; k++) {
        const int idx = a.order[(size_t)(n.first + k)];
        Hit h{false, kInfinity, -1, 0, 0};
        IntersectTriangle(r, sc.tris[(size_t)idx], idx, h);
        if (h.hit && (h.t < hit.t || (h.t == hit.t && from_leaf && idx < hit.tri))) { hit = h; from_leaf = true; }
      }
    } else {
      if (sp + 2 > 128) abort();
      stack[sp++] = n.right;
      stack[sp++] = n.left;
    }
  }
}

// TraverseBVH :225-267
static Hit TraverseBVH(const Scene& sc, const Ray& r, Counters& c) {
  Hit hit{false, kInfinity, -1, 0, 0};
  if (sc.nodes.empty()) return hit;  // defined behaviour for the empty scene
  int stack[256];
  int sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    int ni = stack[--sp];
    const GpuNode& n = sc.nodes[ni];
    c.nodes_visited++;
    float dst = IntersectAABB(r, n.mn, n.mx);
    // The reference culls with `dst >= hit.t` (:245-246) on FP32 slab distances of exact boxes.  Between two nearly coincident
    // surfaces that discards, now and then, the box of the NEARER triangle (its entry distance rounds to >= a best t that is only an
    // ulp or two farther), and the reference keeps the farther hit.  exact_closest keeps such nodes: the cull needs a clear margin.
    if (sc.exact_closest ? (dst == kInfinity || dst > hit.t + (1e-4f * fabsf(hit.t) + 1e-5f)) : (dst >= hit.t)) continue;
    if (n.count > 0) {
      const int accel = sc.accel_of_node[(size_t)ni];
      if (accel >= 0) { c.tris_tested += n.count; IntersectLeafAccelerated(sc, sc.leaf_accel[(size_t)accel], r, hit); }
      else for (int i = 0; i < n.count; i++) { c.tris_tested++; IntersectTriangle(r, sc.tris[n.leftOrFirst + i], n.leftOrFirst + i, hit); }
    } else {
      stack[sp++] = n.leftOrFirst + 1;
      stack[sp++] = n.leftOrFirst;
      if (sp > c.max_stack) c.max_stack = sp;
      if (sp > 254) abort();
    }
  }
  return hit;
}
// SphereInstance.IntersectUnitSphere, HittableObjects.cs:82-107
static inline bool IntersectUnitSphere(V3 o, V3 d, float* t) {
  float a = dot(d, d);
  float b = 2.0f * dot(o, d);
  float c = dot(o, o) - 1.0f;
  float disc = b * b - (4.0f * a) * c;
  if (disc < 0.0f) return false;
  float s = sqrtf(disc);
  float t0 = (-b - s) / (2.0f * a), t1 = (-b + s) / (2.0f * a);
  *t = (t0 > 1e-3f) ? t0 : t1;
  return *t > 1e-3f;
}
// BoxInstance.IntersectUnitBox, HittableObjects.cs:180-223; face: 0 none, 1 -x, 2 +x, 3 -y, 4 +y, 5 -z, 6 +z
static inline bool IntersectUnitBox(V3 o, V3 d, float* t, int* face) {
  float tmin = -1e20f, tmax = 1e20f;
  int nmin = 0, nmax = 0;
  const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
  for (int axis = 0; axis < 3; axis++) {
    float invD = fabsf(dd[axis]) > 1e-8f ? 1.0f / dd[axis] : INFINITY;
    float t1 = (-0.5f - oo[axis]) * invD, t2 = (0.5f - oo[axis]) * invD;
    int n1 = 1 + 2 * axis, n2 = 2 + 2 * axis;
    if (t1 > t2) { std::swap(t1, t2); std::swap(n1, n2); }
    if (t1 > tmin) { tmin = t1; nmin = n1; }
    if (t2 < tmax) { tmax = t2; nmax = n2; }
    if (tmin > tmax) return false;
    if (tmax < 1e-3f) return false;
  }
  *t = tmin >= 1e-3f ? tmin : tmax;
  *face = (*t == tmin) ? nmin : nmax;
  return *t >= 1e-3f;
}
static inline V3 face_normal(int face) {
  switch (face) {
    case 1: return v3(-1, 0, 0); case 2: return v3(1, 0, 0); case 3: return v3(0, -1, 0);
    case 4: return v3(0, 1, 0); case 5: return v3(0, 0, -1); case 6: return v3(0, 0, 1);
    default: return v3(0, 0, 0);
  }
}
static inline V3 analytic_point_os(const Analytic& p, const Ray& r, float tOS) {  // pOS = rOS.origin + tOS * rOS.direction
  V3 o = mul_point3x4(p.W, r.o), d = unity_normalized(mul_vector(p.W, r.d));
  return o + tOS * d;
}
// SphereInstance.Hit :45-77 / BoxInstance.Hit :149-174 with tMin = Epsilon, tMax = best.t of the live kernel's loop
static inline void IntersectAnalytic(const Ray& r, const Analytic& p, int index, Hit& best) {
  V3 o = mul_point3x4(p.W, r.o), d = unity_normalized(mul_vector(p.W, r.d));
  float tOS; int face = 0;
  if (p.kind == 1 ? !IntersectUnitSphere(o, d, &tOS) : !IntersectUnitBox(o, d, &tOS, &face)) return;
  V3 pOS = o + tOS * d;
  V3 pWS = mul_point3x4(p.M, pOS);
  V3 dv = pWS - r.o;
  float tWS = sqrtf((dv.x * dv.x + dv.y * dv.y) + dv.z * dv.z);  // .magnitude
  if (tWS < best.t && tWS > kEpsilon) { best.hit = true; best.t = tWS; best.tri = -1; best.prim = index; best.u = tOS; best.v = (float)face; }
}
// worldToObject.transpose.MultiplyVector(nOS).normalized
static inline V3 analytic_normal(const Analytic& p, const Ray& r, const Hit& h) {
  V3 nOS = p.kind == 1 ? unity_normalized(analytic_point_os(p, r, h.u)) : face_normal((int)h.v);
  const M4& W = p.W;
  return unity_normalized(v3((W.m[0][0] * nOS.x + W.m[1][0] * nOS.y) + W.m[2][0] * nOS.z, (W.m[0][1] * nOS.x + W.m[1][1] * nOS.y) + W.m[2][1] * nOS.z,
                             (W.m[0][2] * nOS.x + W.m[1][2] * nOS.y) + W.m[2][2] * nOS.z));
}
static Hit TraverseBVH(const Scene& sc, const Ray& r, Counters& c);
// The scene query of analytic mode: triangles through the BVH, then every analytic primitive in emission order; the
// strict "<" keeps the first of equal t, as everywhere else.
static Hit TraverseScene(const Scene& sc, const Ray& r, Counters& c) {
  Hit hit = TraverseBVH(sc, r, c);
  for (size_t i = 0; i < sc.prims.size(); i++) IntersectAnalytic(r, sc.prims[i], (int)i, hit);
  return hit;
}

static inline V3 hit_normal(const Tri& t, float u, float v) {  // :186-187
  float w = 1.0f - u - v;
  return hlsl_normalize((w * t.n0 + u * t.n1) + v * t.n2);
}

struct Frame {  // everything RayTracer.cs:221-355 resolves into uniforms
  int w, h, n_samples, gridW, gridH;
  M4 camToObj;
  float camDist, tanHalf, orthoSize;
  V3 lightPos, bg;
  int maxDepth, amb, dif, spec, refr, ortho, soft, glossy, blur, debug;
  float lightIntensity, lightSize, roughness, shutter;
};

static Frame resolve_frame(const Scene& sc, const rtb_render_params& p) {
  const rtb_scene_desc& d = sc.owned.d;
  Frame f;
  f.w = p.has_resolution ? p.width : std::max(1, d.has_image ? d.image_w : 256);   // RayTracer.cs:221
  f.h = p.has_resolution ? p.height : std::max(1, d.has_image ? d.image_h : 256);  // :222
  M4 Mscene = m4_identity();
  if (d.has_camera && d.cam_xform >= 0 && d.cam_xform < d.n_xforms) Mscene = build_matrix(d, d.cam_xform);  // :238-243
  if (p.has_cam_pos || p.has_cam_rot) {  // :251-261
    V3 pos = p.has_cam_pos ? v3(p.cam_pos[0], p.cam_pos[1], p.cam_pos[2]) : v3(0, 0, 0);
    V3 rot = p.has_cam_rot ? v3(p.cam_rot_euler_deg[0], p.cam_rot_euler_deg[1], p.cam_rot_euler_deg[2]) : v3(0, 0, 0);
    M4 trs = m4_rotate(quat_euler(rot.x, rot.y, rot.z));
    trs.m[0][3] = pos.x; trs.m[1][3] = pos.y; trs.m[2][3] = pos.z;
    f.camToObj = m4_inverse(trs);
  } else {
    f.camToObj = m4_inverse(Mscene);  // :266
  }
  f.bg = p.has_bg ? v3(p.bg[0], p.bg[1], p.bg[2]) : (d.has_image ? v3(d.bg[0], d.bg[1], d.bg[2]) : v3(0.2f, 0.2f, 0.2f));  // :322
  f.lightPos = v3(0, 0, 0);  // :325-336
  if (d.n_lights > 0) {
    int li = d.light_xforms[0];
    if (li >= 0 && li < d.n_xforms) { M4 lm = build_matrix(d, li); f.lightPos = v3(lm.m[0][3], lm.m[1][3], lm.m[2][3]); }
  }
  float fov = p.has_fov ? p.fov_deg : (d.has_camera ? d.cam_vfov_deg : 50.0f);  // :339
  f.camDist = d.has_camera ? d.cam_distance : 30.0f;                            // :340
  // halfHeight = _CameraDistance * tan(radians(_CameraFOV) * 0.5)  (BVHRayTracing.compute:292); tan evaluated once on the host
  f.tanHalf = (float)tan((double)((fov * 0.017453292f) * 0.5f));
  f.orthoSize = f.camDist * (float)tan((double)(kDeg2Rad * fov * 0.5f));        // RayTracer.cs:347
  f.n_samples = std::max(1, p.aa_samples);                                      // compute:283
  float gs = sqrtf((float)f.n_samples);
  f.gridW = (int)ceilf(gs);
  f.gridH = (int)ceilf((float)f.n_samples / (float)f.gridW);
  f.maxDepth = p.max_depth; f.amb = p.enable_ambient; f.dif = p.enable_diffuse; f.spec = p.enable_specular; f.refr = p.enable_refraction;
  f.ortho = p.is_orthographic; f.soft = p.soft_shadows; f.glossy = p.glossy; f.blur = p.motion_blur; f.debug = p.debug_mode;
  f.lightIntensity = p.light_intensity; f.lightSize = p.light_size; f.roughness = p.roughness; f.shutter = p.shutter_speed;
  return f;
}

// Ray generation, CSMain :291-349.  i == -1: the un-jittered, un-blurred pixel-centre ray under the current projection
// (what rtb_render_aux reports); i == -2: the perspective centre ray of the debug views (:486-489).
static Ray gen_ray(const Frame& f, int px, int py, int i) {
  float width = (float)f.w, height = (float)f.h;
  float aspect = width / height;
  float halfHeight = f.camDist * f.tanHalf;
  float planeHeight = 2.0f * halfHeight;
  float planeWidth = planeHeight * aspect;
  float ox = 0.5f, oy = 0.5f;
  if (i >= 0 && f.n_samples > 1) {
    int gy = i / f.gridW, gx = i % f.gridW;
    float jx, jy;
    Hash22((float)px + (float)i * 13.0f, (float)py + (float)i * 7.0f, &jx, &jy);
    ox = ((float)gx + jx) / (float)f.gridW;
    oy = ((float)gy + jy) / (float)f.gridH;
  }
  float u = (((float)px + ox) / width - 0.5f) * planeWidth;
  float v = (((float)py + oy) / height - 0.5f) * planeHeight;
  V3 oc, dc;
  if (f.ortho == 1 && i >= -1) {
    float ohh = f.orthoSize, ohw = ohh * aspect;
    float ou = ((((float)px + ox) / width - 0.5f) * 2.0f) * ohw;
    float ov = ((((float)py + oy) / height - 0.5f) * 2.0f) * ohh;
    oc = v3(ou, ov, f.camDist);
    dc = v3(0, 0, -1);
  } else {
    oc = v3(0, 0, f.camDist);
    dc = hlsl_normalize(v3(u, v, 0) - oc);
  }
  const M4& M = f.camToObj;
  V3 o = v3(((M.m[0][0] * oc.x + M.m[0][1] * oc.y) + M.m[0][2] * oc.z) + M.m[0][3] * 1.0f,
            ((M.m[1][0] * oc.x + M.m[1][1] * oc.y) + M.m[1][2] * oc.z) + M.m[1][3] * 1.0f,
            ((M.m[2][0] * oc.x + M.m[2][1] * oc.y) + M.m[2][2] * oc.z) + M.m[2][3] * 1.0f);
  V3 d = hlsl_normalize(mul_vector(M, dc));
  if (f.blur == 1 && i >= 0) {
    V3 r = RandomUnitVector(v3((float)px + (float)i, (float)py, (float)i));
    V3 shake = ((r - v3(0.5f, 0.5f, 0.5f)) * 0.2f) * f.shutter;
    o = o + shake;
  }
  return CreateRay(o, d);
}

static inline void material_of(const Scene& sc, int idx, V3& col, float& ka, float& kd, float& ks, float& kr, float& ior) {
  col = v3(1, 1, 1); ka = 0.1f; kd = 0.7f; ks = 0.0f; kr = 0.0f; ior = 1.0f;  // :371-372
  if (idx >= 0 && idx < (int)sc.mats.size()) {  // out-of-range index -> defaults
    const rtb_material& m = sc.mats[idx];
    col = v3(m.r, m.g, m.b); ka = m.ka; kd = m.kd; ks = m.ks; kr = m.kr; ior = m.ior;
  }
}

// One AA sample of CSMain's depth loop, :356-473
static V3 trace_sample(const Scene& sc, const Frame& f, int px, int py, int i, Counters& c, Hit* primary_out) {
  V3 sampleColor = v3(0, 0, 0), attenuation = v3(1, 1, 1);
  Ray ray = gen_ray(f, px, py, i);
  for (int depth = 0; depth < f.maxDepth; depth++) {
    if (depth == 0) c.rays_primary++; else c.rays_continuation++;
    Hit hit = TraverseScene(sc, ray, c);
    if (depth == 0 && primary_out) *primary_out = hit;
    if (!hit.hit) { sampleColor = sampleColor + attenuation * f.bg; break; }
    c.closest_hits++;
    if (depth == 0) c.primary_hits++;
    V3 pos, n;
    int material_index;
    if (hit.prim >= 0) {  // analytic primitive: rec.positionWS = pWS, rec.normalWS (HittableObjects.cs:65-71)
      const Analytic& ap = sc.prims[hit.prim];
      pos = mul_point3x4(ap.M, analytic_point_os(ap, ray, hit.u));
      n = analytic_normal(ap, ray, hit);
      material_index = ap.material;
    } else {
      const Tri& tri = sc.tris[hit.tri];
      pos = ray.o + hit.t * ray.d;
      n = hit_normal(tri, hit.u, hit.v);
      material_index = tri.material;
    }
    V3 col; float ka, kd, ks, kr, ior;
    material_of(sc, material_index, col, ka, kd, ks, kr, ior);
    V3 local = v3(0, 0, 0);
    if (f.amb == 1) local = local + col * ka;
    V3 lightPos = f.lightPos;
    if (f.soft == 1) {
      V3 j = RandomUnitVector(v3((float)px + (float)i * 9.0f, ((float)py + (float)i * 4.0f) + (float)depth, (float)i)) * f.lightSize;
      lightPos = lightPos + j;
    }
    V3 toL = lightPos - pos;
    V3 lightDir = hlsl_normalize(toL);
    float NdotL = fmax_(0.0f, dot(n, lightDir));
    if (f.dif == 1 && NdotL > 0.0f) {
      Ray sr; sr.o = pos + n * kOffset; sr.d = lightDir; sr.inv = v3(1.0f / lightDir.x, 1.0f / lightDir.y, 1.0f / lightDir.z);
      float dist = hlsl_length(toL);
      c.rays_shadow++;
      const int64_t n0 = c.nodes_visited, t0 = c.tris_tested;
      Hit sh = TraverseScene(sc, sr, c);
      c.nodes_visited_shadow += c.nodes_visited - n0;
      c.tris_tested_shadow += c.tris_tested - t0;
      if (!sh.hit || sh.t > dist) {
        local = local + (col * kd) * NdotL;
        if (f.spec == 1 && ks > 0.0f) {
          V3 viewDir = hlsl_normalize(neg(ray.d));
          V3 halfVec = hlsl_normalize(lightDir + viewDir);
          float sp = pow32(fmax_(dot(n, halfVec), 0.0f));
          float k = ks * sp;
          local = local + v3(k, k, k);
        }
      }
    }
    sampleColor = sampleColor + (attenuation * local) * f.lightIntensity;
    bool shouldReflect = ks > 0.0f;
    bool shouldRefract = (f.refr == 1 && kr > 0.0f);
    if (!shouldReflect && !shouldRefract) break;
    V3 nextDir = v3(0, 0, 0), startPos = pos;
    if (shouldRefract) {
      V3 I = hlsl_normalize(ray.d), N = n;
      float eta = 1.0f / ior;
      if (dot(I, N) > 0) { N = neg(N); eta = ior; }
      float cosi = dot(neg(I), N);
      float k = 1.0f - (eta * eta) * (1.0f - cosi * cosi);
      if (k >= 0.0f) {
        nextDir = eta * I + (eta * cosi - sqrtf(k)) * N;
        attenuation = attenuation * (col * kr);
        startPos = startPos + nextDir * kOffset;
      } else {
        nextDir = hlsl_reflect(I, N);
        attenuation = attenuation * (col * ks);
        startPos = startPos + N * kOffset;
      }
    } else {
      nextDir = hlsl_reflect(hlsl_normalize(ray.d), n);
      attenuation = attenuation * (col * ks);
      startPos = startPos + n * kOffset;
    }
    if (f.glossy == 1 && f.roughness > 0.0f) {
      V3 j = RandomUnitVector(v3(((float)px + (float)i * 55.0f) + (float)depth, (float)py + (float)i * 22.0f, (float)(depth * 13))) * f.roughness;
      nextDir = hlsl_normalize(nextDir + j);
    }
    ray = CreateRay(startPos, hlsl_normalize(nextDir));
  }
  return sampleColor;
}

static inline uint8_t quantize(float c) {
  if (!(c == c)) c = 0.0f;
  c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
  return (uint8_t)(int)floorf(c * 255.0f + 0.5f);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// C API (ctypes-friendly)
// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

struct orc_scene { Scene s; };

typedef struct orc_counters {
  int64_t rays_primary, rays_continuation, rays_shadow, nodes_visited, tris_tested, closest_hits, primary_hits;
  int32_t max_stack, threads;
  double seconds;
  int64_t nodes_visited_shadow, tris_tested_shadow;
} orc_counters;

int orc_build(const rtb_scene_desc* d, orc_scene** out) {
  if (!d || !out) return RTB_E_ARG;
  orc_scene* o = new orc_scene();
  o->s.owned.copy_from(*d);
  build_scene(o->s);
  *out = o;
  return RTB_OK;
}

int orc_parse(const char* text, size_t len, orc_scene** out, char* err, size_t cap) {
  orc_scene* o = new orc_scene();
  try { parse_scene_text(std::string(text, len), o->s.owned); }
  catch (ParseError& e) { if (err && cap) snprintf(err, cap, "%s", e.msg.c_str()); delete o; return RTB_E_PARSE; }
  build_scene(o->s);
  *out = o;
  return RTB_OK;
}

int orc_load(const char* path, orc_scene** out, char* err, size_t cap) {
  FILE* f = fopen(path, "rb");
  if (!f) { if (err && cap) snprintf(err, cap, "cannot open %s", path); return RTB_E_IO; }
  std::string text; char buf[65536]; size_t n;
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
  fclose(f);
  return orc_parse(text.data(), text.size(), out, err, cap);
}

// RTB_PRIM_TESSELLATED (0, the reference's behaviour) or RTB_PRIM_ANALYTIC (1): rebuilds the scene.
int orc_set_primitive_mode(orc_scene* s, int32_t mode) {
  if (!s || (mode != RTB_PRIM_TESSELLATED && mode != RTB_PRIM_ANALYTIC)) return RTB_E_ARG;
  s->s.analytic = mode == RTB_PRIM_ANALYTIC;
  build_scene(s->s);
  return RTB_OK;
}
int64_t orc_n_primitives(const orc_scene* s) { return (int64_t)(s->s.tris_emit.size() + s->s.prims.size()); }

void orc_free(orc_scene* s) { delete s; }
const rtb_scene_desc* orc_desc(const orc_scene* s) { return &s->s.owned.d; }
int64_t orc_n_triangles(const orc_scene* s) { return (int64_t)s->s.tris_emit.size(); }
int64_t orc_n_nodes(const orc_scene* s) { return (int64_t)s->s.nodes.size(); }
int32_t orc_max_leaf(const orc_scene* s) { return s->s.max_leaf; }

// Emission-order triangles: 18 floats (v0 v1 v2 n0 n1 n2) + material; centres separately (3 floats).
void orc_get_triangles(const orc_scene* s, float* vn18, int32_t* material, float* center3) {
  const auto& T = s->s.tris_emit;
  for (size_t i = 0; i < T.size(); i++) {
    const V3 a[6] = {T[i].v0, T[i].v1, T[i].v2, T[i].n0, T[i].n1, T[i].n2};
    if (vn18) for (int k = 0; k < 6; k++) { vn18[i * 18 + k * 3] = a[k].x; vn18[i * 18 + k * 3 + 1] = a[k].y; vn18[i * 18 + k * 3 + 2] = a[k].z; }
    if (material) material[i] = T[i].material;
    if (center3) { center3[i * 3] = T[i].center.x; center3[i * 3 + 1] = T[i].center.y; center3[i * 3 + 2] = T[i].center.z; }
  }
}
// Flattened nodes (8 x 4 bytes each: min.xyz, leftOrFirst, max.xyz, count) and leaf-order -> emission-order map.
void orc_get_bvh(const orc_scene* s, void* nodes32, int32_t* orig) {
  if (nodes32) memcpy(nodes32, s->s.nodes.data(), s->s.nodes.size() * sizeof(GpuNode));
  if (orig) memcpy(orig, s->s.orig.data(), s->s.orig.size() * sizeof(int));
}

int orc_resolve(const orc_scene* s, const rtb_render_params* p, int32_t* w, int32_t* h) {
  Frame f = resolve_frame(s->s, *p);
  *w = f.w; *h = f.h;
  return RTB_OK;
}
// Uniform block the reference would upload (for host-logic parity tests): camToObj row-major 16, then
// camDist, tanHalf, orthoSize, light xyz, bg xyz.
void orc_get_frame(const orc_scene* s, const rtb_render_params* p, float* out25) {
  Frame f = resolve_frame(s->s, *p);
  memcpy(out25, f.camToObj.m, 64);
  float* o = out25 + 16;
  o[0] = f.camDist; o[1] = f.tanHalf; o[2] = f.orthoSize; o[3] = f.lightPos.x; o[4] = f.lightPos.y; o[5] = f.lightPos.z;
  o[6] = f.bg.x; o[7] = f.bg.y; o[8] = f.bg.z;
}

// Renders rows row_begin, row_begin+row_step, ... < row_end (row 0 = bottom).  Outputs are full-frame arrays (W*H);
// rows not rendered are left untouched.  Any output pointer may be NULL.
int orc_render(const orc_scene* s, const rtb_render_params* p, int32_t row_begin, int32_t row_end, int32_t row_step, int32_t threads,
               uint8_t* rgba8, float* rgbf, int32_t* prim, float* tout, int32_t* mat, orc_counters* cnt) {
  if (!s || !p) return RTB_E_ARG;
  const Scene& sc = s->s;
  Frame f = resolve_frame(sc, *p);
  if (row_end < 0 || row_end > f.h) row_end = f.h;
  if (row_step < 1) row_step = 1;
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = threads > 0 ? threads : omp_get_max_threads();
#endif
  std::vector<Counters> cs((size_t)nthreads);
  int nrows = (row_end - row_begin + row_step - 1) / row_step;
  double t0 = 0, t1 = 0;
#ifdef _OPENMP
  t0 = omp_get_wtime();
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
  for (int ri = 0; ri < nrows; ri++) {
    int y = row_begin + ri * row_step;
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    Counters& c = cs[(size_t)tid];
    for (int x = 0; x < f.w; x++) {
      V3 accum = v3(0, 0, 0);
      Hit ph{false, kInfinity, -1, 0, 0};
      bool want_aux = prim || tout || mat;
      for (int i = 0; i < f.n_samples; i++) {
        V3 sc_col = trace_sample(sc, f, x, y, i, c, nullptr);
        accum = accum + sc_col;  // :475
      }
      float ns = (float)f.n_samples;
      V3 fin = v3(accum.x / ns, accum.y / ns, accum.z / ns);  // :478
      if (want_aux) { Counters scratch; ph = TraverseScene(sc, gen_ray(f, x, y, -1), scratch); }
      if (f.debug != 0) {
        Counters scratch;
        Ray cr = gen_ray(f, x, y, -2);  // centre ray, always perspective, :486-489
        Hit ph = TraverseScene(sc, cr, scratch);
        if (f.debug == 1) { float g = ph.t / 100.0f; fin = ph.hit ? v3(g, g, g) : v3(1, 0, 0); }
        else if (f.debug == 2) { if (ph.hit) { V3 n = ph.prim >= 0 ? analytic_normal(sc.prims[ph.prim], cr, ph) : hit_normal(sc.tris[ph.tri], ph.u, ph.v); fin = n * 0.5f + v3(0.5f, 0.5f, 0.5f); } else fin = v3(0, 0, 1); }
        else if (f.debug == 3) fin = ph.hit ? v3(0, 1, 0) : v3(0.2f, 0.2f, 0.2f);
      }
      size_t at = (size_t)y * (size_t)f.w + (size_t)x;
      if (rgba8) { rgba8[at * 4] = quantize(fin.x); rgba8[at * 4 + 1] = quantize(fin.y); rgba8[at * 4 + 2] = quantize(fin.z); rgba8[at * 4 + 3] = 255; }
      if (rgbf) { rgbf[at * 3] = fin.x; rgbf[at * 3 + 1] = fin.y; rgbf[at * 3 + 2] = fin.z; }
      if (prim) prim[at] = !ph.hit ? -1 : (ph.prim >= 0 ? (int)sc.tris_emit.size() + ph.prim : sc.orig[ph.tri]);
      if (tout) tout[at] = ph.t;
      if (mat) mat[at] = !ph.hit ? -1 : (ph.prim >= 0 ? sc.prims[ph.prim].material : sc.tris[ph.tri].material);
    }
  }
#ifdef _OPENMP
  t1 = omp_get_wtime();
#endif
  if (cnt) {
    memset(cnt, 0, sizeof *cnt);
    for (auto& c : cs) {
      cnt->rays_primary += c.rays_primary; cnt->rays_continuation += c.rays_continuation; cnt->rays_shadow += c.rays_shadow;
      cnt->nodes_visited += c.nodes_visited; cnt->tris_tested += c.tris_tested; cnt->closest_hits += c.closest_hits;
      cnt->primary_hits += c.primary_hits; cnt->max_stack = std::max(cnt->max_stack, c.max_stack);
      cnt->nodes_visited_shadow += c.nodes_visited_shadow; cnt->tris_tested_shadow += c.tris_tested_shadow;
    }
    cnt->threads = nthreads;
    cnt->seconds = t1 - t0;
  }
  return RTB_OK;
}

// The pixel-centre primary ray of (px,py): origin and direction (for tie analysis in tests).
void orc_primary_ray(const orc_scene* s, const rtb_render_params* p, int32_t px, int32_t py, float* o3, float* d3) {
  Frame f = resolve_frame(s->s, *p);
  Ray r = gen_ray(f, px, py, -1);
  o3[0] = r.o.x; o3[1] = r.o.y; o3[2] = r.o.z; d3[0] = r.d.x; d3[1] = r.d.y; d3[2] = r.d.z;
}

// RandomUnitVector (compute:124-131) as the oracle evaluates it: Hash33 in FP32 and the fixed polynomial sin / cos.
void orc_random_unit_vector(const float* seed3, float* out3) {
  const V3 r = RandomUnitVector(v3(seed3[0], seed3[1], seed3[2]));
  out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

// The primary ray of AA sample `sample` of pixel (px,py) (stratified cell + Hash22 jitter, compute:296-340); sample = -1: centre ray.
void orc_sample_ray(const orc_scene* s, const rtb_render_params* p, int32_t px, int32_t py, int32_t sample, float* o3, float* d3) {
  Frame f = resolve_frame(s->s, *p);
  Ray r = gen_ray(f, px, py, sample);
  o3[0] = r.o.x; o3[1] = r.o.y; o3[2] = r.o.z; d3[0] = r.d.x; d3[1] = r.d.y; d3[2] = r.d.z;
}

// Brute force over ALL triangles in emission order (no BVH): closest t, and the emission ids attaining exactly that t.
// Returns the number of ties (>=1 on a hit, 0 on a miss); at most `cap` ids are written.
int32_t orc_brute_closest(const orc_scene* s, const float* o3, const float* d3, float* t_out, int32_t* ids, int32_t cap) {
  const Scene& sc = s->s;
  Ray r = CreateRay(v3(o3[0], o3[1], o3[2]), v3(d3[0], d3[1], d3[2]));
  float best = kInfinity;
  for (size_t i = 0; i < sc.tris_emit.size(); i++) {
    Hit h{false, kInfinity, -1, 0, 0};
    IntersectTriangle(r, sc.tris_emit[i], (int)i, h);
    if (h.hit && h.t < best) best = h.t;
  }
  int32_t n = 0;
  if (best < kInfinity)
    for (size_t i = 0; i < sc.tris_emit.size(); i++) {
      Hit h{false, kInfinity, -1, 0, 0};
      IntersectTriangle(r, sc.tris_emit[i], (int)i, h);
      if (h.hit && h.t == best) { if (n < cap && ids) ids[n] = (int32_t)i; n++; }
    }
  if (t_out) *t_out = best;
  return n;
}

// Checker-only leaf accelerator (LeafAccel): leaves with more than `min_count` triangles get a sub-tree; 0 = plain linear scan.
// Applies to scenes built afterwards.  Returns the previous value.
int orc_set_leaf_accel(int32_t min_count) { const int prev = g_leaf_accel_min; g_leaf_accel_min = min_count < 0 ? 0 : min_count; return prev; }
int32_t orc_n_accelerated_leaves(const orc_scene* s) { return (int32_t)s->s.leaf_accel.size(); }

// Checker-only: 1 = TraverseBVH returns the exact closest hit (see Scene::exact_closest), 0 = the reference's traversal.  The
// GPU-LBVH flavour of the product, which tests padded boxes, is held against this mode; the reference-shape flavour against mode 0.
int orc_set_exact_closest(orc_scene* s, int32_t on) {
  if (!s) return RTB_E_ARG;
  s->s.exact_closest = on != 0;
  return RTB_OK;
}

int orc_version(void) { return 2; }

}  // extern "C"
