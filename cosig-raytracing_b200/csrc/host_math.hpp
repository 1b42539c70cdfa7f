// host_math.hpp — host-side restatement of the UnityEngine maths the reference's render set-up depends on
// (Matrix4x4 / Quaternion / Vector3: closed source, not under the reference tree; definitions fixed in SURVEY.md App. D).
//
// Used by the scene-upload path (composite transforms, RayTracer.cs:410-437 / SceneGeometryConverter.cs:83-114) and by the
// per-frame uniform resolve (RayTracer.cs:221-355).  Everything here is FP32 with one rounding per operation, evaluated in
// the order written; transcendental inputs go through double-precision libm and are rounded once.  Build with
// -ffp-contract=off.  Column-major storage like Unity: c[col][row].
#pragma once
#include <cmath>
#include <cstring>

namespace rtb {

struct Vec3f { float x, y, z; };

struct Mat4 {
  float c[4][4];  // c[column][row]
  float& at(int row, int col) { return c[col][row]; }
  float at(int row, int col) const { return c[col][row]; }
  static Mat4 identity() {
    Mat4 m;
    std::memset(&m, 0, sizeof m);
    for (int i = 0; i < 4; i++) m.c[i][i] = 1.0f;
    return m;
  }
};

// Matrix4x4.operator*: each element is a left-to-right FP32 sum of four products.
inline Mat4 operator*(const Mat4& l, const Mat4& r) {
  Mat4 o;
  for (int col = 0; col < 4; col++)
    for (int row = 0; row < 4; row++) {
      float acc = l.at(row, 0) * r.at(0, col);
      acc = acc + l.at(row, 1) * r.at(1, col);
      acc = acc + l.at(row, 2) * r.at(2, col);
      acc = acc + l.at(row, 3) * r.at(3, col);
      o.at(row, col) = acc;
    }
  return o;
}

inline Mat4 translate(Vec3f t) { Mat4 m = Mat4::identity(); m.at(0, 3) = t.x; m.at(1, 3) = t.y; m.at(2, 3) = t.z; return m; }
inline Mat4 scale(Vec3f s) { Mat4 m = Mat4::identity(); m.at(0, 0) = s.x; m.at(1, 1) = s.y; m.at(2, 2) = s.z; return m; }

struct Quatf { float x, y, z, w; };
constexpr float kDeg2Rad = 0.0174532924f;  // Mathf.Deg2Rad

// Quaternion.AngleAxis(deg, unit axis)
inline Quatf angle_axis(float deg, Vec3f axis) {
  const float half = deg * kDeg2Rad * 0.5f;
  const float s = static_cast<float>(std::sin(static_cast<double>(half)));
  const float c = static_cast<float>(std::cos(static_cast<double>(half)));
  return Quatf{axis.x * s, axis.y * s, axis.z * s, c};
}
inline Quatf operator*(Quatf a, Quatf b) {
  Quatf q;
  q.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  q.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  q.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  q.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return q;
}
// Quaternion.Euler: Z first, then X, then Y.
inline Quatf euler(float xd, float yd, float zd) {
  return (angle_axis(yd, Vec3f{0, 1, 0}) * angle_axis(xd, Vec3f{1, 0, 0})) * angle_axis(zd, Vec3f{0, 0, 1});
}
// Matrix4x4.Rotate(q)
inline Mat4 rotate(Quatf q) {
  const float x2 = q.x * 2.0f, y2 = q.y * 2.0f, z2 = q.z * 2.0f;
  const float xx = q.x * x2, yy = q.y * y2, zz = q.z * z2;
  const float xy = q.x * y2, xz = q.x * z2, yz = q.y * z2;
  const float wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
  Mat4 m = Mat4::identity();
  m.at(0, 0) = 1.0f - (yy + zz); m.at(0, 1) = xy - wz;          m.at(0, 2) = xz + wy;
  m.at(1, 0) = xy + wz;          m.at(1, 1) = 1.0f - (xx + zz); m.at(1, 2) = yz - wx;
  m.at(2, 0) = xz - wy;          m.at(2, 1) = yz + wx;          m.at(2, 2) = 1.0f - (xx + yy);
  return m;
}
// Matrix4x4.TRS with unit scale
inline Mat4 trs_unit_scale(Vec3f pos, Quatf q) {
  Mat4 m = rotate(q);
  m.at(0, 3) = pos.x; m.at(1, 3) = pos.y; m.at(2, 3) = pos.z;
  return m;
}

// Matrix4x4.inverse: classical adjugate over determinant, all in double, one rounding to FP32 per element.
// Singular input gives the zero matrix (Unity's documented behaviour).
inline Mat4 inverse(const Mat4& a) {
  double m[16];  // row-major scratch
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) m[r * 4 + c] = static_cast<double>(a.at(r, c));
  double v[16];
  v[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
  v[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
  v[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
  v[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
  v[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
  v[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
  v[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
  v[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
  v[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
  v[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
  v[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
  v[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
  v[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
  v[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
  v[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
  v[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
  const double det = m[0] * v[0] + m[1] * v[4] + m[2] * v[8] + m[3] * v[12];
  Mat4 o;
  if (det == 0.0) { std::memset(&o, 0, sizeof o); return o; }
  const double rdet = 1.0 / det;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) o.at(r, c) = static_cast<float>(v[r * 4 + c] * rdet);
  return o;
}
inline Mat4 transpose(const Mat4& a) {
  Mat4 o;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) o.at(r, c) = a.at(c, r);
  return o;
}

}  // namespace rtb
