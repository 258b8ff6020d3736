// glmath.h — the six GLM 0.9.3.1 (degrees API) functions the reference's host code uses, re-stated in fp32
// with the same operation order so that matrices are bit-identical to the reference's
// (ShadowMapping/include/glm/gtc/matrix_transform.inl:32-78,223-244,383-411,
//  glm/core/type_mat4x4.inl:757-779, glm/gtc/matrix_inverse.inl:77-100, glm/core/func_geometric.inl:239-248).
// Column-major: m[c*4 + r].  Built with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstring>

namespace sgh {

struct Vec3 { float x, y, z; };
struct Mat4 { float m[16]; };
struct Mat3 { float m[9]; };

inline Mat4 identity() { Mat4 r; std::memset(r.m, 0, 64); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f; return r; }

inline Mat4 mul(const Mat4& a, const Mat4& b) {
  Mat4 r;
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < 4; k++)
      r.m[c * 4 + k] = ((a.m[k] * b.m[c * 4] + a.m[4 + k] * b.m[c * 4 + 1]) + a.m[8 + k] * b.m[c * 4 + 2]) + a.m[12 + k] * b.m[c * 4 + 3];
  return r;
}

inline float radians(float deg) { const float pi = (float)3.1415926535897932384626433832795; return deg * (pi / 180.0f); }

inline Vec3 normalize(Vec3 v) {
  float sqr = v.x * v.x + v.y * v.y + v.z * v.z;
  float inv = 1.0f / std::sqrt(sqr);
  return Vec3{v.x * inv, v.y * inv, v.z * inv};
}
inline Vec3 cross(Vec3 a, Vec3 b) { return Vec3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

inline Mat4 perspective(float fovy, float aspect, float zn, float zf) {
  float range = std::tan(radians(fovy / 2.0f)) * zn;
  float left = -range * aspect, right = range * aspect, bottom = -range, top = range;
  Mat4 r; std::memset(r.m, 0, 64);
  r.m[0] = (2.0f * zn) / (right - left);
  r.m[5] = (2.0f * zn) / (top - bottom);
  r.m[10] = -(zf + zn) / (zf - zn);
  r.m[11] = -1.0f;
  r.m[14] = -(2.0f * zf * zn) / (zf - zn);
  return r;
}

inline Mat4 translate(const Mat4& m, Vec3 v) {
  Mat4 r = m;
  for (int k = 0; k < 4; k++) r.m[12 + k] = ((m.m[k] * v.x + m.m[4 + k] * v.y) + m.m[8 + k] * v.z) + m.m[12 + k];
  return r;
}

inline Mat4 lookAt(Vec3 eye, Vec3 center, Vec3 up) {
  Vec3 f = normalize(Vec3{center.x - eye.x, center.y - eye.y, center.z - eye.z});
  Vec3 u = normalize(up);
  Vec3 s = normalize(cross(f, u));
  u = cross(s, f);
  Mat4 r = identity();
  r.m[0] = s.x; r.m[4] = s.y; r.m[8] = s.z;
  r.m[1] = u.x; r.m[5] = u.y; r.m[9] = u.z;
  r.m[2] = -f.x; r.m[6] = -f.y; r.m[10] = -f.z;
  return translate(r, Vec3{-eye.x, -eye.y, -eye.z});
}

inline Mat4 rotate(const Mat4& m, float angle, Vec3 v) {
  float a = radians(angle);
  float c = std::cos(a), s = std::sin(a);
  Vec3 axis = normalize(v);
  Vec3 temp{(1.0f - c) * axis.x, (1.0f - c) * axis.y, (1.0f - c) * axis.z};
  float R[3][3];
  R[0][0] = c + temp.x * axis.x;
  R[0][1] = 0 + temp.x * axis.y + s * axis.z;
  R[0][2] = 0 + temp.x * axis.z - s * axis.y;
  R[1][0] = 0 + temp.y * axis.x - s * axis.z;
  R[1][1] = c + temp.y * axis.y;
  R[1][2] = 0 + temp.y * axis.z + s * axis.x;
  R[2][0] = 0 + temp.z * axis.x + s * axis.y;
  R[2][1] = 0 + temp.z * axis.y - s * axis.x;
  R[2][2] = c + temp.z * axis.z;
  Mat4 r;
  for (int col = 0; col < 3; col++)
    for (int k = 0; k < 4; k++) r.m[col * 4 + k] = (m.m[k] * R[col][0] + m.m[4 + k] * R[col][1]) + m.m[8 + k] * R[col][2];
  for (int k = 0; k < 4; k++) r.m[12 + k] = m.m[12 + k];
  return r;
}
inline Mat4 rotate(float angle, Vec3 v) { return rotate(identity(), angle, v); }      // glm/gtx/transform
inline Mat4 translate(Vec3 v) { return translate(identity(), v); }
inline Mat4 transpose(const Mat4& a) { Mat4 r; for (int c = 0; c < 4; c++) for (int k = 0; k < 4; k++) r.m[c * 4 + k] = a.m[k * 4 + c]; return r; }

inline Mat3 inverseTranspose3(const Mat4& mv) {
#define SGH_M(c, r) mv.m[(c) * 4 + (r)]
  float det = +SGH_M(0, 0) * (SGH_M(1, 1) * SGH_M(2, 2) - SGH_M(1, 2) * SGH_M(2, 1)) -
              SGH_M(0, 1) * (SGH_M(1, 0) * SGH_M(2, 2) - SGH_M(1, 2) * SGH_M(2, 0)) +
              SGH_M(0, 2) * (SGH_M(1, 0) * SGH_M(2, 1) - SGH_M(1, 1) * SGH_M(2, 0));
  Mat3 inv;
  inv.m[0] = +(SGH_M(1, 1) * SGH_M(2, 2) - SGH_M(2, 1) * SGH_M(1, 2));
  inv.m[1] = -(SGH_M(1, 0) * SGH_M(2, 2) - SGH_M(2, 0) * SGH_M(1, 2));
  inv.m[2] = +(SGH_M(1, 0) * SGH_M(2, 1) - SGH_M(2, 0) * SGH_M(1, 1));
  inv.m[3] = -(SGH_M(0, 1) * SGH_M(2, 2) - SGH_M(2, 1) * SGH_M(0, 2));
  inv.m[4] = +(SGH_M(0, 0) * SGH_M(2, 2) - SGH_M(2, 0) * SGH_M(0, 2));
  inv.m[5] = -(SGH_M(0, 0) * SGH_M(2, 1) - SGH_M(2, 0) * SGH_M(0, 1));
  inv.m[6] = +(SGH_M(0, 1) * SGH_M(1, 2) - SGH_M(1, 1) * SGH_M(0, 2));
  inv.m[7] = -(SGH_M(0, 0) * SGH_M(1, 2) - SGH_M(1, 0) * SGH_M(0, 2));
  inv.m[8] = +(SGH_M(0, 0) * SGH_M(1, 1) - SGH_M(1, 0) * SGH_M(0, 1));
#undef SGH_M
  for (int k = 0; k < 9; k++) inv.m[k] = inv.m[k] / det;
  return inv;
}

inline Vec3 mul3(const Mat4& r, Vec3 v) {      // glm::mat3(r) * v
  return Vec3{(r.m[0] * v.x + r.m[4] * v.y) + r.m[8] * v.z, (r.m[1] * v.x + r.m[5] * v.y) + r.m[9] * v.z,
              (r.m[2] * v.x + r.m[6] * v.y) + r.m[10] * v.z};
}

inline Mat4 biasMatrix() {                     // MyGLGeometryViewer.cpp:139-145
  Mat4 b; std::memset(b.m, 0, 64);
  b.m[0] = 0.5f; b.m[5] = 0.5f; b.m[10] = 0.5f; b.m[12] = 0.5f; b.m[13] = 0.5f; b.m[14] = 0.5f; b.m[15] = 1.0f;
  return b;
}

}  // namespace sgh
