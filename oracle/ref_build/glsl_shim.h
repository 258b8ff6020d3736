// glsl_shim.h — just enough GLSL 1.x (compat profile) in C++ to compile the reference's UNMODIFIED
// fragment shaders on the CPU (oracle/ref_build/build_ref.py includes the .frag text verbatim after
// this header).  TEST INFRASTRUCTURE ONLY: it exists to pin oracle/ against the reference's own source.
//
// Semantics chosen where GLSL leaves them to the implementation:
//   * fp32 everywhere, evaluated in source order, no FMA contraction (-ffp-contract=off);
//   * mat*vec and dot() accumulate left to right; normalize(v) = v * (1/sqrt(dot(v,v))) (as GLM does);
//   * mix(x,y,a) = x*(1-a) + y*a, max(x,y) = (x<y)?y:x, min(x,y) = (y<x)?y:x, fract(x) = x-floor(x)
//     (GLSL 1.20 spec §8.3);
//   * texture2D = GL_NEAREST + CLAMP_TO_BORDER with border (0,0,0,0): texel = floor(coord*size)
//     (ShadowMapping/src/Viewers/MyGLTextureViewer.cpp:3-28,76-89); depth textures return (d,d,d,1).
#pragma once
#include <cmath>
#include <cstring>
#include <cstdint>

namespace glsl {

struct vec2; struct vec3; struct vec4;

// ---- swizzle proxies: trivially-constructible views living in a union with the components ------------
template <int N, int A, int B> struct Sw2 {
  float d[N];
  inline operator vec2() const;
  inline Sw2& operator=(const vec2& v);
  inline Sw2& operator+=(const vec2& v); inline Sw2& operator-=(const vec2& v);
  inline Sw2& operator*=(const vec2& v); inline Sw2& operator/=(const vec2& v);
  inline Sw2& operator*=(float s); inline Sw2& operator/=(float s);
  inline Sw2& operator+=(float s); inline Sw2& operator-=(float s);
};
template <int N, int A, int B, int C> struct Sw3 {
  float d[N];
  inline operator vec3() const;
  inline Sw3& operator=(const vec3& v);
  inline Sw3& operator+=(const vec3& v); inline Sw3& operator-=(const vec3& v);
  inline Sw3& operator*=(const vec3& v); inline Sw3& operator/=(const vec3& v);
  inline Sw3& operator*=(float s); inline Sw3& operator/=(float s);
  inline Sw3& operator+=(float s); inline Sw3& operator-=(float s);
};
template <int N, int A, int B, int C, int D> struct Sw4 {
  float d[N];
  inline operator vec4() const;
  inline Sw4& operator=(const vec4& v);
  inline Sw4& operator*=(float s); inline Sw4& operator/=(float s);
};

struct ivec2 { int x, y; };

struct vec2 {
  union {
    struct { float x, y; };
    struct { float r, g; };
    struct { float s, t; };
#include "swizzles_vec2.inc"
  };
  vec2() : x(0), y(0) {}
  vec2(float a) : x(a), y(a) {}
  vec2(float a, float b) : x(a), y(b) {}
  vec2(const ivec2& i) : x((float)i.x), y((float)i.y) {}
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
  vec2& operator+=(const vec2& o) { x += o.x; y += o.y; return *this; }
  vec2& operator-=(const vec2& o) { x -= o.x; y -= o.y; return *this; }
  vec2& operator*=(const vec2& o) { x *= o.x; y *= o.y; return *this; }
  vec2& operator/=(const vec2& o) { x /= o.x; y /= o.y; return *this; }
  vec2& operator*=(float s) { x *= s; y *= s; return *this; }
  vec2& operator/=(float s) { x /= s; y /= s; return *this; }
};

struct vec3 {
  union {
    struct { float x, y, z; };
    struct { float r, g, b; };
    struct { float s, t, p; };
#include "swizzles_vec3.inc"
  };
  vec3() : x(0), y(0), z(0) {}
  vec3(float a) : x(a), y(a), z(a) {}
  vec3(float a, float b, float c) : x(a), y(b), z(c) {}
  vec3(const vec2& v, float c) : x(v.x), y(v.y), z(c) {}
  vec3(float a, const vec2& v) : x(a), y(v.x), z(v.y) {}
  inline vec3(const vec4& v);   // implicit: PlausibleSoftShadow.frag:342 passes a vec4 for a vec3 parameter (NVIDIA accepts it)
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
  vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  vec3& operator*=(const vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
  vec3& operator/=(const vec3& o) { x /= o.x; y /= o.y; z /= o.z; return *this; }
  vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
  vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

struct vec4 {
  union {
    struct { float x, y, z, w; };
    struct { float r, g, b, a; };
    struct { float s, t, p, q; };
#include "swizzles_vec4.inc"
  };
  vec4() : x(0), y(0), z(0), w(0) {}
  vec4(float v) : x(v), y(v), z(v), w(v) {}
  vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
  vec4(const vec2& v, float c, float d) : x(v.x), y(v.y), z(c), w(d) {}
  vec4(const vec2& u, const vec2& v) : x(u.x), y(u.y), z(v.x), w(v.y) {}
  vec4(float a, float b, const vec2& v) : x(a), y(b), z(v.x), w(v.y) {}
  vec4(float a, const vec2& v, float d) : x(a), y(v.x), z(v.y), w(d) {}
  vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
  vec4(float a, const vec3& v) : x(a), y(v.x), z(v.y), w(v.z) {}
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
  vec4& operator+=(const vec4& o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
  vec4& operator-=(const vec4& o) { x -= o.x; y -= o.y; z -= o.z; w -= o.w; return *this; }
  vec4& operator*=(const vec4& o) { x *= o.x; y *= o.y; z *= o.z; w *= o.w; return *this; }
  vec4& operator/=(const vec4& o) { x /= o.x; y /= o.y; z /= o.z; w /= o.w; return *this; }
  vec4& operator*=(float s) { x *= s; y *= s; z *= s; w *= s; return *this; }
  vec4& operator/=(float s) { float k = s; x /= k; y /= k; z /= k; w /= k; return *this; }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

// ---- proxy implementations ------------------------------------------------------------------------------
template <int N, int A, int B> inline Sw2<N, A, B>::operator vec2() const { return vec2(d[A], d[B]); }
template <int N, int A, int B> inline Sw2<N, A, B>& Sw2<N, A, B>::operator=(const vec2& v) { float a = v.x, b = v.y; d[A] = a; d[B] = b; return *this; }
#define SW2_OP(OP) \
  template <int N, int A, int B> inline Sw2<N, A, B>& Sw2<N, A, B>::operator OP(const vec2& v) { float a = v.x, b = v.y; d[A] OP a; d[B] OP b; return *this; }
SW2_OP(+=) SW2_OP(-=) SW2_OP(*=) SW2_OP(/=)
#undef SW2_OP
#define SW2_OPS(OP) \
  template <int N, int A, int B> inline Sw2<N, A, B>& Sw2<N, A, B>::operator OP(float s) { d[A] OP s; d[B] OP s; return *this; }
SW2_OPS(+=) SW2_OPS(-=) SW2_OPS(*=) SW2_OPS(/=)
#undef SW2_OPS
template <int N, int A, int B, int C> inline Sw3<N, A, B, C>::operator vec3() const { return vec3(d[A], d[B], d[C]); }
template <int N, int A, int B, int C> inline Sw3<N, A, B, C>& Sw3<N, A, B, C>::operator=(const vec3& v) { float a = v.x, b = v.y, c = v.z; d[A] = a; d[B] = b; d[C] = c; return *this; }
#define SW3_OP(OP) \
  template <int N, int A, int B, int C> inline Sw3<N, A, B, C>& Sw3<N, A, B, C>::operator OP(const vec3& v) { float a = v.x, b = v.y, c = v.z; d[A] OP a; d[B] OP b; d[C] OP c; return *this; }
SW3_OP(+=) SW3_OP(-=) SW3_OP(*=) SW3_OP(/=)
#undef SW3_OP
#define SW3_OPS(OP) \
  template <int N, int A, int B, int C> inline Sw3<N, A, B, C>& Sw3<N, A, B, C>::operator OP(float s) { d[A] OP s; d[B] OP s; d[C] OP s; return *this; }
SW3_OPS(+=) SW3_OPS(-=) SW3_OPS(*=) SW3_OPS(/=)
#undef SW3_OPS
template <int N, int A, int B, int C, int D> inline Sw4<N, A, B, C, D>::operator vec4() const { return vec4(d[A], d[B], d[C], d[D]); }
template <int N, int A, int B, int C, int D> inline Sw4<N, A, B, C, D>& Sw4<N, A, B, C, D>::operator=(const vec4& v) { float a = v.x, b = v.y, c = v.z, e = v.w; d[A] = a; d[B] = b; d[C] = c; d[D] = e; return *this; }
template <int N, int A, int B, int C, int D> inline Sw4<N, A, B, C, D>& Sw4<N, A, B, C, D>::operator*=(float s) { d[A] *= s; d[B] *= s; d[C] *= s; d[D] *= s; return *this; }
template <int N, int A, int B, int C, int D> inline Sw4<N, A, B, C, D>& Sw4<N, A, B, C, D>::operator/=(float s) { d[A] /= s; d[B] /= s; d[C] /= s; d[D] /= s; return *this; }

// ---- arithmetic (non-template so that swizzle proxies convert implicitly) -------------------------------
#define VEC_BINOPS(V, ...)                                                                   \
  inline V operator+(const V& a, const V& b) { V r = a; r += b; return r; }                  \
  inline V operator-(const V& a, const V& b) { V r = a; r -= b; return r; }                  \
  inline V operator*(const V& a, const V& b) { V r = a; r *= b; return r; }                  \
  inline V operator/(const V& a, const V& b) { V r = a; r /= b; return r; }                  \
  inline V operator+(const V& a, float s) { return a + V(s); }                               \
  inline V operator-(const V& a, float s) { return a - V(s); }                               \
  inline V operator*(const V& a, float s) { return a * V(s); }                               \
  inline V operator/(const V& a, float s) { return a / V(s); }                               \
  inline V operator+(float s, const V& a) { return V(s) + a; }                               \
  inline V operator-(float s, const V& a) { return V(s) - a; }                               \
  inline V operator*(float s, const V& a) { return V(s) * a; }                               \
  inline V operator/(float s, const V& a) { return V(s) / a; }
VEC_BINOPS(vec2) VEC_BINOPS(vec3) VEC_BINOPS(vec4)
#undef VEC_BINOPS
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }

struct mat3 {
  vec3 c[3];
  mat3() {}
  mat3(float d) { c[0] = vec3(d, 0, 0); c[1] = vec3(0, d, 0); c[2] = vec3(0, 0, d); }
  vec3& operator[](int i) { return c[i]; }
  const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
  vec4 c[4];
  mat4() {}
  mat4(float d) { c[0] = vec4(d, 0, 0, 0); c[1] = vec4(0, d, 0, 0); c[2] = vec4(0, 0, d, 0); c[3] = vec4(0, 0, 0, d); }
  vec4& operator[](int i) { return c[i]; }
  const vec4& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v) {
  return vec4(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z + m[3][0] * v.w,
              m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z + m[3][1] * v.w,
              m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z + m[3][2] * v.w,
              m[0][3] * v.x + m[1][3] * v.y + m[2][3] * v.z + m[3][3] * v.w);
}
inline vec3 operator*(const mat3& m, const vec3& v) {
  return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
              m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
              m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}

// ---- built-ins ------------------------------------------------------------------------------------------
inline float abs(float x) { return std::fabs(x); }
inline float floor(float x) { return std::floor(x); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float exp(float x) { return std::exp(x); }
inline float log(float x) { return std::log(x); }
inline float log2(float x) { return std::log2(x); }
inline float pow(float x, float y) { return std::pow(x, y); }
// dFdx / dFdy: the runner supplies the derivative of the fragment (the quad differences are a property of the rasteriser, not
// of the shader source); see build_ref.py RUNNER
inline thread_local float g_dfdx = 0.0f, g_dfdy = 0.0f;
inline float dFdx(float) { return g_dfdx; }
inline float dFdy(float) { return g_dfdy; }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float fract(float x) { return x - std::floor(x); }
inline float max(float a, float b) { return a < b ? b : a; }
inline float min(float a, float b) { return b < a ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float mod(float x, float y) { return x - y * std::floor(x / y); }
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline vec2 normalize(const vec2& v) { return v * inversesqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
inline vec4 normalize(const vec4& v) { return v * inversesqrt(dot(v, v)); }
#define VEC_MAP1(F)                                                                         \
  inline vec2 F(const vec2& v) { return vec2(F(v.x), F(v.y)); }                             \
  inline vec3 F(const vec3& v) { return vec3(F(v.x), F(v.y), F(v.z)); }                     \
  inline vec4 F(const vec4& v) { return vec4(F(v.x), F(v.y), F(v.z), F(v.w)); }
VEC_MAP1(abs) VEC_MAP1(floor) VEC_MAP1(fract) VEC_MAP1(sqrt) VEC_MAP1(exp) VEC_MAP1(log) VEC_MAP1(log2)
#undef VEC_MAP1
inline vec3 reflect(const vec3& I, const vec3& N) { return I - 2.0f * dot(N, I) * N; }
inline vec2 mix(const vec2& x, const vec2& y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
inline vec4 mix(const vec4& x, const vec4& y, float a) { return x * (1.0f - a) + y * a; }
inline vec2 max(const vec2& a, const vec2& b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec2 min(const vec2& a, const vec2& b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
inline vec4 clamp(const vec4& v, float lo, float hi) { return vec4(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi), clamp(v.w, lo, hi)); }
inline vec3 clamp(const vec3& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline vec2 clamp(const vec2& v, float lo, float hi) { return vec2(clamp(v.x, lo, hi), clamp(v.y, lo, hi)); }

// ---- textures -------------------------------------------------------------------------------------------
struct Sampler {            // bound by the runner; plain data so it can be memcpy'd in
  const float* data;
  int32_t w, h;
  int32_t channels;         // 1 = depth texture (returns d,d,d,1), 4 = RGBA32F; | 0x100 = GL_LINEAR (level 0) instead of GL_NEAREST;
                            // | 0x200 = GL_REPEAT instead of CLAMP_TO_BORDER (the scene textures of loadRGBTexture, MyGLTextureViewer.cpp:45-56)
  int32_t layers;
};
typedef Sampler sampler2D;
typedef Sampler sampler2DArray;

inline vec4 sampler_texel(const Sampler& s, float fi, float fj, int layer) {     // integer texel coordinates as floats
  float fw = (float)s.w, fh = (float)s.h;
  const int ch = s.channels & 0xFF;
  if (s.channels & 0x200) { fi = fi - std::floor(fi / fw) * fw; fj = fj - std::floor(fj / fh) * fh; if (!(fi < fw)) fi = 0.0f; if (!(fj < fh)) fj = 0.0f; }
  if (!(fi >= 0.0f && fi < fw && fj >= 0.0f && fj < fh) || layer < 0 || layer >= (s.layers > 0 ? s.layers : 1))
    return ch == 1 ? vec4(0.0f, 0.0f, 0.0f, 1.0f) : vec4(0.0f);
  size_t o = ((size_t)layer * s.h + (size_t)(int)fj) * s.w + (size_t)(int)fi;
  if (ch == 1) { float d = s.data[o]; return vec4(d, d, d, 1.0f); }
  const float* p = s.data + 4 * o;
  return vec4(p[0], p[1], p[2], p[3]);
}
inline vec4 sampler_fetch(const Sampler& s, float u, float v, int layer) {
  float fw = (float)s.w, fh = (float)s.h;
  if (s.channels & 0x100) {
    // GL_LINEAR of level 0 (the level of detail of a GL_LINEAR_MIPMAP_LINEAR lookup inside a data-dependent loop is
    // undefined): weights from fract(u*size - 0.5), texels accumulated in the order 00,10,01,11
    float x = u * fw - 0.5f, y = v * fh - 0.5f;
    float x0 = std::floor(x), y0 = std::floor(y), ax = x - x0, ay = y - y0;
    vec4 acc(0.0f);
    for (int k = 0; k < 4; k++) {
      float wgt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
      acc = acc + sampler_texel(s, x0 + (float)(k & 1), y0 + (float)(k >> 1), layer) * wgt;
    }
    return acc;
  }
  return sampler_texel(s, std::floor(u * fw), std::floor(v * fh), layer);
}
inline vec4 texture2D(const Sampler& s, const vec2& c) { return sampler_fetch(s, c.x, c.y, 0); }
inline vec4 texture(const Sampler& s, const vec2& c) { return sampler_fetch(s, c.x, c.y, 0); }
inline vec4 texture2DLod(const Sampler& s, const vec2& c, float) { return sampler_fetch(s, c.x, c.y, 0); }
inline vec4 texture2DArray(const Sampler& s, const vec3& c) { return sampler_fetch(s, c.x, c.y, (int)std::floor(c.z + 0.5f)); }
inline ivec2 textureSize(const Sampler& s, int) { ivec2 r; r.x = s.w; r.y = s.h; return r; }

struct LightModelProducts { vec4 sceneColor; };

}  // namespace glsl
