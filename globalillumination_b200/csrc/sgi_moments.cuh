// sgi_moments.cuh — pre-filtered ("moment") shadow maps of the ShadowMapping program: VSM, ESM, EVSM, MSM
// (SURVEY.md §8(f) row 4).  Included by sgi_shadow.cu (filter + reconstruction kernels) and sgi_raster.cu (the per-texel
// moment function the light-view resolve calls).
//
// Reference (ShadowMapping/):
//   moment render     Shaders/ShadowMap/Moments.frag:19-49, Exponential.frag:15-22, ExponentialMoments.frag:15-30
//                     (bound by displaySceneFromLightPOV, src/main.cpp:227-243)
//   quantisation      MyGLGeometryViewer::configureMoments, src/Viewers/MyGLGeometryViewer.cpp:188-213
//   separable blur    filterShadowMap, src/main.cpp:374-398; Shaders/Filter/GaussianFilter.frag:10-34,
//                     LogGaussianFilter.frag:10-49; weights Filter::buildGaussianKernel, src/Filter.cpp:17-46
//   reconstruction    Shaders/Shadow.frag:118-220 + main :240-273
//
// Texture lookups of the chain are bilinear filtering of level 0 (fp32 weights fract(u*size - 0.5), texels summed in the
// order 00, 10, 01, 11, border colour 0) and dFdx / dFdy are the fine 2x2-quad differences on the fragment's own plane:
// the two points GL leaves to the implementation, fixed in DESIGN.md §2.  fp32 in source order (-fmad=false): everything
// without exp / log is bit-identical to oracle/; the exponential paths agree to a few ulp of expf / logf.
#pragma once
#include <climits>
#include <cstddef>
#include "sgi_internal.cuh"

#define SGI_MOM_MAX_ORDER 33            // `uniform float kernel[33]`, GaussianFilter.frag:8

__host__ __device__ __forceinline__ bool sgi_is_moment_tech(int t) { return t >= SGI_TECH_VSM && t <= SGI_TECH_MSM; }

// linearize(): Moments.frag:9-16 == Shadow.frag:32-39
__device__ __forceinline__ float mom_linearize(float depth, int z_near, int z_far) {
  const float n = (float)z_near, f = (float)z_far;
  return (2.0f * n) / (f + n - depth * (f - n));
}

struct MomQuant { float m[16], minv[16], t[4]; };

// One texel of the moment target from the window depth of the winning fragment and of its plane at the quad partners
// (x^1, y) and (x, y^1).
__device__ __forceinline__ float4 mom_texel(int tech, float zwin, float zwin_px, float zwin_py, int x_odd, int y_odd, int z_near,
                                            int z_far, const float* __restrict__ q, const float* __restrict__ qt) {
  const float depth = mom_linearize(zwin, z_near, z_far);
  if (tech == SGI_TECH_ESM) return make_float4(depth, 0.0f, 0.0f, 1.0f);                     // Exponential.frag:21
  const float m0 = depth;
  float m1 = depth * depth;
  if (tech == SGI_TECH_VSM || tech == SGI_TECH_EVSM) {
    const float dpx = mom_linearize(zwin_px, z_near, z_far), dpy = mom_linearize(zwin_py, z_near, z_far);
    const float dx = x_odd ? depth - dpx : dpx - depth, dy = y_odd ? depth - dpy : dpy - depth;
    m1 = m1 + 0.25f * (dx * dx + dy * dy);                                                   // Moments.frag:36
    return tech == SGI_TECH_VSM ? make_float4(m0, m1, 0.0f, 0.0f) : make_float4(m0, m1, depth, 1.0f);
  }
  const float m2 = depth * depth * depth, m3 = depth * depth * depth * depth;                // Moments.frag:40-46
  float o[4];
#pragma unroll
  for (int r = 0; r < 4; r++) o[r] = (((q[0 + r] * m0 + q[4 + r] * m1) + q[8 + r] * m2) + q[12 + r] * m3) + qt[r];
  return make_float4(o[0], o[1], o[2], o[3]);
}

#ifdef SGI_MOMENTS_KERNELS
namespace {

// MyGLGeometryViewer.cpp:193-199: the 16 numbers as typed, glm::transpose, glm::inverse (func_matrix.inl:530-587) in fp32
void mom_quantization(MomQuant& Q) {
  static const float typed[4][4] = {
      {-2.07224649f, 32.2370378f, -68.5710746f, 39.3703274f},
      {13.7948857f, -59.4683976f, 82.035975f, -35.3649032f},
      {0.105877704f, -1.90774663f, 9.34965551f, -6.65434907f},
      {9.79240621f, -33.76521106f, 47.9456097f, -23.9728048f}};
  float* m = Q.m;
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) m[c * 4 + r] = typed[r][c];
  Q.t[0] = 0.0359558848f; Q.t[1] = 0.0f; Q.t[2] = 0.0f; Q.t[3] = 0.0f;                       // :208
  auto M = [&](int c, int r) -> float { return m[c * 4 + r]; };
  // volatile: each product / difference is rounded to fp32 on its own, whatever the host compiler would like to fuse
  volatile float s[18];
  const int pairs[18][4] = {   // sub-determinant M(a,b)*M(c,d) - M(c,b)*M(a,d) with columns (a,c) and rows (b,d)
      {2, 2, 3, 3}, {1, 2, 3, 3}, {1, 2, 2, 3}, {2, 1, 3, 3}, {1, 1, 3, 3}, {1, 1, 2, 3}, {2, 1, 3, 2}, {1, 1, 3, 2}, {1, 1, 2, 2},
      {2, 0, 3, 3}, {1, 0, 3, 3}, {1, 0, 2, 3}, {2, 0, 3, 2}, {1, 0, 3, 2}, {1, 0, 2, 2}, {2, 0, 3, 1}, {1, 0, 3, 1}, {1, 0, 2, 1}};
  for (int k = 0; k < 18; k++) {
    volatile float a = M(pairs[k][0], pairs[k][1]) * M(pairs[k][2], pairs[k][3]);
    volatile float b = M(pairs[k][2], pairs[k][1]) * M(pairs[k][0], pairs[k][3]);
    s[k] = a - b;
  }
  // Fac0..Fac5 = (s0,s0,s1,s2), (s3,s3,s4,s5), (s6,s6,s7,s8), (s9,s9,s10,s11), (s12,s12,s13,s14), (s15,s15,s16,s17)
  auto F = [&](int f, int k) -> float { return s[3 * f + (k == 0 ? 0 : k - 1)]; };
  auto V = [&](int row, int k) -> float { return k == 0 ? M(1, row) : M(0, row); };
  const float sa[4] = {1.0f, -1.0f, 1.0f, -1.0f}, sb[4] = {-1.0f, 1.0f, -1.0f, 1.0f};
  float inv[16];
  auto comb = [&](float a0, float a1, float b0, float b1, float c0, float c1) -> float {
    volatile float p0 = a0 * a1, p1 = b0 * b1, p2 = c0 * c1;
    volatile float d = p0 - p1;
    volatile float e = d + p2;
    return e;
  };
  for (int k = 0; k < 4; k++) {
    inv[0 * 4 + k] = sa[k] * comb(V(1, k), F(0, k), V(2, k), F(1, k), V(3, k), F(2, k));
    inv[1 * 4 + k] = sb[k] * comb(V(0, k), F(0, k), V(2, k), F(3, k), V(3, k), F(4, k));
    inv[2 * 4 + k] = sa[k] * comb(V(0, k), F(1, k), V(1, k), F(3, k), V(3, k), F(5, k));
    inv[3 * 4 + k] = sb[k] * comb(V(0, k), F(2, k), V(1, k), F(4, k), V(2, k), F(5, k));
  }
  volatile float d0 = M(0, 0) * inv[0], d1 = M(0, 1) * inv[4], d2 = M(0, 2) * inv[8], d3 = M(0, 3) * inv[12];
  volatile float det = d0 + d1;
  det = det + d2;
  det = det + d3;
  for (int k = 0; k < 16; k++) { volatile float v = inv[k] / det; Q.minv[k] = v; }
}

// Filter::buildGaussianKernel, Filter.cpp:17-46
void mom_gaussian_kernel(int order, float* kernel) {
  const float norm = powf(2.0f, (float)(order - 1));
  int coef = 1;
  for (int j = 0; j < order; j++) {
    if (j > 0) coef = coef * (order - 1 - j + 1) / j;
    volatile float w = (float)coef / norm;
    kernel[j] = w;
  }
}

// GL_LINEAR of level 0, CLAMP_TO_BORDER (0,0,0,0), RGBA32F.  MomTap = where a bilinear sample lands: the four texels are
// (ix, iy), (ix+1, iy), (ix, iy+1), (ix+1, iy+1); a false c* flag means that column / row is outside the image (border colour).
struct MomTap {
  int ix, iy; bool cx0, cx1, cy0, cy1;
  float w00, w10, w01, w11;
};
__device__ __forceinline__ MomTap mom_tap(float fw, float fh, float u, float v) {
  MomTap t;
  const float x = u * fw - 0.5f, y = v * fh - 0.5f;
  const float x0 = floorf(x), y0 = floorf(y), ax = x - x0, ay = y - y0;
  const float bx = 1.0f - ax, by = 1.0f - ay;
  t.cx0 = x0 >= 0.0f && x0 < fw; t.cx1 = x0 + 1.0f >= 0.0f && x0 + 1.0f < fw;
  t.cy0 = y0 >= 0.0f && y0 < fh; t.cy1 = y0 + 1.0f >= 0.0f && y0 + 1.0f < fh;
  // NaN coordinates fail every range test: all four texels are the border colour
  t.ix = t.cx0 ? (int)x0 : (t.cx1 ? -1 : INT_MIN / 2); t.iy = t.cy0 ? (int)y0 : (t.cy1 ? -1 : INT_MIN / 2);
  t.w00 = bx * by; t.w10 = ax * by; t.w01 = bx * ay; t.w11 = ax * ay;
  return t;
}
__device__ __forceinline__ float4 mom_texel_at(const float4* __restrict__ img, int w, int ix, int iy, bool inside) {
  return inside ? __ldg(img + ((ptrdiff_t)iy * w + ix)) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ float4 mom_blend(const MomTap& t, float4 t00, float4 t10, float4 t01, float4 t11) {
  float4 r;
  r.x = (((0.0f + t00.x * t.w00) + t10.x * t.w10) + t01.x * t.w01) + t11.x * t.w11;
  r.y = (((0.0f + t00.y * t.w00) + t10.y * t.w10) + t01.y * t.w01) + t11.y * t.w11;
  r.z = (((0.0f + t00.z * t.w00) + t10.z * t.w10) + t01.z * t.w01) + t11.z * t.w11;
  r.w = (((0.0f + t00.w * t.w00) + t10.w * t.w10) + t01.w * t.w01) + t11.w * t.w11;
  return r;
}
__device__ __forceinline__ float4 mom_fetch4(const float4* __restrict__ img, int w, int h, float fw, float fh, float u, float v) {
  const MomTap t = mom_tap(fw, fh, u, v);
  return mom_blend(t, mom_texel_at(img, w, t.ix, t.iy, t.cx0 && t.cy0), mom_texel_at(img, w, t.ix + 1, t.iy, t.cx1 && t.cy0),
                   mom_texel_at(img, w, t.ix, t.iy + 1, t.cx0 && t.cy1), mom_texel_at(img, w, t.ix + 1, t.iy + 1, t.cx1 && t.cy1));
}

struct MomFilterArgs {
  const float4* src; int sw, sh; float4* dst; int W, H;
  int order; float kernel[SGI_MOM_MAX_ORDER];
  float step_s, step_t;
};

// One pass of filterShadowMap (main.cpp:380-392): a W x H target over the full-screen quad of GaussianFilter.vert:4-9; the
// source is read with the TARGET's step (drawTextureOnShader's imageWidth / imageHeight).  One thread per target texel;
// the taps of neighbouring threads overlap almost entirely, so the source is served from L1 and crosses HBM once.
// LOGSPACE: LogGaussianFilter.frag (ESM) on .x, replicated into the four channels.
// ORDER: compile-time tap count of the specialised variant (7 = the reference's default, main.cpp:859: all 28 texel loads of a
// target texel are independent and issue together), 0 = run-time loop.
template <bool HORIZONTAL, bool LOGSPACE, int ORDER>
__global__ void __launch_bounds__(256) k_mom_filter(const MomFilterArgs a) {
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
  if (i >= a.W || j >= a.H) return;
  const float nx = ((float)i + 0.5f) / (float)a.W * 2.0f - 1.0f, ny = ((float)j + 0.5f) / (float)a.H * 2.0f - 1.0f;
  const float cs = nx * 0.5f + 0.5f, ct = ny * 0.5f + 0.5f;
  const float dir_s = HORIZONTAL ? 1.0f : 0.0f, dir_t = HORIZONTAL ? 0.0f : 1.0f;
  const float fw = (float)a.sw, fh = (float)a.sh;
  const int order = ORDER ? ORDER : a.order;
  const int kc = order / 2;
  float4 out;
  if (!LOGSPACE) {
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = -kc; s <= kc; s++) {
      const float4 t = mom_fetch4(a.src, a.sw, a.sh, fw, fh, cs + dir_s * (float)s * a.step_s, ct + dir_t * (float)s * a.step_t);
      const float k = a.kernel[kc + s];
      sum.x = sum.x + t.x * k; sum.y = sum.y + t.y * k; sum.z = sum.z + t.z * k; sum.w = sum.w + t.w * k;
    }
    out = sum;
  } else {
    int ks = -kc;
    const float s0 = mom_fetch4(a.src, a.sw, a.sh, fw, fh, cs + dir_s * (float)ks * a.step_s, ct + dir_t * (float)ks * a.step_t).x;
    ks++;
    const float s1 = mom_fetch4(a.src, a.sw, a.sh, fw, fh, cs + dir_s * (float)ks * a.step_s, ct + dir_t * (float)ks * a.step_t).x;
    float sum = s0 + logf(a.kernel[0] + (a.kernel[1] * expf(s1 - s0)));                      // log_conv, :10-13
#pragma unroll
    for (int k = 2; k < order; k++) {
      ks++;
      const float sk = mom_fetch4(a.src, a.sw, a.sh, fw, fh, cs + dir_s * (float)ks * a.step_s, ct + dir_t * (float)ks * a.step_t).x;
      sum = sum + logf(1.0f + (a.kernel[k] * expf(sk - sum)));
    }
    out = make_float4(sum, sum, sum, sum);
  }
  a.dst[(size_t)j * a.W + i] = out;
}

// Shadow.frag:118-132
__device__ __forceinline__ float mom_chebyshev(float m0, float m1, float z, float si) {
  if (z <= m0) return 1.0f;
  const float variance = m1 - (m0 * m0);
  const float d = z - m0;
  float p_max = variance / (variance + d * d);
  const float p = (z <= m0) ? 1.0f : 0.0f;
  p_max = g_max(p, p_max);
  return g_mix(p_max, 1.0f, si);
}
__device__ __forceinline__ float mom_clamp(float x, float lo, float hi) { return g_min(g_max(x, lo), hi); }

// Shadow.frag:168-220
__device__ __forceinline__ float mom_hamburger(float4 bq, float zx, float si, const float* __restrict__ qi, const float* __restrict__ qt) {
  const float v0 = bq.x - qt[0], v1 = bq.y - qt[1], v2 = bq.z - qt[2], v3 = bq.w - qt[3];
  float b[4];
#pragma unroll
  for (int r = 0; r < 4; r++) b[r] = ((qi[0 + r] * v0 + qi[4 + r] * v1) + qi[8 + r] * v2) + qi[12 + r] * v3;
  const float bias = 0.00003f;
#pragma unroll
  for (int r = 0; r < 4; r++) b[r] = (1.0f - bias) * b[r] + bias * 0.5f;
  const float d0 = 1.0f, d1 = zx, d2 = zx * zx;
  const float L10 = b[0], L20 = b[1];
  const float D11 = b[1] - L10 * L10;
  const float L21 = (b[2] - L20 * L10) / D11;
  const float D22 = b[3] - L20 * L20 - L21 * L21 * D11;
  const float y0 = d0;
  float y1 = d1 - L10 * y0;
  float y2 = d2 - L20 * y0 - L21 * y1;
  y1 = y1 / D11; y2 = y2 / D22;
  const float cz = y2, cy = y1 - L21 * cz, cx = y0 - L10 * cy - L20 * cz;
  const float p = cy / cz, qq = cx / cz;
  const float D = ((p * p) / 4.0f) - qq;
  const float r = sqrtf(D);
  const float zy = -(p / 2.0f) - r, zz = -(p / 2.0f) + r;
  if (zx <= zy) return 1.0f;
  else if (zx <= zz)
    return mom_clamp((1.0f - mom_clamp((zx * zz - b[0] * (zx + zz) + b[1]) / ((zz - zy) * (zx - zy)), 0.0f, 1.0f)), si, 1.0f);
  else
    return mom_clamp((1.0f - mom_clamp(1.0f - (zy * zz - b[0] * (zy + zz) + b[1]) / ((zx - zy) * (zx - zz)), 0.0f, 1.0f)), si, 1.0f);
}

struct MomVisArgs {
  float mv[16], nm[9], lpos[3], lmvp[16];
  float shadow_intensity; int z_near, z_far;
  const float4* pos4; const float4* nrm4; float* vis;
  int W, H, rx0, ry0, rx1, ry1;
  const float4* fmap; int mw, mh;
  float minv[16], qt[4];
};

// Shadow.frag main (:240-273) with VSM / ESM / EVSM / MSM == 1 over the twice-filtered map (FILTER_Y_MAP_COLOR, main.cpp:316)
template <int TECH>
__global__ void __launch_bounds__(256) k_mom_visibility(const MomVisArgs a) {
  const int x = a.rx0 + blockIdx.x * 32 + threadIdx.x, y = a.ry0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= a.rx1 || y >= a.ry1) return;
  const size_t o = (size_t)y * a.W + x;
  const float4 vertex = __ldg(&a.pos4[o]);
  if (vertex.x == 0.0f) { a.vis[o] = 0.0f; return; }
  const float4 normal = __ldg(&a.nrm4[o]);
  const float4 sc = mat4_mul(a.lmvp, vertex);
  float4 c;
  sgi_div3(sc.x, sc.y, sc.z, sc.w, c.x, c.y, c.z);
  c.w = sc.w / sc.w;
  const float si = a.shadow_intensity;
  float shadow;
  {
    // computePreEvaluationBasedOnNormalOrientation, Shadow.frag:222-238 (same arithmetic as pre_evaluation())
    const float4 ev = mat4_mul(a.mv, vertex);
    float n0 = (a.nm[0] * normal.x + a.nm[3] * normal.y) + a.nm[6] * normal.z;
    float n1 = (a.nm[1] * normal.x + a.nm[4] * normal.y) + a.nm[7] * normal.z;
    float n2 = (a.nm[2] * normal.x + a.nm[5] * normal.y) + a.nm[8] * normal.z;
    const float inv = 1.0f / sqrtf((n0 * n0 + n1 * n1) + n2 * n2);
    n0 = n0 * inv; n1 = n1 * inv; n2 = n2 * inv;
    const float d0 = a.lpos[0] - ev.x, d1 = a.lpos[1] - ev.y, d2 = a.lpos[2] - ev.z;
    const float invl = 1.0f / sqrtf((d0 * d0 + d1 * d1) + d2 * d2);
    const float L0 = d0 * invl, L1 = d1 * invl, L2 = d2 * invl;
    if (!(normal.w != 0.0f)) { n0 *= -1.0f; n1 *= -1.0f; n2 *= -1.0f; }
    const float dt = (n0 * L0 + n1 * L1) + n2 * L2;
    shadow = (g_max(dt, 0.0f) == 0.0f) ? si : 1.0f;
  }
  if (sc.w > 0.0f && shadow == 1.0f) {
    const float4 b = mom_fetch4(a.fmap, a.mw, a.mh, (float)a.mw, (float)a.mh, c.x, c.y);
    const float z = mom_linearize(c.z, a.z_near, a.z_far);
    if (TECH == SGI_TECH_VSM) shadow = mom_chebyshev(b.x, b.y, z, si);                         // :134-141
    else if (TECH == SGI_TECH_ESM) {                                                          // :144-157
      const float e2 = expf(80.0f * b.x);
      const float e1 = expf(-80.0f * z);
      shadow = mom_clamp(e1 * e2, si, 1.0f);
    } else if (TECH == SGI_TECH_EVSM) {                                                       // :160-175
      const float variance = mom_chebyshev(b.x, b.y, z, si);
      const float e1 = expf(-60.0f * z);
      const float e2 = expf(60.0f * b.z);
      shadow = g_min(variance, mom_clamp(e1 * e2, si, 1.0f));
    } else shadow = mom_hamburger(b, z, si, a.minv, a.qt);
  }
  a.vis[o] = shadow;
}

}  // namespace
#endif  // SGI_MOMENTS_KERNELS
