// sgi_shadow.cu — per-pixel shadow test / filter kernels for sm_100a (K3a-d, K5).  One thread per screen
// pixel reads the G-buffer (2 x 128-bit loads), projects into light space and evaluates the technique the
// reference's full-screen fragment program would:
//   hard, PCF            ShadowMapping/Shaders/Shadow.frag:86-116,222-273
//   PCSS                 SoftShadowMapping/Shaders/SoftShadow/PlausibleSoftShadow.frag:33-49,166-194,365-398,556-563,605-633
//   RBSM / RPCF / RSMSS  ShadowMapping/Shaders/RBSM/{NonConservativeSMSR,ConservativeSMSR,FilteredRBSM}.frag
//   many-light           SoftShadowMapping/Shaders/SoftShadow/AccurateSoftShadow.frag:52-133
// Shadow-map taps are GL_NEAREST + CLAMP_TO_BORDER(0): texel = floor(coord*size) (MyGLTextureViewer.cpp:3-28).
// fp32 in source order, -fmad=false: results are bit-identical to oracle/ (DESIGN.md §3).
#include <cstdlib>
#include <cstring>
#include <vector>
#include "sgi_internal.cuh"

namespace {

struct VisArgs {
  sgi_params p;
  float mv[16], nm[9], lpos[3];
  float lmvp[16];
  const float4* pos4; const float4* nrm4; float* vis;
  int W, H, rx0, ry0, rx1, ry1;
  const float* sm; int SW, SH; float fw, fh;
  float sx, sy;
  float pcf_off[SGI_MAX_PCF_TAPS]; int pcf_n;
  float rpcf_off[SGI_MAX_PCF_TAPS]; int rpcf_n;
  // pixel-independent parts of the tap coordinates, evaluated once on the host with the shader's own fp32 operations:
  float pcf_du[SGI_MAX_PCF_TAPS], pcf_dv[SGI_MAX_PCF_TAPS];   // pcf_off[k]*incrWidth, pcf_off[k]*incrHeight   (Shadow.frag:104)
  float bs_q[SGI_MAX_PCF_TAPS]; int bs_w0, bs_n;              // (float(w)*blockerSearchWidth)/filterWidth      (PlausibleSoftShadow.frag:180)
  const float4* trans; int N; size_t layer;
  int pcss_early_out;      // option "pcss_early_out" (validated on the host): see pcss_t
  // min-max cull: extrema of the depth map over the reach of the tap window, per 32x32-texel block of the window's centre texel
  const float* dmin; const float* dmax; int mm_w; float mm_limit;     // mm_limit: largest reach (texels) the dilation covers
  float pcf_reach, bs_reach;                                          // reach of the PCF grid / the PCSS blocker search, in texels
  // multi_fused: the camera pass's primitive ids and raster records (positions are resolved here instead of read from a G-buffer)
  const unsigned int* ids; const SgiRec* rec; const SgiRecAttr* attr; const int32_t* ovf_base;
  // multi_partial == 2: the lights this rank sampled, one bit each at the light's index in the whole set, 8 lights per byte plane,
  // laid out [rank strip][plane][strip pixels] so that every rank's part is contiguous (one in-place reduce-scatter)
  unsigned char* mask; int mask_planes, mask_rows; size_t mask_strip; const int* gid; int mask_total;
  int own_r0, own_r1;                                                 // rows of this rank's strip (they receive the discard decision)
};

struct Smap { const float* __restrict__ d; int w, h; float fw, fh; };

__device__ __forceinline__ float sm_fetch(const Smap& s, float u, float v) {
  float fi = floorf(u * s.fw), fj = floorf(v * s.fh);
  if (!(fi >= 0.0f && fi < s.fw && fj >= 0.0f && fj < s.fh)) return 0.0f;
  return __ldg(&s.d[(size_t)(int)fj * s.w + (int)fi]);
}
__device__ __forceinline__ float g_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float g_max(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float g_min(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float g_fract(float x) { return x - floorf(x); }

__device__ __forceinline__ float4 mat4_mul(const float* __restrict__ m, float4 v) {
  float4 r;
  r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
  r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
  r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
  r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
  return r;
}

// Shadow.frag:222-238
__device__ __forceinline__ float pre_evaluation(const VisArgs& a, float4 vertex, float4 normal) {
  float4 ev = mat4_mul(a.mv, vertex);
  float n0 = (a.nm[0] * normal.x + a.nm[3] * normal.y) + a.nm[6] * normal.z;
  float n1 = (a.nm[1] * normal.x + a.nm[4] * normal.y) + a.nm[7] * normal.z;
  float n2 = (a.nm[2] * normal.x + a.nm[5] * normal.y) + a.nm[8] * normal.z;
  float inv = 1.0f / sqrtf((n0 * n0 + n1 * n1) + n2 * n2);
  n0 = n0 * inv; n1 = n1 * inv; n2 = n2 * inv;
  float d0 = a.lpos[0] - ev.x, d1 = a.lpos[1] - ev.y, d2 = a.lpos[2] - ev.z;
  float invl = 1.0f / sqrtf((d0 * d0 + d1 * d1) + d2 * d2);
  float L0 = d0 * invl, L1 = d1 * invl, L2 = d2 * invl;
  if (!(normal.w != 0.0f)) { n0 *= -1.0f; n1 *= -1.0f; n2 *= -1.0f; }
  float dt = (n0 * L0 + n1 * L1) + n2 * L2;
  return (g_max(dt, 0.0f) == 0.0f) ? a.p.shadow_intensity : 1.0f;
}

// Separable addressing of a tap grid: the texel column of a tap depends only on its u, the row only on its v
// (texel = floor(coord*size), A.3), so both are computed once per axis position instead of once per tap; a tap is
// then one add, one load and one compare.  Values are identical to calling sm_fetch per tap.
__device__ __forceinline__ int axis_texel(float coord, float size) {      // texel index, or -1 outside [0,size) / NaN
  float f = floorf(coord * size);
  return (f >= 0.0f && f < size) ? (int)f : -1;
}

// Where taps are read from: the depth map in global memory (through L1/L2), or a window of it that the CTA staged
// in shared memory with 128-bit loads.  Rows/columns are turned into keys once per axis position: key < 0 = border.
template <bool SHARED>
struct TapSrc {
  const float* base; int pitch, x0, y0;
  __device__ __forceinline__ int rowkey(int r) const { return r < 0 ? -1 : (r - y0) * pitch; }
  __device__ __forceinline__ int colkey(int c) const { return c < 0 ? -1 : c - x0; }
  __device__ __forceinline__ float tap(int rk, int ck) const {
    if ((rk | ck) < 0) return 0.0f;
    return SHARED ? base[rk + ck] : __ldg(&base[(size_t)rk + ck]);
  }
  // Row-wise form of the same taps: when the row and every column of the grid are inside the map (all but the pixels whose
  // footprint crosses the map border), a tap is one address add and one load off the row pointer - no border test, no 64-bit
  // index arithmetic per tap.  Same texels, same values.
  // (the element offset is formed in 32 bits - it is below S*S <= 2^26 - and widened once by the address multiply-add)
  __device__ __forceinline__ int rowptr(int rk) const { return rk; }
  __device__ __forceinline__ float tap_in(int rk, int ck) const { const unsigned off = (unsigned)(rk + ck); return SHARED ? base[off] : __ldg(base + off); }
};
template <int N>
__device__ __forceinline__ int keys_or(const int (&k)[N], int n) {
  int m = 0;
#pragma unroll
  for (int i = 0; i < N; i++) if (i < n) m |= k[i];
  return m;
}

// Min-max cull.  dmin / dmax hold, for the 32x32-texel block of the window's centre texel, the smallest / largest depth of every
// texel a tap window of reach <= mm_limit around that texel can touch (taps beyond the map edge read 0 and are folded into the
// minimum).  z <= dmin: every tap compares lit; z > dmax: every tap compares shadowed - the taps' values are not needed, only
// their number, and the result is the one the tap loop would produce, bit for bit.  Returns 0 (undecided), 1 (all lit), 2 (all shadowed).
__device__ __forceinline__ int mm_classify(const VisArgs& a, const Smap& s, float4 c, float reach) {
  if (!a.dmin || !(reach <= a.mm_limit)) return 0;
  const int tx = axis_texel(c.x, s.fw), ty = axis_texel(c.y, s.fh);
  if ((tx | ty) < 0) return 0;
  const int b = (ty >> 5) * a.mm_w + (tx >> 5);
  if (c.z <= __ldg(&a.dmin[b])) return 1;
  if (c.z > __ldg(&a.dmax[b])) return 2;
  return 0;
}
// what `count` taps that all compare shadowed add up to, in the shaders' own order of additions
__device__ __forceinline__ float sum_shadowed(float si, int count) {
  float illum = 0.0f;
  for (int i = 0; i < count; i++) illum += si;
  return illum;
}

// ---- Shadow.frag:86-116 (tap offsets precomputed on the host with the same fp32 loop) ----
// N = taps per axis known at compile time (0 = run-time count, columns kept in local memory)
template <int N, bool SHARED>
__device__ __forceinline__ float pcf_t(const VisArgs& a, const Smap& s, const TapSrc<SHARED>& src, float4 c) {
  const int n = N ? N : a.pcf_n;
  if (n <= 0) return 1.0f;
  if (!SHARED) {
    const int cls = mm_classify(a, s, c, a.pcf_reach);
    if (cls == 1) return (float)(n * n) / (float)(n * n);             // n*n additions of 1.0 are exact
    if (cls == 2) return sum_shadowed(a.p.shadow_intensity, n * n) / (float)(n * n);
  }
  int rows[N ? N : SGI_MAX_PCF_TAPS];
#pragma unroll
  for (int ih = 0; ih < (N ? N : SGI_MAX_PCF_TAPS); ih++)
    if (ih < n) rows[ih] = src.rowkey(axis_texel(c.y + a.pcf_dv[ih], s.fh));
  float illum = 0.0f;
  const int rows_or = keys_or(rows, n);
  for (int iw = 0; iw < n; iw++) {                     // Shadow.frag:98-99: w outer, h inner
    const int col = src.colkey(axis_texel(c.x + a.pcf_du[iw], s.fw));
    if ((rows_or | col) >= 0) {
      const int cp = src.rowptr(col);     // base + column; the row keys are the offsets
#pragma unroll
      for (int ih = 0; ih < (N ? N : SGI_MAX_PCF_TAPS); ih++)
        if (ih < n) { if (c.z <= src.tap_in(cp, rows[ih])) illum += 1.0f; else illum += a.p.shadow_intensity; }
    } else {
#pragma unroll
      for (int ih = 0; ih < (N ? N : SGI_MAX_PCF_TAPS); ih++)
        if (ih < n) { if (c.z <= src.tap(rows[ih], col)) illum += 1.0f; else illum += a.p.shadow_intensity; }
    }
  }
  return illum / (float)(n * n);
}

// ---- Shadow.frag:41-84: cubic() weights and textureBicubic() on the NEAREST depth texture, .z; :86-116 with tricubicPCF == 1 ----
__device__ __forceinline__ void cubic4(float v, float (&o)[4]) {
  const float n0 = 1.0f - v, n1 = 2.0f - v, n2 = 3.0f - v, n3 = 4.0f - v;
  const float s0 = n0 * n0 * n0, s1 = n1 * n1 * n1, s2 = n2 * n2 * n2, s3 = n3 * n3 * n3;
  (void)s3;
  const float x = s0;
  const float y = s1 - 4.0f * s0;
  const float z = s2 - 4.0f * s1 + 6.0f * s0;
  const float w = 6.0f - x - y - z;
  const float sixth = 1.0f / 6.0f;
  o[0] = x * sixth; o[1] = y * sixth; o[2] = z * sixth; o[3] = w * sixth;
}
__device__ __forceinline__ float texture_bicubic_z(const Smap& sm, float u, float v) {
  const float invx = 1.0f / sm.fw, invy = 1.0f / sm.fh;
  float tx = u * sm.fw - 0.5f, ty = v * sm.fh - 0.5f;
  const float fx = g_fract(tx), fy = g_fract(ty);
  tx -= fx; ty -= fy;
  float xc[4], yc[4];
  cubic4(fx, xc); cubic4(fy, yc);
  const float c0 = tx + -0.5f, c1 = tx + 1.5f, c2 = ty + -0.5f, c3 = ty + 1.5f;
  const float s0 = xc[0] + xc[1], s1 = xc[2] + xc[3], s2 = yc[0] + yc[1], s3 = yc[2] + yc[3];
  float o0 = c0 + xc[1] / s0, o1 = c1 + xc[3] / s1, o2 = c2 + yc[1] / s2, o3 = c3 + yc[3] / s3;
  o0 *= invx; o1 *= invx; o2 *= invy; o3 *= invy;
  const float sample0 = sm_fetch(sm, o0, o2), sample1 = sm_fetch(sm, o1, o2), sample2 = sm_fetch(sm, o0, o3), sample3 = sm_fetch(sm, o1, o3);
  const float sx = s0 / (s0 + s1), sy = s2 / (s2 + s3);
  return g_mix(g_mix(sample3, sample2, sx), g_mix(sample1, sample0, sx), sy);
}
__device__ __forceinline__ float pcf_tricubic(const VisArgs& a, const Smap& s, float4 c) {
  const int n = a.pcf_n;
  if (n <= 0) return 1.0f;
  float illum = 0.0f;
  for (int iw = 0; iw < n; iw++)                       // Shadow.frag:98-99: w outer, h inner
    for (int ih = 0; ih < n; ih++) {
      const float dfl = texture_bicubic_z(s, c.x + a.pcf_du[iw], c.y + a.pcf_dv[ih]);
      if (c.z <= dfl) illum += 1.0f; else illum += a.p.shadow_intensity;
    }
  return illum / (float)(n * n);
}

// ---- PlausibleSoftShadow.frag:166-194: mean depth of the blockers (1.0 if none) ----
template <int NB, bool SHARED>
__device__ __forceinline__ float pcss_blockers(const VisArgs& a, const Smap& s, const TapSrc<SHARED>& src, float4 c) {
  float averageDepth = 0.0f;
  int numberOfBlockers = 0;
  // `for(int w = -filterWidth; w <= filterWidth; w++)`: the int start truncates toward zero (A.7).  The offsets
  // (float(w)*blockerSearchWidth)/filterWidth do not depend on the pixel: a.bs_q[] holds them (same fp32 ops, host)
  const int nb = NB ? NB : a.bs_n;
  int cols[NB ? NB : SGI_MAX_PCF_TAPS];
#pragma unroll
  for (int k = 0; k < (NB ? NB : SGI_MAX_PCF_TAPS); k++)
    if (k < nb) cols[k] = src.colkey(axis_texel(c.x + a.bs_q[k], s.fw));
  const int cols_or = keys_or(cols, nb);
#pragma unroll
  for (int j = 0; j < nb; j++) {
    const int row = src.rowkey(axis_texel(c.y + a.bs_q[j], s.fh));
    if ((cols_or | row) >= 0) {
      const int rp = src.rowptr(row);
      if (NB > 0 && NB <= 8) {
        // specialised grid: the row's blockers are collected as bits and counted with one popc instead of one add per tap
        unsigned int hit = 0u;
#pragma unroll
        for (int k = 0; k < (NB ? NB : 1); k++) {
          const float dfl = src.tap_in(rp, cols[k]);
          if (c.z > dfl) { averageDepth += dfl; hit |= 1u << k; }
        }
        numberOfBlockers += __popc(hit);
      } else {
#pragma unroll
        for (int k = 0; k < (NB ? NB : SGI_MAX_PCF_TAPS); k++)
          if (k < nb) {
            const float dfl = src.tap_in(rp, cols[k]);
            if (c.z > dfl) { averageDepth += dfl; numberOfBlockers++; }
          }
      }
    } else {
#pragma unroll
      for (int k = 0; k < (NB ? NB : SGI_MAX_PCF_TAPS); k++)
        if (k < nb) {
          const float dfl = src.tap(row, cols[k]);
          if (c.z > dfl) { averageDepth += dfl; numberOfBlockers++; }
        }
    }
  }
  if (numberOfBlockers == 0) return 1.0f;
  return averageDepth / (float)numberOfBlockers;
}

// ---- PlausibleSoftShadow.frag:365-374 ----
__device__ __forceinline__ float pcss_penumbra(const sgi_params& p, float averageDepth, float z) {
  if (averageDepth < 0.99f) return 0.0f;
  float pw = ((z - averageDepth) / averageDepth) * (float)p.light_source_radius;
  return ((float)p.z_near * pw) / z;
}

// ---- PlausibleSoftShadow.frag:376-398 (the caller has checked 0 < stepSize < 1) ----
template <int NK, bool SHARED>
__device__ __forceinline__ float pcss_filter(const VisArgs& a, const Smap& s, const TapSrc<SHARED>& src, float4 c, float penumbraWidth) {
  const sgi_params& p = a.p;
  const float fw2 = ((float)p.kernel_size - 1.0f) * 0.5f;
  float illum = 0.0f;
  const int w0 = (int)(-fw2);
  const int nk = NK ? NK : ((fw2 >= 0.0f) ? (int)fw2 - w0 + 1 : 0);
  int cols[NK ? NK : SGI_MAX_PCF_TAPS];
#pragma unroll
  for (int k = 0; k < (NK ? NK : SGI_MAX_PCF_TAPS); k++)
    if (k < nk) cols[k] = src.colkey(axis_texel(c.x + ((float)(w0 + k) * penumbraWidth) / fw2, s.fw));
  const int cols_or = keys_or(cols, nk);
  for (int h = w0; (float)h <= fw2; h++) {
    const int row = src.rowkey(axis_texel(c.y + ((float)h * penumbraWidth) / fw2, s.fh));
    if ((cols_or | row) >= 0) {
      const int rp = src.rowptr(row);
#pragma unroll
      for (int k = 0; k < (NK ? NK : SGI_MAX_PCF_TAPS); k++)
        if (k < nk) { if (c.z <= src.tap_in(rp, cols[k])) illum += 1.0f; else illum += p.shadow_intensity; }
    } else {
#pragma unroll
      for (int k = 0; k < (NK ? NK : SGI_MAX_PCF_TAPS); k++)
        if (k < nk) { if (c.z <= src.tap(row, cols[k])) illum += 1.0f; else illum += p.shadow_intensity; }
    }
  }
  return illum / (float)(p.kernel_size * p.kernel_size);
}

// Option "pcss_early_out": for 0 < z < 0.989 the program's result is 1.0 whatever the shadow map holds, so no tap is needed.
// Proof, in the shader's own fp32 terms: every blocker depth d satisfies 0 <= d < z (the maps hold [0,1], border taps 0), so the
// fp32 mean of up to 64x64 of them is below z(1 + 4096 eps) < 0.98925 < 0.99, computePenumbraWidth (PlausibleSoftShadow.frag:368)
// returns 0 and the step size 0 ends the program with 1.0 (:386); without blockers the mean is 1, the width
// ((z - 1) / 1) * lightSourceRadius * zNear / z is <= 0 for z in (0, 1) and the non-negative parameters the host checked, and
// the step size <= 0 ends it with 1.0 as well.  (under a near light this is every pixel.)
#define SGI_PCSS_EARLY_Z 0.989f
template <int NB, int NK>
__device__ __forceinline__ float pcss_t(const VisArgs& a, const Smap& s, float4 c) {
  if (a.pcss_early_out && c.z > 0.0f && c.z < SGI_PCSS_EARLY_Z) return 1.0f;
  // no depth around the blocker-search window is nearer than the pixel: no blocker, the average is 1.0, the penumbra width
  // ((z - 1) / 1 ...) is not positive and the program returns 1.0 (PlausibleSoftShadow.frag:189-190,386)
  if (c.z > 0.0f && c.z <= 1.0f && a.p.light_source_radius >= 0 && a.p.z_near >= 0 && a.p.kernel_size > 0 && mm_classify(a, s, c, a.bs_reach) == 1) return 1.0f;
  const TapSrc<false> g = {s.d, s.w, 0, 0};
  const float avg = pcss_blockers<NB, false>(a, s, g, c);
  const float pw = pcss_penumbra(a.p, avg, c.z);
  const float stepSize = 2.0f * pw / (float)a.p.kernel_size;
  if (stepSize <= 0.0f || stepSize >= 1.0f) return 1.0f;
  {
    // the filter's window reaches |pw| (in map units) from the centre: decided as a whole where the extrema allow it
    const int cls = mm_classify(a, s, c, fabsf(pw) * fmaxf(s.fw, s.fh) + 2.0f);
    if (cls) {
      const float fw2 = ((float)a.p.kernel_size - 1.0f) * 0.5f;
      const int w0 = (int)(-fw2);
      const int nk = NK ? NK : ((fw2 >= 0.0f) ? (int)fw2 - w0 + 1 : 0);
      const float kk = (float)(a.p.kernel_size * a.p.kernel_size);
      return (cls == 1 ? (float)(nk * nk) : sum_shadowed(a.p.shadow_intensity, nk * nk)) / kk;
    }
  }
  return pcss_filter<NK, false>(a, s, g, c, pw);
}

// ---- shared-memory staging of the CTA's light-space footprint ---------------------------------------------------
#define SGI_STAGE_FLOATS 12288                 // 48 KB window

// texel index clamped into the map (for footprints only: taps outside the map never touch memory)
__device__ __forceinline__ int clamp_texel(float coord, float size, bool& bad) {
  float f = floorf(coord * size);
  if (!(f == f)) { bad = true; return 0; }
  f = fminf(fmaxf(f, 0.0f), size - 1.0f);
  return (int)f;
}

// CTA-wide min/max of the per-thread footprints (warp shuffles, then one shared-memory round across the 8 warps).
// Returns true if at least one thread takes part and the window (x0 rounded down, width rounded up to 4 texels)
// fits the staging buffer; then stages it with 128-bit loads.  All 256 threads must call this.
__device__ __forceinline__ bool stage_window(const Smap& s, float* smem, int* red, bool take, bool bad, int lx, int hx, int ly, int hy,
                                             TapSrc<true>& out) {
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.y * 32 + threadIdx.x, lane = threadIdx.x, warp = threadIdx.y;
  int mnx = take ? lx : 0x7fffffff, mxx = take ? hx : -1, mny = take ? ly : 0x7fffffff, mxy = take ? hy : -1;
  int anybad = (take && bad) ? 1 : 0;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    mnx = min(mnx, __shfl_xor_sync(full, mnx, d)); mxx = max(mxx, __shfl_xor_sync(full, mxx, d));
    mny = min(mny, __shfl_xor_sync(full, mny, d)); mxy = max(mxy, __shfl_xor_sync(full, mxy, d));
    anybad |= __shfl_xor_sync(full, anybad, d);
  }
  __syncthreads();                                   // previous users of `red` and of the staging buffer are done
  if (lane == 0) { red[warp * 5 + 0] = mnx; red[warp * 5 + 1] = mxx; red[warp * 5 + 2] = mny; red[warp * 5 + 3] = mxy; red[warp * 5 + 4] = anybad; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; w++) {
    mnx = min(mnx, red[w * 5 + 0]); mxx = max(mxx, red[w * 5 + 1]); mny = min(mny, red[w * 5 + 2]); mxy = max(mxy, red[w * 5 + 3]);
    anybad |= red[w * 5 + 4];
  }
  if (mxx < 0 || anybad || (s.w & 3)) return false;
  const int x0 = mnx & ~3, wv = ((mxx - x0 + 4) >> 2), hh = mxy - mny + 1;     // wv = window width in float4
  if (wv * 4 * hh > SGI_STAGE_FLOATS) return false;
  const float4* __restrict__ g = reinterpret_cast<const float4*>(s.d);
  float4* t4 = reinterpret_cast<float4*>(smem);
  const int gw = s.w >> 2, gx = x0 >> 2;
  for (int v = tid; v < wv * hh; v += 256) {
    const int r = v / wv, cv = v - r * wv;
    t4[v] = __ldg(&g[(size_t)(mny + r) * gw + gx + cv]);
  }
  __syncthreads();
  out.base = smem; out.pitch = wv * 4; out.x0 = x0; out.y0 = mny;
  return true;
}

// ================================ RBSM ===========================================================
struct Rb {
  const Smap& s; float sx, sy; float thr; int max_search; float si; int filtered; float SWf, SHf;
  float newDepth;
};

// NonConservativeSMSR.frag:23-54 / ConservativeSMSR.frag:23-48 (same fetch sequence)
__device__ __forceinline__ void getdisc4(Rb& r, float4 c, float dir[4]) {
  c.x -= r.sx;
  dir[0] = (c.z <= sm_fetch(r.s, c.x, c.y)) ? 1.0f : 0.0f;
  c.x += 2.0f * r.sx;
  dir[1] = (c.z <= sm_fetch(r.s, c.x, c.y)) ? 1.0f : 0.0f;
  c.x -= r.sx;
  c.y += r.sy;
  dir[2] = (c.z <= sm_fetch(r.s, c.x, c.y)) ? 1.0f : 0.0f;
  c.y -= 2.0f * r.sy;
  dir[3] = (c.z <= sm_fetch(r.s, c.x, c.y)) ? 1.0f : 0.0f;
}

// NonConservativeSMSR.frag:56-94
__device__ bool nc_getdisc_f(Rb& r, float4 c, float dx, float dy, float discType) {
  if (dx == 0.0f) {
    c.x -= r.sx;
    float left = (c.z <= sm_fetch(r.s, c.x, c.y)) ? 1.0f : 0.0f;
    if (fabsf(left - discType) == 0.0f) return true;
    c.x += 2.0f * r.sx;
    float right = (c.z <= sm_fetch(r.s, c.x, c.y)) ? 1.0f : 0.0f;
    if (fabsf(right - discType) == 0.0f) return true;
    c.x -= r.sx;
  }
  if (dy == 0.0f) {
    c.y += r.sy;
    float bottom = (c.z <= sm_fetch(r.s, c.x, c.y)) ? 1.0f : 0.0f;
    if (fabsf(bottom - discType) == 0.0f) return true;
    c.y -= 2.0f * r.sy;
    float top = (c.z <= sm_fetch(r.s, c.x, c.y)) ? 1.0f : 0.0f;
    if (fabsf(top - discType) == 0.0f) return true;
  }
  return false;
}

__device__ __forceinline__ bool nc_side(Rb& r, float4& c, float ux, float uy, float discB) {
  float dfl = sm_fetch(r.s, ux, uy);
  if (discB == 1.0f) {
    if (fabsf(c.z - dfl) < r.thr) { c.z -= r.thr; r.newDepth = c.z; }
  }
  float side = (c.z <= dfl) ? 1.0f : 0.0f;
  return fabsf(side - discB) == 0.0f;
}

// NonConservativeSMSR.frag:96-178
__device__ bool nc_getdisc_v(Rb& r, float4 c, float dx, float dy, float dr, float dg, float db) {
  float relx = c.x, rely = c.y;
  r.newDepth = c.z;
  if (dx == 0.0f) {
    if (dr == 0.5f || dr == 0.75f) { relx = c.x - r.sx; if (nc_side(r, c, relx, rely, db)) return true; }
    if (dr == 0.75f || dr == 0.25f) { relx = c.x + r.sx; if (nc_side(r, c, relx, rely, db)) return true; }
  }
  if (dy == 0.0f) {
    if (dg == 0.5f || dg == 0.75f) { rely = c.y + r.sy; if (nc_side(r, c, relx, rely, db)) return true; }
    if (dg == 0.75f || dg == 0.25f) { rely = c.y - r.sy; if (nc_side(r, c, relx, rely, db)) return true; }
  }
  return false;
}

// NonConservativeSMSR.frag:180-232
__device__ float nc_disc_length(Rb& r, float dr, float dg, float db, float4 c, float dx, float dy, float subCoord) {
  float foundEdgeEnd = 0.0f, dist = 0.0f;
  float stx = dx * r.sx, sty = dy * r.sy;
  c.x += stx; c.y += sty;
  for (int it = 0; it < r.max_search; it++) {
    float dfl = sm_fetch(r.s, c.x, c.y);
    if (db == 0.0f)
      if (fabsf(c.z - dfl) < r.thr) c.z -= r.thr;
    float center = (c.z <= dfl) ? 1.0f : 0.0f;
    if (fabsf(center - db) == 0.0f) {
      foundEdgeEnd = nc_getdisc_f(r, c, 0.0f, 0.0f, db) ? 1.0f : 0.0f;
      break;
    } else {
      if (!nc_getdisc_v(r, c, dx, dy, dr, dg, db)) break;
    }
    dist += 1.0f;
    c.x += stx; c.y += sty;
    if (db == 1.0f) c.z = r.newDepth;
  }
  return g_mix(-(dist + (1.0f - subCoord)), dist + (1.0f - subCoord), foundEdgeEnd);
}

// :234-243 (FilteredRBSM.frag:241)
__device__ __forceinline__ float nc_rel_pos(const Rb& r, float ax, float ay, float shadow) {
  float T = 1.0f;
  if (ax < 0.0f && ay < 0.0f) T = 0.0f;
  if (ax > 0.0f && ay > 0.0f) T = -2.0f;
  float edgeLength = g_min(fabsf(ax) + fabsf(ay), (float)r.max_search);
  float lead = r.filtered ? T : g_max(T, 2.0f * shadow - 1.0f);
  return (lead * fabsf(g_max(T * ax, T * ay))) / edgeLength;
}

// :266-274 / FilteredRBSM.frag:266-272
__device__ __forceinline__ float nc_revectorize(const Rb& r, float rx, float ry, float shadow) {
  if (r.filtered) {
    if (rx * ry < 0.0f) return (1.0f - shadow) + (2.0f * shadow - 1.0f) * g_max(rx, ry);
    else if (rx * ry == 0.0f) return shadow;
    else {
      float v = (1.0f - shadow) + (2.0f * shadow - 1.0f) * (rx + ry);
      return g_min(g_max(v, 0.0f), 1.0f);
    }
  }
  if ((rx * ry == 2.0f * shadow) ||
      ((fabsf(rx) * fabsf(ry) > 0.0f) && ((1.0f - shadow) + (2.0f * shadow - 1.0f) * (fabsf(rx) + fabsf(ry)) < 0.5f)))
    return 0.0f;
  return 1.0f;
}

// :276-286
__device__ __forceinline__ void nc_compute_disc(Rb& r, float4 c, float dfl, float& dr, float& dg, float& db) {
  float center = (c.z <= dfl) ? 1.0f : 0.0f;
  float dir[4];
  getdisc4(r, c, dir);
  float d0 = fabsf(dir[0] - center), d1 = fabsf(dir[1] - center), d2 = fabsf(dir[2] - center), d3 = fabsf(dir[3] - center);
  dr = (2.0f * d0 + d1) / 4.0f;
  dg = (2.0f * d2 + d3) / 4.0f;
  db = 1.0f - center;
}

__device__ void nc_rel(Rb& r, float4 c, float dr, float dg, float db, float subx, float suby, float shadow, float& rx, float& ry) {
  float left = nc_disc_length(r, dr, dg, db, c, -1.0f, 0.0f, (1.0f - subx));
  float right = nc_disc_length(r, dr, dg, db, c, 1.0f, 0.0f, subx);
  float down = nc_disc_length(r, dr, dg, db, c, 0.0f, -1.0f, (1.0f - suby));
  float up = nc_disc_length(r, dr, dg, db, c, 0.0f, 1.0f, suby);
  rx = nc_rel_pos(r, left, right, shadow);
  ry = nc_rel_pos(r, down, up, shadow);
}

// :288-304
__device__ float nc_smsr(Rb& r, float4 c) {
  float dfl = sm_fetch(r.s, c.x, c.y);
  float dr, dg, db;
  nc_compute_disc(r, c, dfl, dr, dg, db);
  float subx = g_fract(c.x * r.SWf), suby = g_fract(c.y * r.SHf);
  float shadow = (c.z <= dfl) ? 1.0f : 0.0f;
  if (dr > 0.0f || dg > 0.0f) {
    if (dr == 0.75f && dg == 0.75f) return g_mix(1.0f - shadow, 1.0f, r.si);
    float rx, ry;
    nc_rel(r, c, dr, dg, db, subx, suby, shadow, rx, ry);
    return g_mix(nc_revectorize(r, rx, ry, shadow), 1.0f, r.si);
  }
  return g_mix(shadow, 1.0f, r.si);
}

// :306-349
__device__ float nc_rpcf(Rb& r, const VisArgs& a, float4 c) {
  float incrWidth = 1.0f / (float)a.SW, incrHeight = 1.0f / (float)a.SH;
  float illum = 0.0f;
  int n = a.rpcf_n;
  if (n <= 0) return 1.0f;
  for (int iw = 0; iw < n; iw++)
    for (int ih = 0; ih < n; ih++) {
      float4 sc = make_float4(c.x + a.rpcf_off[iw] * incrWidth, c.y + a.rpcf_off[ih] * incrHeight, c.z, c.w);
      float dfl = sm_fetch(r.s, sc.x, sc.y);
      float shadow = (sc.z <= dfl) ? 1.0f : 0.0f;
      float dr, dg, db;
      nc_compute_disc(r, sc, dfl, dr, dg, db);
      if (dr > 0.0f || dg > 0.0f) {
        float subx = g_fract(sc.x * r.SWf), suby = g_fract(sc.y * r.SHf);
        float rx, ry;
        nc_rel(r, sc, dr, dg, db, subx, suby, shadow, rx, ry);
        illum += g_mix(nc_revectorize(r, rx, ry, shadow), 1.0f, r.si);
      } else {
        illum += g_mix(shadow, 1.0f, r.si);
      }
    }
  return illum / (float)(n * n);
}

// ---- conservative: ConservativeSMSR.frag ----
__device__ __forceinline__ void cs_disc(Rb& r, float4 c, float d[4]) {
  float dir[4];
  getdisc4(r, c, dir);
#pragma unroll
  for (int k = 0; k < 4; k++) d[k] = fabsf(dir[k] - 1.0f);
}

// :50-86
__device__ float cs_rel_distance(Rb& r, float4 t, float dx, float dy, float cc) {
  float foundSilhouetteEnd = 0.0f, distance = 0.0f;
  float stx = dx * r.sx, sty = dy * r.sy;
  t.x += stx; t.y += sty;
  for (int it = 0; it < r.max_search; it++) {
    float dfl = sm_fetch(r.s, t.x, t.y);
    if (fabsf(t.z - dfl) < r.thr) t.z -= r.thr;
    float center = (t.z <= dfl) ? 1.0f : 0.0f;
    if (!(center != 0.0f)) { foundSilhouetteEnd = 1.0f; break; }
    else {
      float d[4];
      cs_disc(r, t, d);
      if ((d[0] + d[1] + d[2] + d[3]) == 0.0f) break;
    }
    distance += 1.0f;
    t.x += stx; t.y += sty;
  }
  distance = distance + (1.0f - cc);
  return g_mix(-distance, distance, foundSilhouetteEnd);
}

// :99-108
__device__ __forceinline__ float cs_norm(const Rb& r, float ax, float ay) {
  float T = 1.0f;
  if (ax < 0.0f && ay < 0.0f) T = 0.0f;
  if (ax > 0.0f && ay > 0.0f) T = -2.0f;
  float length = g_min(fabsf(ax) + fabsf(ay), (float)r.max_search);
  return fabsf(g_max(T * ax, T * ay)) / length;
}

// :88-97,110-126
__device__ float cs_revec(Rb& r, float4 sc, float cx, float cy) {
  float dl = cs_rel_distance(r, sc, -1.0f, 0.0f, (1.0f - cx));
  float dr = cs_rel_distance(r, sc, 1.0f, 0.0f, cx);
  float db = cs_rel_distance(r, sc, 0.0f, -1.0f, (1.0f - cy));
  float dt = cs_rel_distance(r, sc, 0.0f, 1.0f, cy);
  float rx = cs_norm(r, dl, dr), ry = cs_norm(r, db, dt);
  if ((rx * ry > 0.0f) && (1.0f - rx > ry)) return r.si;
  return 1.0f;
}

// :128-144
__device__ float cs_smsr(Rb& r, float4 c) {
  float dfl = sm_fetch(r.s, c.x, c.y);
  float shadow = (c.z <= dfl) ? 1.0f : 0.0f;
  if (shadow == 0.0f) return r.si;
  float d[4];
  cs_disc(r, c, d);
  if ((d[0] + d[1] + d[2] + d[3]) == 0.0f) return 1.0f;
  else if ((d[0] + d[1]) == 2.0f || (d[2] + d[3]) == 2.0f) return r.si;
  float cx = g_fract(c.x * r.SWf), cy = g_fract(c.y * r.SHf);
  return cs_revec(r, c, cx, cy);
}

// :146-198
__device__ float cs_rpcf(Rb& r, const VisArgs& a, float4 c) {
  float incrWidth = 1.0f / (float)a.SW, incrHeight = 1.0f / (float)a.SH;
  float illum = 0.0f;
  int n = a.rpcf_n;
  if (n <= 0) return 1.0f;
  for (int iw = 0; iw < n; iw++)
    for (int ih = 0; ih < n; ih++) {
      float u = c.x + a.rpcf_off[iw] * incrWidth, v = c.y + a.rpcf_off[ih] * incrHeight;
      float dfl = sm_fetch(r.s, u, v);
      float shadow = (c.z <= dfl) ? 1.0f : r.si;
      if (shadow == 1.0f) {
        float4 sc = make_float4(u, v, c.z, c.w);
        float d[4];
        cs_disc(r, sc, d);
        if (d[0] == 0.0f && d[1] == 0.0f) illum += 1.0f;
        else {
          float subx = g_fract(sc.x * r.SWf), suby = g_fract(sc.y * r.SHf);
          illum += cs_revec(r, sc, subx, suby);
        }
      } else illum += r.si;
    }
  return illum / (float)(n * n);
}

#include "sgi_rbssm.cuh"
#include "sgi_edt.cuh"
}  // namespace
#define SGI_MOMENTS_KERNELS
#include "sgi_moments.cuh"
namespace {

// ================================ kernels ========================================================
// VA / VB: compile-time tap counts of the specialised variants (PCF: VA = taps per axis; PCSS: VA = blocker taps,
// VB = filter taps; 0 = generic run-time loops)
// (the specialised PCF / PCSS variants wait on their taps: capped at 40 registers = 6 CTAs per SM, where they fit without spills;
//  left to itself the compiler took 48 = 5 CTAs)
template <int TECH, int VA, int VB>
__global__ void __launch_bounds__(256, (VA > 0 ? 6 : 1)) k_visibility(const VisArgs a) {
  int x = a.rx0 + blockIdx.x * 32 + threadIdx.x, y = a.ry0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= a.rx1 || y >= a.ry1) return;
  size_t o = (size_t)y * a.W + x;
  float4 vertex = __ldg(&a.pos4[o]);
  if (vertex.x == 0.0f) { a.vis[o] = 0.0f; return; }              // discard (Shadow.frag:244): the target keeps its clear value 0
  float4 normal = __ldg(&a.nrm4[o]);
  float4 sc = mat4_mul(a.lmvp, vertex);
  float4 c;
  sgi_div3(sc.x, sc.y, sc.z, sc.w, c.x, c.y, c.z);
  c.w = sc.w / sc.w;
  float shadow = pre_evaluation(a, vertex, normal);
  Smap s = {a.sm, a.SW, a.SH, a.fw, a.fh};
  if (TECH == SGI_TECH_HARD || TECH == SGI_TECH_PCF || TECH == SGI_TECH_PCSS || TECH == SGI_TECH_RBSSM || TECH == SGI_TECH_PCF_TRICUBIC) {
    if (sc.w > 0.0f && shadow == 1.0f) {
      if (TECH == SGI_TECH_HARD) shadow = (c.z <= sm_fetch(s, c.x, c.y)) ? 1.0f : a.p.shadow_intensity;
      else if (TECH == SGI_TECH_PCF_TRICUBIC) shadow = pcf_tricubic(a, s, c);
      else if (TECH == SGI_TECH_RBSSM) {                             // RBSSM.frag:1372-1373
        Rb r = {s, a.sx, a.sy, a.p.depth_threshold, a.p.max_search, a.p.shadow_intensity, 0, a.fw, a.fh, 0.0f};
        shadow = ss_rbssm(a, r, c);
      }
      else if (TECH == SGI_TECH_PCF) { const TapSrc<false> g = {s.d, s.w, 0, 0}; shadow = pcf_t<VA, false>(a, s, g, c); }
      else shadow = pcss_t<VA, VB>(a, s, c);
    }
  } else if (shadow == 1.0f) {
    Rb r = {s, a.sx, a.sy, a.p.depth_threshold, a.p.max_search, a.p.shadow_intensity, TECH == SGI_TECH_RSMSS, a.fw, a.fh, 0.0f};
    if (TECH == SGI_TECH_RBSM_NONCONS) shadow = nc_smsr(r, c);
    else if (TECH == SGI_TECH_RBSM_CONS) shadow = cs_smsr(r, c);
    else if (TECH == SGI_TECH_RPCF_NONCONS || TECH == SGI_TECH_RSMSS) shadow = nc_rpcf(r, a, c);
    else if (TECH == SGI_TECH_RPCF_CONS) shadow = cs_rpcf(r, a, c);
  }
  a.vis[o] = shadow;
}

// PCF / PCSS with the shadow-map window of the CTA's 32x8 pixels staged in shared memory (128-bit loads, footprint found
// by warp-shuffle min/max reductions).  Identical arithmetic; only where the taps are read from changes.  A CTA whose
// footprint does not fit the 48 KB window (depth discontinuities, grazing angles) reads through L1/L2 as k_visibility.
template <int TECH, int VA, int VB>
__global__ void __launch_bounds__(256) k_visibility_staged(const VisArgs a) {
  extern __shared__ __align__(16) float stage[];
  __shared__ int red[40];
  const int x = a.rx0 + blockIdx.x * 32 + threadIdx.x, y = a.ry0 + blockIdx.y * 8 + threadIdx.y;
  const bool inb = x < a.rx1 && y < a.ry1;
  const size_t o = inb ? (size_t)y * a.W + x : 0;
  float4 vertex = make_float4(0.f, 0.f, 0.f, 1.f), normal = vertex, sc = vertex, c = vertex;
  float shadow = 0.0f;
  bool need = false;
  if (inb) {
    vertex = __ldg(&a.pos4[o]);
    if (vertex.x != 0.0f) {
      normal = __ldg(&a.nrm4[o]);
      sc = mat4_mul(a.lmvp, vertex);
      sgi_div3(sc.x, sc.y, sc.z, sc.w, c.x, c.y, c.z);
      c.w = sc.w / sc.w;
      shadow = pre_evaluation(a, vertex, normal);
      need = sc.w > 0.0f && shadow == 1.0f;
    }
  }
  const Smap s = {a.sm, a.SW, a.SH, a.fw, a.fh};
  const TapSrc<false> g = {s.d, s.w, 0, 0};
  TapSrc<true> sh;
  if (TECH == SGI_TECH_PCF) {
    const int n = VA ? VA : a.pcf_n;
    bool bad = false;
    const int lx = clamp_texel(c.x + a.pcf_du[0], s.fw, bad), hx = clamp_texel(c.x + a.pcf_du[max(n - 1, 0)], s.fw, bad);
    const int ly = clamp_texel(c.y + a.pcf_dv[0], s.fh, bad), hy = clamp_texel(c.y + a.pcf_dv[max(n - 1, 0)], s.fh, bad);
    const bool staged = stage_window(s, stage, red, need, bad, lx, hx, ly, hy, sh);
    if (need) shadow = staged ? pcf_t<VA, true>(a, s, sh, c) : pcf_t<VA, false>(a, s, g, c);
  } else {
    const int nb = VA ? VA : a.bs_n;
    bool bad = false;
    const int lx = clamp_texel(c.x + a.bs_q[0], s.fw, bad), hx = clamp_texel(c.x + a.bs_q[max(nb - 1, 0)], s.fw, bad);
    const int ly = clamp_texel(c.y + a.bs_q[0], s.fh, bad), hy = clamp_texel(c.y + a.bs_q[max(nb - 1, 0)], s.fh, bad);
    const bool staged = stage_window(s, stage, red, need, bad, lx, hx, ly, hy, sh);
    float pw = 0.0f;
    bool need2 = false;
    if (need) {
      const float avg = staged ? pcss_blockers<VA, true>(a, s, sh, c) : pcss_blockers<VA, false>(a, s, g, c);
      pw = pcss_penumbra(a.p, avg, c.z);
      const float stepSize = 2.0f * pw / (float)a.p.kernel_size;
      need2 = !(stepSize <= 0.0f || stepSize >= 1.0f);
      if (!need2) shadow = 1.0f;
    }
    if (__syncthreads_or(need2)) {                     // the filter runs only where a penumbra exists
      const float fw2 = ((float)a.p.kernel_size - 1.0f) * 0.5f;
      const int w0 = (int)(-fw2), w1 = (fw2 >= 0.0f) ? (int)fw2 : w0;
      bool bad2 = false;
      const int flx = clamp_texel(c.x + ((float)w0 * pw) / fw2, s.fw, bad2), fhx = clamp_texel(c.x + ((float)w1 * pw) / fw2, s.fw, bad2);
      const int fly = clamp_texel(c.y + ((float)w0 * pw) / fw2, s.fh, bad2), fhy = clamp_texel(c.y + ((float)w1 * pw) / fw2, s.fh, bad2);
      const bool staged2 = stage_window(s, stage, red, need2, bad2, flx, fhx, fly, fhy, sh);
      if (need2) shadow = staged2 ? pcss_filter<VB, true>(a, s, sh, c, pw) : pcss_filter<VB, false>(a, s, g, c, pw);
    }
  }
  if (inb) a.vis[o] = (vertex.x == 0.0f) ? 0.0f : shadow;
}

// AccurateSoftShadow.frag:52-133, monteCarlo branch
__global__ void __launch_bounds__(256) k_visibility_multi(const VisArgs a) {
  int x = a.rx0 + blockIdx.x * 32 + threadIdx.x, y = a.ry0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= a.rx1 || y >= a.ry1) return;
  size_t o = (size_t)y * a.W + x;
  float4 vertex = __ldg(&a.pos4[o]);
  if (vertex.x == 0.0f) { a.vis[o] = 0.0f; return; }
  const float* m = a.lmvp;
  float cx = m[0] * vertex.x + m[4] * vertex.y + m[8] * vertex.z;
  float cy = m[1] * vertex.x + m[5] * vertex.y + m[9] * vertex.z;
  float cz = m[2] * vertex.x + m[6] * vertex.y + m[10] * vertex.z;
  float cw = m[3] * vertex.x + m[7] * vertex.y + m[11] * vertex.z;
  float accShadow = 0.0f, count = 0.0f;
  const float accFactor = 1.0f;
  for (int l = 0; l < a.N; l++) {
    float4 t = __ldg(&a.trans[l]);
    float sx = cx + t.x, sy = cy + t.y, sz = cz + t.z, sw = cw + t.w;
    sgi_div3(sx, sy, sz, sw, sx, sy, sz);
    Smap s = {a.sm + a.layer * l, a.SW, a.SH, a.fw, a.fh};
    float dfl = sm_fetch(s, sx, sy);
    accShadow += ((sz <= dfl) ? 1.0f : a.p.shadow_intensity) * accFactor;
    count += accFactor;
  }
  a.vis[o] = a.p.multi_partial ? accShadow : accShadow / count;
}

// The same accumulation with the vertex map resolved on the fly: GBuffer.vert/frag's world position of the pixel's winning
// primitive, interpolated exactly as the tile rasteriser's resolve does (perspective-correct, same expression order), so the
// positions - and with them every tap - are identical to the materialised-G-buffer path while 16 B/pixel less is written and read.
// world position of the pixel's winning primitive; false: the fragment is discarded (background, or vertex.x == 0)
__device__ __forceinline__ bool fused_position(const VisArgs& a, int x, int y, size_t o, float& vx, float& vy, float& vz) {
  const unsigned int prim = __ldg(&a.ids[o]);
  if (prim == 0xFFFFFFFFu) return false;                         // background: vertex map (0,0,0,1), discarded (x == 0)
  const int t = (int)(prim >> 3), sub = (int)(prim & 7u);
  const int slot = (sub == 0) ? t : __ldg(&a.ovf_base[t]) + sub - 1;
  if (slot < 0) return false;                                    // (an id that does not belong to this camera pass: never dereferenced)
  SgiRec r; SgiRecAttr at;
  {
    const uint4* rq = reinterpret_cast<const uint4*>(&a.rec[slot]);
    const uint4* aq = reinterpret_cast<const uint4*>(&a.attr[slot]);
    uint4* rd = reinterpret_cast<uint4*>(&r); uint4* ad = reinterpret_cast<uint4*>(&at);
    rd[0] = __ldg(rq); rd[1] = __ldg(rq + 1); rd[2] = __ldg(rq + 2);
    ad[0] = __ldg(aq); ad[1] = __ldg(aq + 1); ad[2] = __ldg(aq + 2); ad[3] = __ldg(aq + 3); ad[4] = __ldg(aq + 4); ad[5] = __ldg(aq + 5);
  }
  const int PX = x * SGI_SUBPIX + SGI_SUBPIX / 2, PY = y * SGI_SUBPIX + SGI_SUBPIX / 2;
  const long long E0 = (long long)(r.X2 - r.X1) * (long long)(PY - r.Y1) - (long long)(r.Y2 - r.Y1) * (long long)(PX - r.X1);
  const long long E1 = (long long)(r.X0 - r.X2) * (long long)(PY - r.Y2) - (long long)(r.Y0 - r.Y2) * (long long)(PX - r.X2);
  const long long E2 = (long long)(r.X1 - r.X0) * (long long)(PY - r.Y0) - (long long)(r.Y1 - r.Y0) * (long long)(PX - r.X0);
  const float q0 = ((float)E0 * r.ia) * at.iw[0];
  const float q1 = ((float)E1 * r.ia) * at.iw[1];
  const float q2 = ((float)E2 * r.ia) * at.iw[2];
  const float iq = 1.0f / ((q0 + q1) + q2);
  vx = ((q0 * at.A[0][0] + q1 * at.A[1][0]) + q2 * at.A[2][0]) * iq;
  vy = ((q0 * at.A[0][1] + q1 * at.A[1][1]) + q2 * at.A[2][1]) * iq;
  vz = ((q0 * at.A[0][2] + q1 * at.A[1][2]) + q2 * at.A[2][2]) * iq;
  return vx != 0.0f;
}
// byte of plane p of pixel (x, y) in the mask layout
__device__ __forceinline__ size_t mask_index(const VisArgs& a, int x, int y, int p) {
  const int rk = y / a.mask_rows;
  return ((size_t)rk * a.mask_planes + p) * a.mask_strip + (size_t)(y - rk * a.mask_rows) * a.W + x;
}

// MASK = false: AccurateSoftShadow.frag's accumulation (the sum, or with multi_partial = 1 the un-normalised sum).  MASK = true
// (multi_partial = 2): which of this rank's lights reach the pixel, as bits at the lights' indices in the whole set; the rank's own
// strip of the visibility target additionally receives the discard decision (0 = discarded, 1 = a fragment), which is all
// k_mask_resolve needs besides the masks.
// (8 CTAs per SM = 32 registers: the kernel waits on its taps, and at 34 registers - 6 CTAs - it ran 20 % longer)
template <bool MASK>
__global__ void __launch_bounds__(256, 8) k_visibility_multi_fused(const VisArgs a) {
  int x = a.rx0 + blockIdx.x * 32 + threadIdx.x, y = a.ry0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= a.rx1 || y >= a.ry1) return;
  size_t o = (size_t)y * a.W + x;
  float vx, vy, vz;
  const bool valid = fused_position(a, x, y, o, vx, vy, vz);
  if (MASK && y >= a.own_r0 && y < a.own_r1) a.vis[o] = valid ? 1.0f : 0.0f;
  if (!valid) {
    if (MASK) { for (int p = 0; p < a.mask_planes; p++) a.mask[mask_index(a, x, y, p)] = 0; }
    else a.vis[o] = 0.0f;
    return;
  }
  const float* m = a.lmvp;
  float cx = m[0] * vx + m[4] * vy + m[8] * vz;
  float cy = m[1] * vx + m[5] * vy + m[9] * vz;
  float cz = m[2] * vx + m[6] * vy + m[10] * vz;
  float cw = m[3] * vx + m[7] * vy + m[11] * vz;
  float accShadow = 0.0f, count = 0.0f;
  const float accFactor = 1.0f;
  unsigned int lit = 0u;
  for (int l = 0; l < a.N; l++) {
    float4 tr = __ldg(&a.trans[l]);
    float sx = cx + tr.x, sy = cy + tr.y, sz = cz + tr.z, sw = cw + tr.w;
    sgi_div3(sx, sy, sz, sw, sx, sy, sz);
    Smap s = {a.sm + a.layer * l, a.SW, a.SH, a.fw, a.fh};
    float dfl = sm_fetch(s, sx, sy);
    if (MASK) { if (sz <= dfl) lit |= 1u << __ldg(&a.gid[l]); }
    else {
      accShadow += ((sz <= dfl) ? 1.0f : a.p.shadow_intensity) * accFactor;
      count += accFactor;
    }
  }
  if (MASK) { for (int p = 0; p < a.mask_planes; p++) a.mask[mask_index(a, x, y, p)] = (unsigned char)((lit >> (8 * p)) & 255u); }
  else a.vis[o] = a.p.multi_partial ? accShadow : accShadow / count;
}

// The lit masks of all ranks, summed (disjoint bits: the sum is their union), turned into the visibility of this rank's strip: the
// accumulation loop of AccurateSoftShadow.frag:100-127 replayed over the whole light set in its own order, so the result has the
// bits of the un-sharded frame for every shadow intensity (partial float sums would depend on how the lights were dealt).
// Reads nothing of the camera pass: the discard decision was left in the strip by this rank's own accumulation kernel.
__global__ void __launch_bounds__(256) k_mask_resolve(const VisArgs a) {
  int x = a.rx0 + blockIdx.x * 32 + threadIdx.x, y = a.ry0 + blockIdx.y * 8 + threadIdx.y;
  if (x >= a.rx1 || y >= a.ry1) return;
  size_t o = (size_t)y * a.W + x;
  if (a.vis[o] == 0.0f) return;                                  // discarded fragment: the target keeps its 0
  unsigned int lit = 0u;
  for (int p = 0; p < a.mask_planes; p++) lit |= (unsigned int)a.mask[mask_index(a, x, y, p)] << (8 * p);
  float accShadow = 0.0f, count = 0.0f;
  const float accFactor = 1.0f;
  for (int l = 0; l < a.mask_total; l++) {
    accShadow += (((lit >> l) & 1u) ? 1.0f : a.p.shadow_intensity) * accFactor;
    count += accFactor;
  }
  a.vis[o] = accShadow / count;
}

// ShadowMapping/Shaders/GBuffer/PhongShading.frag:11-47 (shadeScene).  Note `vec3 E = normalize(-vertex)` normalises the
// vec4 (w included) before truncation, and the specular term is scaled by (shadow - shadowIntensity).
struct ShadeArgs {
  float mv[16], nm[9], lpos[3]; float si; float clear[4];
  const float4* pos4; const float4* nrm4; const float4* albedo4; const float* vis; float4* out; int W, H;
};
__global__ void __launch_bounds__(256) k_shade_phong(const ShadeArgs a) {
  int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= a.W || y >= a.H) return;
  size_t o = (size_t)y * a.W + x;
  float4 vertex = __ldg(&a.pos4[o]);
  if (vertex.x == 0.0f) { a.out[o] = make_float4(a.clear[0], a.clear[1], a.clear[2], a.clear[3]); return; }
  float4 normal = __ldg(&a.nrm4[o]);
  float shadow = __ldg(&a.vis[o]);
  float specShadow = shadow - a.si;
  float4 ev = mat4_mul(a.mv, vertex);
  float n0 = (a.nm[0] * normal.x + a.nm[3] * normal.y) + a.nm[6] * normal.z;
  float n1 = (a.nm[1] * normal.x + a.nm[4] * normal.y) + a.nm[7] * normal.z;
  float n2 = (a.nm[2] * normal.x + a.nm[5] * normal.y) + a.nm[8] * normal.z;
  { float inv = 1.0f / sqrtf((n0 * n0 + n1 * n1) + n2 * n2); n0 *= inv; n1 *= inv; n2 *= inv; }
  float L0 = a.lpos[0] - ev.x, L1 = a.lpos[1] - ev.y, L2 = a.lpos[2] - ev.z;
  { float inv = 1.0f / sqrtf((L0 * L0 + L1 * L1) + L2 * L2); L0 *= inv; L1 *= inv; L2 *= inv; }
  float mx = -ev.x, my = -ev.y, mz = -ev.z, mw = -ev.w;
  float einv = 1.0f / sqrtf(((mx * mx + my * my) + mz * mz) + mw * mw);
  float E0 = mx * einv, E1 = my * einv, E2 = mz * einv;
  float d = (n0 * L0 + n1 * L1) + n2 * L2;
  float r0 = -(L0 - 2.0f * d * n0), r1 = -(L1 - 2.0f * d * n1), r2 = -(L2 - 2.0f * d * n2);
  { float inv = 1.0f / sqrtf((r0 * r0 + r1 * r1) + r2 * r2); r0 *= inv; r1 *= inv; r2 *= inv; }
  float ndl = g_max(d, 0.0f);
  float rde = g_max((r0 * E0 + r1 * E1) + r2 * E2, 0.0f);
  float pw = powf(rde, 0.3f * 10.0f);
  float4 col = a.albedo4 ? __ldg(&a.albedo4[o]) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float cc[4] = {col.x, col.y, col.z, col.w};
  const float amb[4] = {0.4f, 0.4f, 0.4f, 1.0f}, spec[4] = {0.25f, 0.25f, 0.25f, 1.0f}, diff[4] = {0.5f, 0.5f, 0.5f, 1.0f};
  float res[4];
#pragma unroll
  for (int c = 0; c < 4; c++) res[c] = (shadow * cc[c]) * ((diff[c] * ndl + (specShadow * spec[c]) * pw) + amb[c]);
  a.out[o] = make_float4(res[0], res[1], res[2], res[3]);
}

}  // namespace

namespace {
__device__ __forceinline__ unsigned int hash_u32(unsigned int x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
// operands for the divide check: raw random bits (every exponent, denormals, infinities, NaNs), or - three draws in four - values
// with exponents near each other, where light-space coordinates live
__device__ __forceinline__ float test_operand(unsigned int h, unsigned int sel) {
  if ((sel & 3u) == 0u) return __uint_as_float(h);
  const unsigned int e = 100u + (hash_u32(h ^ 0x9e3779b9u) % 56u);             // 2^-27 .. 2^28
  return __uint_as_float((h & 0x807FFFFFu) | (e << 23));
}
__global__ void __launch_bounds__(256) k_divide_selftest(unsigned long long n, unsigned int seed, unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256ull) {
    const unsigned int k = hash_u32((unsigned int)i ^ seed) + (unsigned int)(i >> 32) * 0x632be5abu;
    const unsigned int sel = hash_u32(k + 4u);
    const float a0 = test_operand(hash_u32(k), sel), a1 = test_operand(hash_u32(k + 1u), sel >> 2), a2 = test_operand(hash_u32(k + 2u), sel >> 4);
    const float b = test_operand(hash_u32(k + 3u), sel >> 6);
    float q0, q1, q2;
    sgi_div3(a0, a1, a2, b, q0, q1, q2);
    const float p0 = a0 / b, p1 = a1 / b, p2 = a2 / b;
    const bool same0 = __float_as_uint(q0) == __float_as_uint(p0) || (q0 != q0 && p0 != p0);
    const bool same1 = __float_as_uint(q1) == __float_as_uint(p1) || (q1 != q1 && p1 != p1);
    const bool same2 = __float_as_uint(q2) == __float_as_uint(p2) || (q2 != q2 && p2 != p2);
    bad += (same0 ? 0 : 1) + (same1 ? 0 : 1) + (same2 ? 0 : 1);
  }
  if (bad) atomicAdd(mismatches, bad);
}
}  // namespace

// sgi_divide_selftest: the shared-reciprocal divide against the plain division on n random operand quadruples
int sgi_divide_selftest_run(sgi_ctx* ctx, unsigned long long n, unsigned int seed, unsigned long long* mismatches) {
  unsigned long long* d = nullptr;
  SGI_CUDA(ctx, cudaMalloc((void**)&d, 8));
  SGI_CUDA(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
  k_divide_selftest<<<ctx->n_sm * 8, 256, 0, ctx->stream>>>(n, seed, d);
  ctx->launches++;
  cudaError_t e = cudaMemcpyAsync(mismatches, d, 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d);
  SGI_CUDA(ctx, e);
  return SGI_OK;
}

// geometry of the lit-mask target (SGI_BUF_LIGHT_MASK): ceil(total / 8) byte planes per rank strip
static void fill_mask_args(const sgi_ctx* ctx, VisArgs& a) {
  a.mask = (unsigned char*)ctx->buf[SGI_BUF_LIGHT_MASK];
  a.mask_total = ctx->mask_total; a.mask_planes = (ctx->mask_total + 7) / 8;
  a.mask_rows = sgi_strip_rows(ctx); a.mask_strip = (size_t)a.mask_rows * ctx->W;
  a.gid = ctx->d_light_gid;
  const int rk = ctx->comm_n > 1 ? ctx->comm_rank : 0;
  a.own_r0 = rk * a.mask_rows < ctx->H ? rk * a.mask_rows : ctx->H;
  a.own_r1 = (rk + 1) * a.mask_rows < ctx->H ? (rk + 1) * a.mask_rows : ctx->H;
}

// sgi_reduce_lights in mask mode: visibility of rows [r0, r1) from the summed masks
int sgi_mask_resolve_run(sgi_ctx* ctx, int r0, int r1, cudaStream_t st) {
  if (r1 <= r0) return SGI_OK;
  VisArgs a;
  memset(&a, 0, sizeof(a));
  a.p = ctx->params;
  a.W = ctx->W; a.H = ctx->H; a.rx0 = 0; a.ry0 = r0; a.rx1 = ctx->W; a.ry1 = r1;
  a.vis = (float*)ctx->buf[SGI_BUF_VISIBILITY];
  fill_mask_args(ctx, a);
  dim3 block(32, 8), grid((ctx->W + 31) / 32, (r1 - r0 + 7) / 8);
  k_mask_resolve<<<grid, block, 0, st>>>(a);
  ctx->launches++;
  SGI_CUDA(ctx, cudaGetLastError());
  return SGI_OK;
}

int sgi_shade_run(sgi_ctx* ctx, const float clear_rgba[4]) {
  ShadeArgs a;
  for (int k = 0; k < 16; k++) a.mv[k] = ctx->cam_mv[k];
  for (int k = 0; k < 9; k++) a.nm[k] = ctx->cam_nm[k];
  for (int k = 0; k < 3; k++) a.lpos[k] = ctx->light_pos[k];
  for (int k = 0; k < 4; k++) a.clear[k] = clear_rgba[k];
  a.si = ctx->params.shadow_intensity;
  a.pos4 = (const float4*)ctx->buf[SGI_BUF_GBUF_POS]; a.nrm4 = (const float4*)ctx->buf[SGI_BUF_GBUF_NRM];
  const bool tex_on = ctx->has_uv && ctx->uv_V == ctx->V && (ctx->d_tex[0] || ctx->d_tex[1] || ctx->d_tex[2]);
  a.albedo4 = (ctx->has_rgb || tex_on) ? (const float4*)ctx->buf[SGI_BUF_GBUF_ALBEDO] : nullptr;
  a.vis = (const float*)ctx->buf[SGI_BUF_VISIBILITY]; a.out = (float4*)ctx->buf[SGI_BUF_SHADED];
  a.W = ctx->W; a.H = ctx->H;
  dim3 block(32, 8), grid((ctx->W + 31) / 32, (ctx->H + 7) / 8);
  k_shade_phong<<<grid, block, 0, ctx->stream>>>(a);
  ctx->launches++;
  SGI_CUDA(ctx, cudaGetLastError());
  return SGI_OK;
}

int sgi_host_pcf_offsets(int kernel_order, int penumbra_size, int inclusive, float* out, int cap) {
  // Shadow.frag:93-98 / NonConservativeSMSR.frag:313-319, evaluated in fp32 exactly as the shader's float loop
  volatile float offset = (float)penumbra_size;
  volatile float stepSize = 2 * offset / (float)kernel_order;
  int n = 0;
  if (!(stepSize > 0.0f)) return 0;
  for (volatile float w = -offset; inclusive ? (w <= offset) : (w < offset); w = w + stepSize) {
    if (n >= cap) return -1;
    out[n++] = w;
  }
  return n;
}

void sgi_moments_quantization(float m[16], float minv[16], float t[4]) {
  MomQuant Q;
  mom_quantization(Q);
  for (int k = 0; k < 16; k++) { m[k] = Q.m[k]; minv[k] = Q.minv[k]; }
  for (int k = 0; k < 4; k++) t[k] = Q.t[k];
}

// filterShadowMap(), ShadowMapping/src/main.cpp:374-398: X pass (moment target -> FILTER_X) then Y pass (FILTER_X -> FILTER_Y),
// both into window-sized targets with the window's texel step
int sgi_moments_filter_run(sgi_ctx* ctx, cudaStream_t st) {
  MomFilterArgs f;
  f.W = ctx->W; f.H = ctx->H;
  f.order = ctx->params.kernel_order;
  for (int k = 0; k < SGI_MOM_MAX_ORDER; k++) f.kernel[k] = 0.0f;
  mom_gaussian_kernel(f.order, f.kernel);
  { volatile float ss = 1.0f / (float)ctx->W, tt = 1.0f / (float)ctx->H; f.step_s = ss; f.step_t = tt; }      // GaussianFilter.frag:27-28
  const bool logs = ctx->params.technique == SGI_TECH_ESM;
  dim3 block(32, 8), grid((ctx->W + 31) / 32, (ctx->H + 7) / 8);
  f.src = (const float4*)ctx->buf[SGI_BUF_MOMENTS]; f.sw = ctx->SW; f.sh = ctx->SH; f.dst = (float4*)ctx->buf[SGI_BUF_MOMENTS_X];
  const bool o7 = f.order == 7;
  if (logs) { if (o7) k_mom_filter<true, true, 7><<<grid, block, 0, st>>>(f); else k_mom_filter<true, true, 0><<<grid, block, 0, st>>>(f); }
  else { if (o7) k_mom_filter<true, false, 7><<<grid, block, 0, st>>>(f); else k_mom_filter<true, false, 0><<<grid, block, 0, st>>>(f); }
  f.src = (const float4*)ctx->buf[SGI_BUF_MOMENTS_X]; f.sw = ctx->W; f.sh = ctx->H; f.dst = (float4*)ctx->buf[SGI_BUF_MOMENTS_FILTERED];
  if (logs) { if (o7) k_mom_filter<false, true, 7><<<grid, block, 0, st>>>(f); else k_mom_filter<false, true, 0><<<grid, block, 0, st>>>(f); }
  else { if (o7) k_mom_filter<false, false, 7><<<grid, block, 0, st>>>(f); else k_mom_filter<false, false, 0><<<grid, block, 0, st>>>(f); }
  ctx->launches += 2;
  SGI_CUDA(ctx, cudaGetLastError());
  return SGI_OK;
}

static int moments_visibility_run(sgi_ctx* ctx, cudaStream_t st) {
  MomVisArgs a;
  for (int k = 0; k < 16; k++) { a.mv[k] = ctx->cam_mv[k]; a.lmvp[k] = ctx->h_light_mvp_b[k]; }
  for (int k = 0; k < 9; k++) a.nm[k] = ctx->cam_nm[k];
  for (int k = 0; k < 3; k++) a.lpos[k] = ctx->light_pos[k];
  a.shadow_intensity = ctx->params.shadow_intensity; a.z_near = ctx->params.z_near; a.z_far = ctx->params.z_far;
  a.pos4 = (const float4*)ctx->buf[SGI_BUF_GBUF_POS]; a.nrm4 = (const float4*)ctx->buf[SGI_BUF_GBUF_NRM];
  a.vis = (float*)ctx->buf[SGI_BUF_VISIBILITY];
  a.W = ctx->W; a.H = ctx->H;
  a.rx0 = ctx->params.rect_x0; a.ry0 = ctx->params.rect_y0; a.rx1 = ctx->params.rect_x1; a.ry1 = ctx->params.rect_y1;
  if (a.rx1 <= a.rx0 || a.ry1 <= a.ry0) { a.rx0 = 0; a.ry0 = 0; a.rx1 = ctx->W; a.ry1 = ctx->H; }
  a.rx0 = a.rx0 < 0 ? 0 : a.rx0; a.ry0 = a.ry0 < 0 ? 0 : a.ry0;
  a.rx1 = a.rx1 > ctx->W ? ctx->W : a.rx1; a.ry1 = a.ry1 > ctx->H ? ctx->H : a.ry1;
  a.fmap = (const float4*)ctx->buf[SGI_BUF_MOMENTS_FILTERED]; a.mw = ctx->filtered_w; a.mh = ctx->filtered_h;
  float m[16];
  sgi_moments_quantization(m, a.minv, a.qt);
  const int rw = a.rx1 - a.rx0, rh = a.ry1 - a.ry0;
  if (rw <= 0 || rh <= 0) return SGI_OK;
  dim3 block(32, 8), grid((rw + 31) / 32, (rh + 7) / 8);
  int tslot = sgi_timing_begin(ctx, SGI_PASS_VIS_KERNEL, st);
  switch (ctx->params.technique) {
    case SGI_TECH_VSM: k_mom_visibility<SGI_TECH_VSM><<<grid, block, 0, st>>>(a); break;
    case SGI_TECH_ESM: k_mom_visibility<SGI_TECH_ESM><<<grid, block, 0, st>>>(a); break;
    case SGI_TECH_EVSM: k_mom_visibility<SGI_TECH_EVSM><<<grid, block, 0, st>>>(a); break;
    default: k_mom_visibility<SGI_TECH_MSM><<<grid, block, 0, st>>>(a); break;
  }
  ctx->launches++;
  sgi_timing_end(ctx, SGI_PASS_VIS_KERNEL, tslot, st);
  SGI_CUDA(ctx, cudaGetLastError());
  return SGI_OK;
}

// texels the current technique's fixed tap window reaches from its centre texel (PCF grid, PCSS blocker search), + margin; 0 = the
// technique has no window the min-max cull applies to
int sgi_minmax_reach(const sgi_ctx* ctx) {
  const sgi_params& p = ctx->params;
  const float smax = (float)(ctx->SW > ctx->SH ? ctx->SW : ctx->SH);
  if (p.technique == SGI_TECH_PCF) {
    float m = 0.0f;
    for (int k = 0; k < ctx->pcf_n; k++) m = fmaxf(m, fabsf(ctx->pcf_off[k]));
    const float r = m / (float)(ctx->SW < ctx->SH ? ctx->SW : ctx->SH) * smax + 3.0f;
    return r < 4096.0f ? (int)r + 1 : 0;
  }
  if (p.technique == SGI_TECH_PCSS) {
    const float bsw = ((float)ctx->SW <= 1024.0f) ? (float)p.light_source_radius / (float)ctx->SW : (float)p.light_source_radius / 1024.0f;
    const float r = fabsf(bsw) * smax + 3.0f;             // |w| <= filterWidth: |(w * bsw) / filterWidth| <= bsw
    return r < 4096.0f ? (int)r + 1 : 0;
  }
  return 0;
}

int sgi_shadow_run(sgi_ctx* ctx, cudaStream_t stream) {
  if (sgi_is_moment_tech(ctx->params.technique)) return moments_visibility_run(ctx, stream);
  VisArgs a;
  a.p = ctx->params;
  for (int k = 0; k < 16; k++) a.mv[k] = ctx->cam_mv[k];
  for (int k = 0; k < 9; k++) a.nm[k] = ctx->cam_nm[k];
  for (int k = 0; k < 3; k++) a.lpos[k] = ctx->light_pos[k];
  const bool multi = ctx->params.technique == SGI_TECH_MULTI_HARD;
  const float* lm = ctx->h_light_mvp_b + (multi ? (size_t)(ctx->N - 1) * 16 : 0);
  if (multi && ctx->has_multi_common) lm = ctx->multi_common;
  for (int k = 0; k < 16; k++) a.lmvp[k] = lm[k];
  a.pos4 = (const float4*)ctx->buf[SGI_BUF_GBUF_POS];
  a.nrm4 = (const float4*)ctx->buf[SGI_BUF_GBUF_NRM];
  a.vis = (float*)ctx->buf[SGI_BUF_VISIBILITY];
  a.W = ctx->W; a.H = ctx->H;
  a.rx0 = ctx->params.rect_x0; a.ry0 = ctx->params.rect_y0; a.rx1 = ctx->params.rect_x1; a.ry1 = ctx->params.rect_y1;
  if (a.rx1 <= a.rx0 || a.ry1 <= a.ry0) { a.rx0 = 0; a.ry0 = 0; a.rx1 = ctx->W; a.ry1 = ctx->H; }
  a.rx0 = a.rx0 < 0 ? 0 : a.rx0; a.ry0 = a.ry0 < 0 ? 0 : a.ry0;
  a.rx1 = a.rx1 > ctx->W ? ctx->W : a.rx1; a.ry1 = a.ry1 > ctx->H ? ctx->H : a.ry1;
  a.sm = (const float*)ctx->buf[SGI_BUF_SHADOW_MAP];
  a.SW = ctx->SW; a.SH = ctx->SH; a.fw = (float)ctx->SW; a.fh = (float)ctx->SH;
  a.sx = (float)(1.0 / ctx->SW); a.sy = (float)(1.0 / ctx->SH);      // MyGLGeometryViewer.cpp:238
  a.pcf_n = ctx->pcf_n; a.rpcf_n = ctx->rpcf_n;
  for (int k = 0; k < SGI_MAX_PCF_TAPS; k++) { a.pcf_off[k] = ctx->pcf_off[k]; a.rpcf_off[k] = ctx->rpcf_off[k]; }
  a.trans = (const float4*)ctx->d_light_trans; a.N = ctx->N; a.layer = (size_t)ctx->SW * ctx->SH;
  a.ids = nullptr; a.rec = nullptr; a.attr = nullptr; a.ovf_base = nullptr;
  a.mask = nullptr; a.mask_planes = 0; a.mask_rows = 1; a.mask_strip = 0; a.mask_total = 0; a.gid = nullptr; a.own_r0 = a.own_r1 = 0;
  a.dmin = nullptr; a.dmax = nullptr; a.mm_w = 0; a.mm_limit = 0.0f; a.pcf_reach = 1.0e30f; a.bs_reach = 1.0e30f;
  a.pcss_early_out = (ctx->pcss_early_out && ctx->params.light_source_radius >= 0 && ctx->params.z_near >= 0 && ctx->params.kernel_size > 0 &&
                      ctx->params.blocker_search_size <= SGI_MAX_PCF_TAPS && !ctx->vis_staged) ? 1 : 0;
  {
    volatile float incrWidth = 1.0f / (float)ctx->SW, incrHeight = 1.0f / (float)ctx->SH;      // Shadow.frag:89-90
    for (int k = 0; k < SGI_MAX_PCF_TAPS; k++) {
      volatile float du = ctx->pcf_off[k] * incrWidth, dv = ctx->pcf_off[k] * incrHeight;
      a.pcf_du[k] = du; a.pcf_dv[k] = dv;
    }
    // PlausibleSoftShadow.frag:172-180
    volatile float bsw = ((float)ctx->SW <= 1024.0f) ? (float)ctx->params.light_source_radius / (float)ctx->SW
                                                     : (float)ctx->params.light_source_radius / 1024.0f;
    volatile float filterWidth = ((float)ctx->params.blocker_search_size - 1.0f) * 0.5f;
    a.bs_w0 = (int)(-filterWidth);
    a.bs_n = 0;
    for (int w = a.bs_w0; (float)w <= filterWidth && a.bs_n < SGI_MAX_PCF_TAPS; w++) {
      volatile float num = (float)w * bsw;
      volatile float q = num / filterWidth;
      a.bs_q[a.bs_n++] = q;
    }
  }

  if (ctx->mm_valid && ctx->vis_minmax_cull && !ctx->vis_staged && (ctx->params.technique == SGI_TECH_PCF || ctx->params.technique == SGI_TECH_PCSS)) {
    const int n = ctx->mm_w * ctx->mm_h;
    a.dmin = (const float*)(ctx->d_mm + (size_t)(2 + 2 * ctx->mm_set) * n); a.dmax = (const float*)(ctx->d_mm + (size_t)(3 + 2 * ctx->mm_set) * n);
    a.mm_w = ctx->mm_w; a.mm_limit = (float)(32 * ctx->mm_radius - 1);
    const float smax = (float)(ctx->SW > ctx->SH ? ctx->SW : ctx->SH);
    float pr = 0.0f, br = 0.0f;
    for (int k = 0; k < a.pcf_n; k++) { pr = fmaxf(pr, fabsf(a.pcf_du[k]) * smax); pr = fmaxf(pr, fabsf(a.pcf_dv[k]) * smax); }
    for (int k = 0; k < a.bs_n; k++) br = fmaxf(br, fabsf(a.bs_q[k]) * smax);
    a.pcf_reach = pr + 2.0f; a.bs_reach = br + 2.0f;
  }
  int rw = a.rx1 - a.rx0, rh = a.ry1 - a.ry0;
  if (rw <= 0 || rh <= 0) return SGI_OK;
  cudaStream_t st = stream;
  // EDT shadow mapping: the hard shadows are the SMSR branch of the RBSM programs (main.cpp:406-409,
  // NonConservativeSMSR.frag:381) rendered into a side buffer, the EDT filter chain then produces the visibility
  const bool edtsm = ctx->params.technique == SGI_TECH_EDTSM_NONCONS || ctx->params.technique == SGI_TECH_EDTSM_CONS;
  if (edtsm) {
    if (rw != ctx->W || rh != ctx->H) { ctx->err = "EDT shadow mapping needs the whole screen (no rect)"; return SGI_ERR_INVALID; }
    const size_t px = (size_t)ctx->W * ctx->H;
    const int nbands = (ctx->H + 31) / 32, nblk = (ctx->W + 31) / 32;
    const size_t need[SGI_EDT_NBUF] = {px * 4, px * 8, px, px * 2, px * 8, px * 8, 16, (size_t)nbands * ctx->W * 4, (size_t)ctx->H * nblk * 4};   // hard, aux, site, col, a2, b2, flag, band ends, block minima
    for (int k = 0; k < SGI_EDT_NBUF; k++)
      if (ctx->edt_bytes[k] != need[k]) {
        SGI_CUDA(ctx, cudaStreamSynchronize(st));
        if (ctx->edt_buf[k]) cudaFree(ctx->edt_buf[k]);
        ctx->edt_buf[k] = nullptr; ctx->edt_bytes[k] = 0;
        SGI_CUDA(ctx, cudaMalloc(&ctx->edt_buf[k], need[k]));
        ctx->edt_bytes[k] = need[k];
      }
    a.vis = (float*)ctx->edt_buf[0];
    a.p.technique = ctx->params.technique == SGI_TECH_EDTSM_CONS ? SGI_TECH_RBSM_CONS : SGI_TECH_RBSM_NONCONS;
  }
  if (multi && ctx->trans_dirty) {
    // lightMVPTrans[i] = column 3 of bias*lightMVP_i (SoftShadowMapping/src/Viewers/MyGLGeometryViewer.cpp:187-190)
    std::vector<float> tmp((size_t)ctx->N * 4);
    for (int i = 0; i < ctx->N; i++) for (int k = 0; k < 4; k++) tmp[4 * (size_t)i + k] = ctx->h_light_mvp_b[16 * (size_t)i + 12 + k];
    SGI_CUDA(ctx, cudaMemcpyAsync(ctx->d_light_trans, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice, st));
    SGI_CUDA(ctx, cudaStreamSynchronize(st));      // tmp is pageable
    ctx->trans_dirty = false;
  }
  // the reference clears the target to 0 before the full-screen pass (main.cpp:403-405) and discarded pixels keep it:
  // here the kernel writes that 0 itself (every pixel of the rectangle is visited), so there is no separate clear pass
  dim3 block(32, 8), grid((rw + 31) / 32, (rh + 7) / 8);
  int tslot = sgi_timing_begin(ctx, SGI_PASS_VIS_KERNEL, st);
  const sgi_params& P = a.p;
  // shared-memory staging of the shadow-map window (PCF / PCSS) is an option (sgi_set_option "vis_staged"): measured on
  // B200 it only pays when the map is much larger than L1 can cover (4096^2 PCSS: 1.01 -> 0.89 ms) and loses at the
  // c2 sizes (0.077 -> 0.097 ms), because the kernel is issue-bound, not L1-bound (profiles/r1_vis_staging.txt)
  const size_t stage_bytes = (size_t)SGI_STAGE_FLOATS * 4;
  if (ctx->vis_staged && !(ctx->func_cfg & (1ull << 40))) {        // per-device attribute: once per context
    ctx->func_cfg |= 1ull << 40;
    cudaFuncSetAttribute(k_visibility_staged<SGI_TECH_PCF, 7, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
    cudaFuncSetAttribute(k_visibility_staged<SGI_TECH_PCF, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
    cudaFuncSetAttribute(k_visibility_staged<SGI_TECH_PCSS, 7, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
    cudaFuncSetAttribute(k_visibility_staged<SGI_TECH_PCSS, 7, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
    cudaFuncSetAttribute(k_visibility_staged<SGI_TECH_PCSS, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
  }
  const bool staged = ctx->vis_staged != 0;
  // tap counts the float/int loops of the shaders produce for the current parameters
  const int nb_taps = a.bs_n;
  const int nk_taps = 2 * (int)(((float)P.kernel_size - 1.0f) * 0.5f) + 1;
  switch (P.technique) {
    case SGI_TECH_HARD: k_visibility<SGI_TECH_HARD, 0, 0><<<grid, block, 0, st>>>(a); break;
    case SGI_TECH_PCF_TRICUBIC: k_visibility<SGI_TECH_PCF_TRICUBIC, 0, 0><<<grid, block, 0, st>>>(a); break;
    case SGI_TECH_PCF:
      if (staged) {
        if (a.pcf_n == 7) k_visibility_staged<SGI_TECH_PCF, 7, 0><<<grid, block, stage_bytes, st>>>(a);
        else k_visibility_staged<SGI_TECH_PCF, 0, 0><<<grid, block, stage_bytes, st>>>(a);
      } else {
        if (a.pcf_n == 7) k_visibility<SGI_TECH_PCF, 7, 0><<<grid, block, 0, st>>>(a);          // kernelOrder 7 (reference default)
        else k_visibility<SGI_TECH_PCF, 0, 0><<<grid, block, 0, st>>>(a);
      }
      break;
    case SGI_TECH_PCSS:
      if (staged) {
        if (nb_taps == 7 && nk_taps == 15) k_visibility_staged<SGI_TECH_PCSS, 7, 15><<<grid, block, stage_bytes, st>>>(a);
        else if (nb_taps == 7 && nk_taps == 7) k_visibility_staged<SGI_TECH_PCSS, 7, 7><<<grid, block, stage_bytes, st>>>(a);
        else k_visibility_staged<SGI_TECH_PCSS, 0, 0><<<grid, block, stage_bytes, st>>>(a);
      } else {
        if (nb_taps == 7 && nk_taps == 15) k_visibility<SGI_TECH_PCSS, 7, 15><<<grid, block, 0, st>>>(a);   // reference default
        else if (nb_taps == 7 && nk_taps == 7) k_visibility<SGI_TECH_PCSS, 7, 7><<<grid, block, 0, st>>>(a); // after "reset"
        else k_visibility<SGI_TECH_PCSS, 0, 0><<<grid, block, 0, st>>>(a);
      }
      break;
    case SGI_TECH_RBSM_NONCONS: k_visibility<SGI_TECH_RBSM_NONCONS, 0, 0><<<grid, block, 0, st>>>(a); break;
    case SGI_TECH_RBSM_CONS: k_visibility<SGI_TECH_RBSM_CONS, 0, 0><<<grid, block, 0, st>>>(a); break;
    case SGI_TECH_RPCF_NONCONS: k_visibility<SGI_TECH_RPCF_NONCONS, 0, 0><<<grid, block, 0, st>>>(a); break;
    case SGI_TECH_RPCF_CONS: k_visibility<SGI_TECH_RPCF_CONS, 0, 0><<<grid, block, 0, st>>>(a); break;
    case SGI_TECH_RSMSS: k_visibility<SGI_TECH_RSMSS, 0, 0><<<grid, block, 0, st>>>(a); break;
    case SGI_TECH_MULTI_HARD:
      if (ctx->params.multi_fused) {
        const SgiScratch& sc = ctx->scratch[1];          // the camera pass's records (sgi_render_prim_ids)
        a.ids = (const unsigned int*)ctx->buf[SGI_BUF_PRIM_ID]; a.rec = sc.d_rec; a.attr = sc.d_attr; a.ovf_base = sc.d_ovf_base;
        if (ctx->params.multi_partial == 2) fill_mask_args(ctx, a);
        if (ctx->params.multi_partial == 2) k_visibility_multi_fused<true><<<grid, block, 0, st>>>(a);
        else k_visibility_multi_fused<false><<<grid, block, 0, st>>>(a);
      } else k_visibility_multi<<<grid, block, 0, st>>>(a);
      break;
    case SGI_TECH_RBSSM: {
      if (!ctx->rbssm_compact) { k_visibility<SGI_TECH_RBSSM, 0, 0><<<grid, block, 0, st>>>(a); break; }
      // work list of the penumbra pixels + one warp per listed pixel (sgi_rbssm.cuh)
      const size_t need = (size_t)rw * rh * sizeof(RbssmItem) + 64;
      if (ctx->rbssm_bytes < need) {
        SGI_CUDA(ctx, cudaStreamSynchronize(st));
        if (ctx->rbssm_buf) cudaFree(ctx->rbssm_buf);
        ctx->rbssm_buf = nullptr; ctx->rbssm_bytes = 0;
        SGI_CUDA(ctx, cudaMalloc(&ctx->rbssm_buf, need));
        ctx->rbssm_bytes = need;
      }
      int* counters = (int*)ctx->rbssm_buf;                    // [0] items, [1] cursor
      RbssmItem* items = (RbssmItem*)((char*)ctx->rbssm_buf + 64);
      SGI_CUDA(ctx, cudaMemsetAsync(counters, 0, 8, st));
      k_rbssm_prepare<<<grid, block, 0, st>>>(a, items, counters);
      k_rbssm_taps<<<ctx->n_sm * 4, 256, 0, st>>>(a, items, counters, counters + 1);
      ctx->launches++;
      break;
    }
    default: ctx->err = "unknown technique"; return SGI_ERR_INVALID;
  }
  ctx->launches++;
  if (edtsm) {
    // filterHardShadowsUsingEDT (main.cpp:416-447)
    EdtArgs e;
    e.pos4 = a.pos4; e.nrm4 = a.nrm4; e.vis_in = (const float*)ctx->edt_buf[0]; e.vis_out = (float*)ctx->buf[SGI_BUF_VISIBILITY];
    e.aux = (float2*)ctx->edt_buf[1]; e.site = (unsigned char*)ctx->edt_buf[2]; e.col = (short*)ctx->edt_buf[3];
    e.a2 = (float2*)ctx->edt_buf[4]; e.b2 = (float2*)ctx->edt_buf[5]; e.any_site = (int*)ctx->edt_buf[6];
    e.nearest = (short2*)ctx->buf[SGI_BUF_EDT_NEAREST];
    e.W = ctx->W; e.H = ctx->H;
    for (int k = 0; k < 16; k++) { e.cmvp[k] = ctx->cam_mvp[k]; e.mv[k] = ctx->cam_mv[k]; }
    { volatile double pen = (double)ctx->params.penumbra_size / 5.0; e.penumbra = (float)pen; }      // main.cpp:421: int / 5.0, narrowed
    e.si = ctx->params.shadow_intensity; e.order = ctx->params.kernel_order; e.z_near = ctx->params.z_near; e.z_far = ctx->params.z_far;
    { volatile float t = tanf(45.0f / 2.0f); volatile float d = 2.0f * t; e.dscreen = 1.0f / d; }    // MeanFilter.frag:43, fov = 45 as is (MyGLGeometryViewer.cpp:6)
    SGI_CUDA(ctx, cudaMemsetAsync(e.any_site, 0, 4, st));
    k_edt_prepare<<<grid, block, 0, st>>>(a, e);
    k_edt_sites<<<grid, block, 0, st>>>(e);
    const int nbands = (ctx->H + SGI_EDT_BAND - 1) / SGI_EDT_BAND, nblk = (ctx->W + 31) / 32;
    short2* ends = (short2*)ctx->edt_buf[7]; int* bmin = (int*)ctx->edt_buf[8];
    k_edt_band_ends<<<dim3((ctx->W + 127) / 128, nbands), 128, 0, st>>>(e, ends, nbands);
    k_edt_cols<<<dim3((ctx->W + 127) / 128, nbands), 128, 0, st>>>(e, ends, nbands);
    k_edt_blockmin<<<dim3((nblk + 7) / 8, ctx->H), dim3(32, 8), 0, st>>>(e, bmin, nblk);
    k_edt_rows<<<grid, block, 0, st>>>(e, bmin, nblk);
    k_edt_normalize<<<grid, block, 0, st>>>(e);
    k_mean_filter<false, false><<<grid, block, 0, st>>>(e, e.a2, e.b2, 1);
    k_mean_filter<true, true><<<grid, block, 0, st>>>(e, e.b2, nullptr, 0);
    ctx->launches += 9;
  }
  sgi_timing_end(ctx, SGI_PASS_VIS_KERNEL, tslot, st);
  SGI_CUDA(ctx, cudaGetLastError());
  return SGI_OK;
}
