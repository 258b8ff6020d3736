// sgi_edt.cuh — Euclidean-distance-transform shadow mapping (EDTSM), device side + pass sequence.  Included by
// sgi_shadow.cu (uses its VisArgs, mat4_mul, pre_evaluation).
//
// Replaces, in the ShadowMapping program with shadowParams.EDTSM,
//   the extra channels of the hard-shadow target  Shaders/RBSM/NonConservativeSMSR.frag:384-393   -> k_edt_prepare
//   initializeInput                               include/EDT/pba2DKernel.h:513-535               -> k_edt_sites
//   pba2DVoronoiDiagram(16,16,16)                 src/EDT/pba2DHost.cu:200-229 (13 PBA kernels)   -> k_edt_cols, k_edt_rows
//   pbaNormalizeDistanceTransform                 include/EDT/pba2DKernel.h:551-569               -> k_edt_normalize
//   two MeanFilter.frag passes                    src/main.cpp:425-445, Shaders/Filter/MeanFilter.frag -> k_mean_filter
// (the only CUDA in the reference: legacy texture references + CUDA-GL interop; here plain global-memory kernels on the
// context's buffers).  The Voronoi diagram is exact (what the Parallel Banding Algorithm computes) but found differently:
//   phase 1  nearest site row of the same column for every pixel: columns cut into 32-row bands, band ends exchanged
//            through a small table, two register sweeps per band (k_edt_band_ends, k_edt_cols)
//   phase 2  one thread per pixel: best of the columns' candidates, blocks of 32 columns visited outwards from the pixel;
//            a per-row table of each block's smallest vertical distance (k_edt_blockmin) lets whole blocks be skipped, and
//            the walk ends when the horizontal gap alone exceeds the best distance found (k_edt_rows)
// Ties between equidistant sites go to the smallest (y, x) (PBA's own pick depends on its band schedule).  fp32 / fp64
// sub-expressions follow the CUDA and GLSL sources literally; bit-identical to oracle/oracle_edt_impl.h.
#pragma once

#define SGI_EDT_MARKER (-32768)                                   // pba2D.h:61

struct EdtArgs {
  const float4* pos4; const float4* nrm4; const float* vis_in;   // G-buffer, hard shadows (RBSM) of this frame
  float* vis_out;
  float2* aux;                // (camera window depth, pre-evaluated shadow); background (0, 0)
  unsigned char* site; int* any_site;
  short* col; short2* nearest;
  float2* a2; float2* b2;
  int W, H;
  float cmvp[16], mv[16];
  float penumbra, si;
  int order, z_near, z_far;
  float dscreen;
};

// pba2DKernel.h:503-510 — `2.0 * n` and the quotient are double
__device__ __forceinline__ float edt_linearize_cuda(float depth) {
  const float n = 1.0f, f = 1000.0f;
  const float den = f + n - depth * (f - n);
  return (float)((2.0 * (double)n) / (double)den);
}
// MeanFilter.frag:15-22 — fp32
__device__ __forceinline__ float edt_linearize_glsl(float depth, int z_near, int z_far) {
  const float n = (float)z_near, f = (float)z_far;
  return (2.0f * n) / (f + n - depth * (f - n));
}

// NonConservativeSMSR.frag:384-393: depth = (MVP*vertex).z/w*0.5+0.5, preEvaluatedShadow
__global__ void __launch_bounds__(256) k_edt_prepare(const VisArgs va, const EdtArgs e) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= e.W || y >= e.H) return;
  const size_t o = (size_t)y * e.W + x;
  const float4 vertex = __ldg(&e.pos4[o]);
  if (vertex.x == 0.0f) { e.aux[o] = make_float2(0.0f, 0.0f); return; }
  const float4 normal = __ldg(&e.nrm4[o]);
  const float4 position = mat4_mul(e.cmvp, vertex);
  float depth = position.z / position.w;
  depth = depth * 0.5f + 0.5f;
  e.aux[o] = make_float2(depth, pre_evaluation(va, vertex, normal));
}

// initializeInput: a pixel is a site when one of its 8 neighbours has another shadow value, a close linearised depth and
// a light-facing pre-evaluation
__global__ void __launch_bounds__(256) k_edt_sites(const EdtArgs e) {
  const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
  if (px >= e.W || py >= e.H) return;
  const size_t o = (size_t)py * e.W + px;
  const float cx = __ldg(&e.vis_in[o]);
  const float cl = edt_linearize_cuda(__ldg(&e.aux[o]).x);
  bool is_site = false;
  for (int x = -1; x <= 1; x++)
    for (int y = -1; y <= 1; y++)
      if (px + x >= 0 && px + x < e.W && py + y >= 0 && py + y < e.H) {
        const size_t q = (size_t)(py + y) * e.W + (px + x);
        const float2 qa = __ldg(&e.aux[q]);
        if (__ldg(&e.vis_in[q]) != cx && (double)fabsf(cl - edt_linearize_cuda(qa.x)) <= 0.0025 && qa.y == 1.0f) is_site = true;
      }
  e.site[o] = is_site ? 1 : 0;
  if (is_site && *e.any_site == 0) *e.any_site = 1;
}

// ---- phase 1: nearest site row of the same column (tie: the smaller row); MARKER if the column holds no site -----------
// The column is cut into bands of SGI_EDT_BAND rows so that W x bands threads work instead of W:
//   k_edt_band_ends   per (column, band): lowest and highest site row inside the band
//   k_edt_cols        per (column, band): carry the nearest site below / above the band in from the other bands' ends
//                     (a walk over <= H/32 band records), then two register sweeps over the band's rows
#define SGI_EDT_BAND 32
__global__ void __launch_bounds__(128) k_edt_band_ends(const EdtArgs e, short2* __restrict__ ends, int nbands) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (x >= e.W) return;
  int lo = SGI_EDT_MARKER, hi = SGI_EDT_MARKER;
  const int y0 = b * SGI_EDT_BAND;
#pragma unroll 8
  for (int k = 0; k < SGI_EDT_BAND; k++) {
    const int y = y0 + k;
    if (y < e.H && e.site[(size_t)y * e.W + x]) { if (lo == SGI_EDT_MARKER) lo = y; hi = y; }
  }
  ends[(size_t)b * e.W + x] = make_short2((short)lo, (short)hi);
}

__global__ void __launch_bounds__(128) k_edt_cols(const EdtArgs e, const short2* __restrict__ ends, int nbands) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (x >= e.W) return;
  int below = SGI_EDT_MARKER, above = SGI_EDT_MARKER;           // nearest sites strictly outside this band
  for (int k = b - 1; k >= 0 && below == SGI_EDT_MARKER; k--) below = ends[(size_t)k * e.W + x].y;
  for (int k = b + 1; k < nbands && above == SGI_EDT_MARKER; k++) above = ends[(size_t)k * e.W + x].x;
  const int y0 = b * SGI_EDT_BAND;
  unsigned char sv[SGI_EDT_BAND];
  short dn[SGI_EDT_BAND];
#pragma unroll
  for (int k = 0; k < SGI_EDT_BAND; k++) sv[k] = (y0 + k < e.H) ? e.site[(size_t)(y0 + k) * e.W + x] : 0;
  int last = below;
#pragma unroll
  for (int k = 0; k < SGI_EDT_BAND; k++) { if (sv[k]) last = y0 + k; dn[k] = (short)last; }       // nearest at or below
  last = above;
#pragma unroll
  for (int k = SGI_EDT_BAND - 1; k >= 0; k--) {                                                   // nearest above: strictly closer wins
    const int y = y0 + k;
    if (y >= e.H) continue;
    if (sv[k]) last = y;
    int best = dn[k];
    if (last != SGI_EDT_MARKER && (best == SGI_EDT_MARKER || last - y < y - best)) best = last;
    e.col[(size_t)y * e.W + x] = (short)best;
  }
}

// ---- phase 2: per pixel, the best of the columns' candidates ---------------------------------------------------------
// k_edt_blockmin: per row and block of 32 columns, the smallest vertical distance^2 of the block's candidates; lets the
// row scan skip whole blocks that cannot beat (or tie) the best distance found so far.
__global__ void __launch_bounds__(256) k_edt_blockmin(const EdtArgs e, int* __restrict__ bmin, int nblk) {
  const int lane = threadIdx.x, cb = blockIdx.x * 8 + threadIdx.y, y = blockIdx.y;
  if (cb >= nblk) return;
  const int c = cb * 32 + lane;
  int v = 0x7fffffff;
  if (c < e.W) {
    const int sy = e.col[(size_t)y * e.W + c];
    if (sy != SGI_EDT_MARKER) v = (y - sy) * (y - sy);
  }
  v = __reduce_min_sync(0xffffffffu, v);
  if (lane == 0) bmin[(size_t)y * nblk + cb] = v;
}

__global__ void __launch_bounds__(256) k_edt_rows(const EdtArgs e, const int* __restrict__ bmin, int nblk) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= e.W || y >= e.H) return;
  const size_t row = (size_t)y * e.W;
  int bx = SGI_EDT_MARKER, by = SGI_EDT_MARKER;
  if (*e.any_site) {
    int best = 0x7fffffff;
    const int cb0 = x >> 5;
    const int* __restrict__ bm = bmin + (size_t)y * nblk;
    // blocks in order of their horizontal gap to x: own block, then left/right pairs.  A block is scanned only if
    // gap^2 + (its smallest vertical distance^2) can beat or tie the best; the walk ends when the gap alone cannot.
    for (int db = 0; db < nblk; db++) {
      bool any_reachable = false;
      for (int s = 0; s < (db ? 2 : 1); s++) {
        const int cb = s ? cb0 + db : cb0 - db;
        if (cb < 0 || cb >= nblk) continue;
        const int c_lo = cb << 5, c_hi = min(c_lo + 31, e.W - 1);
        const int gap = (x < c_lo) ? c_lo - x : ((x > c_hi) ? x - c_hi : 0);
        if ((long long)gap * gap > (long long)best) continue;
        any_reachable = true;
        const int m = __ldg(&bm[cb]);
        if (m == 0x7fffffff || (long long)gap * gap + m > (long long)best) continue;
        for (int c = c_lo; c <= c_hi; c++) {
          const int sy = __ldg(&e.col[row + c]);
          if (sy == SGI_EDT_MARKER) continue;
          const int d = c - x;
          const int dd = d * d + (y - sy) * (y - sy);
          if (dd < best || (dd == best && (sy < by || (sy == by && c < bx)))) { best = dd; bx = c; by = sy; }
        }
      }
      if (!any_reachable && db > 0) break;
    }
  }
  e.nearest[row + x] = make_short2((short)bx, (short)by);
}

// pbaNormalizeDistanceTransform: world-space distance to the nearest site -> penumbra ramp.  A MARKER site reads texel
// (0,0) (tex2D clamps).
__global__ void __launch_bounds__(256) k_edt_normalize(const EdtArgs e) {
  const int px = blockIdx.x * 32 + threadIdx.x, py = blockIdx.y * 8 + threadIdx.y;
  if (px >= e.W || py >= e.H) return;
  const size_t o = (size_t)py * e.W + px;
  const short2 ns = e.nearest[o];
  const int sx = min(max((int)ns.x, 0), e.W - 1), sy = min(max((int)ns.y, 0), e.H - 1);
  const size_t so = (size_t)sy * e.W + sx;
  const float ix = __ldg(&e.vis_in[o]);
  const float2 ia = __ldg(&e.aux[o]), sa = __ldg(&e.aux[so]);
  const float4 p1 = __ldg(&e.pos4[o]), p2 = __ldg(&e.pos4[so]);
  const float dx = p1.x - p2.x, dy = p1.y - p2.y, dz = p1.z - p2.z;
  const float distance = sqrtf((dx * dx + dy * dy) + dz * dz);
  float r;
  if (sa.y != 1.0f || (double)fabsf(edt_linearize_cuda(sa.x) - edt_linearize_cuda(ia.x)) > 0.0005 || distance > e.penumbra / 2) r = ix;
  else {
    const float q = distance / e.penumbra;
    const float v0 = (ix == e.si) ? (float)(0.5 - (double)q) : (float)(0.5 + (double)q);
    r = (1 - e.si) * v0 + e.si * 1.0f;
  }
  e.a2[o] = make_float2(r, ia.x);
}

// one texel of an (r,g) image, CLAMP_TO_BORDER(0): GL_NEAREST, or GL_LINEAR of level 0 (weights from fract(u*size-0.5),
// texels accumulated in the order 00,10,01,11)
template <bool LINEAR>
__device__ __forceinline__ float2 edt_fetch2(const float2* __restrict__ img, int W, int H, float u, float v) {
  if (!LINEAR) {
    const float fi = floorf(u * (float)W), fj = floorf(v * (float)H);
    if (!(fi >= 0.0f && fi < (float)W && fj >= 0.0f && fj < (float)H)) return make_float2(0.0f, 0.0f);
    return __ldg(&img[(size_t)(int)fj * W + (int)fi]);
  }
  const float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
  const float x0 = floorf(x), y0 = floorf(y);
  const float ax = x - x0, ay = y - y0;
  float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float fx = x0 + (float)(k & 1), fy = y0 + (float)(k >> 1);
    const float wgt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
    float2 t = make_float2(0.0f, 0.0f);
    if (fx >= 0.0f && fx < (float)W && fy >= 0.0f && fy < (float)H) t = __ldg(&img[(size_t)(int)fy * W + (int)fx]);
    acc.x += wgt * t.x; acc.y += wgt * t.y;
  }
  return acc;
}

// MeanFilter.frag:35-77, one separable pass; FINAL writes the red channel into the visibility buffer
template <bool LINEAR, bool FINAL>
__global__ void __launch_bounds__(256) k_mean_filter(const EdtArgs e, const float2* __restrict__ in2, float2* __restrict__ out2, int horizontal) {
  const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
  if (i >= e.W || j >= e.H) return;
  const size_t o = (size_t)j * e.W + i;
  const float4 vertex = __ldg(&e.pos4[o]);
  if (vertex.x == 0.0f) {                                   // discard: the target keeps its clear value
    if (FINAL) e.vis_out[o] = 0.0f; else out2[o] = make_float2(0.0f, 0.0f);
    return;
  }
  const float cs = (((float)i + 0.5f) / (float)e.W * 2.0f - 1.0f) * 0.5f + 0.5f;      // f_texcoord (MeanFilter.vert)
  const float ct = (((float)j + 0.5f) / (float)e.H * 2.0f - 1.0f) * 0.5f + 0.5f;
  const float steps = 1.0f / (float)e.W, stept = 1.0f / (float)e.H;
  const float dirs = horizontal ? 1.0f : 0.0f, dirt = horizontal ? 0.0f : 1.0f;
  const float2 color = edt_fetch2<LINEAR>(in2, e.W, e.H, cs, ct);
  int count = 0;
  float sum = 0.0f;
  const float deye = -(mat4_mul(e.mv, vertex)).z;
  float kernelCenter = (e.dscreen * (float)e.order * 50.0f) / (deye * 2.0f);
  if (kernelCenter > 4096.0f) kernelCenter = 4096.0f;     // guard (eye distance ~ 0): the shader would loop without end
  const float lc = edt_linearize_glsl(color.y, e.z_near, e.z_far);
  for (float sample = -kernelCenter; sample <= kernelCenter; sample++) {
    const float2 cur = edt_fetch2<LINEAR>(in2, e.W, e.H, cs + dirs * sample * steps, ct + dirt * sample * stept);
    float r = cur.x;
    if (r == 0.0f) r = color.x;                                                              // adjustColor :24-33
    else if (fabsf(edt_linearize_glsl(cur.y, e.z_near, e.z_far) - lc) >= 0.0005f) r = color.x;
    sum += r;
    count++;
  }
  sum /= (float)count;
  if (FINAL) e.vis_out[o] = sum; else out2[o] = make_float2(sum, color.y);
}
