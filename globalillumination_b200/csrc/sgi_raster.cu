// sgi_raster.cu — tile-binned triangle rasteriser for sm_100a (K1 light-view depth, K2 camera G-buffer,
// K4 shadow-volume counting).  Replaces the fixed-function GL raster behind
//   renderShadowMap()   ShadowMapping/src/main.cpp:350-361   (Scene.vert:11-20, glPolygonOffset :246)
//   renderGBuffer()     ShadowMapping/src/main.cpp:363-372   (GBuffer.vert:12-23, GBuffer.frag:32-38)
//   stencil pass        ShadowVolumes/src/main.cpp:154-172   (INCR_WRAP/DECR_WRAP on z-pass)
//
// Pipeline (all on one stream, no host round trip, no memset):
//   k_setup_bin  1 thread / source triangle: transform, clip (near/far + 16x guard band), snap to 1/256 px,
//                integer edge set-up, depth plane, polygon offset -> 64-byte SgiRec (+ attribute record); the same
//                thread then appends the record to the list of every 64x64 tile it really overlaps (exact edge /
//                tile-corner test; the larger records of a warp are flattened into (record, tile) pairs walked by all
//                32 lanes).  Lists are fixed-capacity segments (one atomic per pair): single pass, no count + scan + fill.
//   k_order      one CTA: list lengths -> work items of the tile kernel (hot tiles subdivided, busiest first); snapshots
//                and re-zeroes the binner's counters for the next pass
//   k_tile<MODE> 1 CTA / tile: the tile lives in shared memory (u32 depth, u64 depth|prim key or i32 count);
//                warps pull triangles off the tile's list, reject 8x4-pixel blocks of the bounding box 32 at
//                a time (one block per lane, conservative corner test) and rasterise the surviving blocks
//                with one lane per pixel (exact int64 edge functions, top-left rule); shared-memory atomics
//                resolve visibility; the tile is written to HBM exactly once (depth tiles: one bulk copy per row,
//                cp.async.bulk shared -> global; the clear is fused: there is no separate memset pass).
//
// HBM traffic per pass = geometry once + every output texel once; depth never bounces
// through global atomics.  Numerics follow DESIGN.md §3 to the bit (-fmad=false).
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "sgi_internal.cuh"
#include "sgi_moments.cuh"

// modes whose tile payload is the 64-bit (depth | primitive) key: the attribute passes that need to know the winning triangle
#define SGI_KEYED(M) ((M) == SGI_MODE_GBUFFER || (M) == SGI_MODE_GBUFFER_RGB || (M) == SGI_MODE_MOMENTS || (M) == SGI_MODE_IDS)

#define SGI_GRID_DEP_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")     // programmatic dependent launch: see launch_pdl

namespace {

struct CV { float x, y, z, w, b0, b1, b2; };

__device__ __forceinline__ float plane_dist(const CV& v, int p) {
  switch (p) {
    case 0: return v.w + v.z;
    case 1: return v.w - v.z;
    case 2: return SGI_GUARD * v.w + v.x;
    case 3: return SGI_GUARD * v.w - v.x;
    case 4: return SGI_GUARD * v.w + v.y;
    default: return SGI_GUARD * v.w - v.y;
  }
}

// a inside, b outside; always evaluated inside -> outside so shared edges clip identically
__device__ __forceinline__ CV clip_lerp(const CV& a, const CV& b, float da, float db) {
  float t = da / (da - db);
  CV r;
  r.x = a.x + t * (b.x - a.x);
  r.y = a.y + t * (b.y - a.y);
  r.z = a.z + t * (b.z - a.z);
  r.w = a.w + t * (b.w - a.w);
  r.b0 = a.b0 + t * (b.b0 - a.b0);
  r.b1 = a.b1 + t * (b.b1 - a.b1);
  r.b2 = a.b2 + t * (b.b2 - a.b2);
  return r;
}

__device__ int clip_polygon(CV* poly, int n, bool& clipped, int skip_far) {
  CV tmp[10];
  float d[10];
  clipped = false;
  for (int p = 0; p < 6; p++) {
    if (p == 1 && skip_far) continue;
    bool any_out = false;
    for (int i = 0; i < n; i++) { d[i] = plane_dist(poly[i], p); if (!(d[i] >= 0.0f)) any_out = true; }
    if (!any_out) continue;
    clipped = true;
    int m = 0;
    for (int i = 0; i < n; i++) {
      int j = (i + 1 == n) ? 0 : i + 1;
      bool in_i = d[i] >= 0.0f, in_j = d[j] >= 0.0f;
      if (in_i) {
        tmp[m++] = poly[i];
        if (!in_j) tmp[m++] = clip_lerp(poly[i], poly[j], d[i], d[j]);
      } else if (in_j) {
        tmp[m++] = clip_lerp(poly[j], poly[i], d[j], d[i]);
      }
    }
    n = m;
    if (n < 3) return 0;
    for (int i = 0; i < n; i++) poly[i] = tmp[i];
  }
  return n;
}

__device__ __forceinline__ CV xform(const float* __restrict__ m, float x, float y, float z) {
  CV o;
  o.x = ((m[0] * x + m[4] * y) + m[8] * z) + m[12];
  o.y = ((m[1] * x + m[5] * y) + m[9] * z) + m[13];
  o.z = ((m[2] * x + m[6] * y) + m[10] * z) + m[14];
  o.w = ((m[3] * x + m[7] * y) + m[11] * z) + m[15];
  return o;
}

struct SetupBinArgs {
  // set-up
  const float* xyz; const float* nrm; const float* rgb; const int32_t* idx; int T;
  float mvp[16];
  int W, H;
  int use_offset; float factor, units;
  int no_far_clip;                               // depth clamp (shadow volumes, depth-fail mode): the far plane does not clip, depths saturate at 1
  SgiRec* rec; SgiRecAttr* attr; int32_t* ovf_base; int32_t* counters;
  const float* uv; SgiRecUV* uvrec;              // texture coordinates (optional)
  // binning
  int tiles_x, tx0, ty0, tx1, ty1;               // tile grid pitch and the inclusive tile range of the job rectangle
  int32_t* tile_cnt; int32_t* pairs; int cap;    // per-tile append cursor; list of tile t = pairs[t * cap .. t * cap + cap)
  int2* spill; int spill_cap;                    // (tile, record) pairs that found their tile's list full (counters[5] = count)
  int32_t* big_list;                             // records spanning > big_tiles tiles: not binned by the set-up kernel (counters[3] = count); with k_bin_big
                                                 // the ones beyond huge_tiles are listed from the far end of the buffer instead (counters[6] = count)
  int few_tiles;                                 // the pass has at most 128 tiles: the CTA's larger records are walked tile-major (one atomic per warp and tile)
  int big_tiles, huge_tiles, big_cap;            // thresholds of the two classes (huge_tiles = INT_MAX without k_bin_big), entries the buffer holds
  const unsigned int* tile_zmax;                 // shadow volumes: largest scene depth of each tile (float bits), or null
};

__device__ __forceinline__ void invalidate(SgiRec* r) {
  SgiRec z;
  z.X0 = z.Y0 = z.X1 = z.Y1 = z.X2 = z.Y2 = 0;
  z.z0 = z.dz1 = z.dz2 = z.ia = z.zoff = 0.0f;
  z.prim_front = -1;
  z.px0 = z.py0 = 1; z.px1 = z.py1 = 0;
  z.pad0 = z.pad1 = 0;
  *r = z;
}

// Records of one (clipped) polygon of source triangle t: window transform, snap, fan triangulation, edge / depth-plane set-up,
// attribute record.  NP = capacity of the polygon arrays: instantiated with 3 for triangles that need no clipping (the common
// case: everything stays in registers) and 10 for clipped ones (local memory).  Identical arithmetic either way.
template <int NP>
__device__ __forceinline__ int emit_records(const SetupBinArgs& a, int t, int i0, int i1, int i2, const CV* poly, int n, bool was_clipped,
                                            SgiRec& first, int& base) {
  float hw = (float)a.W * 0.5f, hh = (float)a.H * 0.5f;
  int32_t X[NP], Y[NP];
  float Z[NP], IW[NP];
#pragma unroll
  for (int k = 0; k < NP; k++) {
    if (k >= n) break;
    float nx = poly[k].x / poly[k].w, ny = poly[k].y / poly[k].w, nz = poly[k].z / poly[k].w;
    float xw = nx * hw + hw, yw = ny * hh + hh;
    Z[k] = nz * 0.5f + 0.5f;
    IW[k] = 1.0f / poly[k].w;
    float sx = xw * (float)SGI_SUBPIX, sy = yw * (float)SGI_SUBPIX;
    if (!(fabsf(sx) < 1.0e9f) || !(fabsf(sy) < 1.0e9f) || !(fabsf(Z[k]) < 1.0e9f)) { invalidate(&a.rec[t]); return 0; }
    X[k] = __float2int_rn(sx);
    Y[k] = __float2int_rn(sy);
  }
  if (n > 3) {
    base = a.T + atomicAdd(&a.counters[0], n - 3);
    a.ovf_base[t] = base;
  }
#pragma unroll
  for (int f = 1; f + 1 < NP; f++) {
    if (f + 1 >= n) break;
    int slot = (f == 1) ? t : base + (f - 2);
    int id0 = 0, id1 = f, id2 = f + 1;
    long long area2 = (long long)(X[id1] - X[id0]) * (long long)(Y[id2] - Y[id0]) -
                      (long long)(X[id2] - X[id0]) * (long long)(Y[id1] - Y[id0]);
    if (area2 == 0) { invalidate(&a.rec[slot]); continue; }
    int front = area2 > 0;
    if (area2 < 0) { int s = id1; id1 = id2; id2 = s; area2 = -area2; }
    SgiRec r;
    r.X0 = X[id0]; r.Y0 = Y[id0]; r.X1 = X[id1]; r.Y1 = Y[id1]; r.X2 = X[id2]; r.Y2 = Y[id2];
    int mnx = min(r.X0, min(r.X1, r.X2)), mxx = max(r.X0, max(r.X1, r.X2));
    int mny = min(r.Y0, min(r.Y1, r.Y2)), mxy = max(r.Y0, max(r.Y1, r.Y2));
    r.ia = 1.0f / (float)area2;
    r.z0 = Z[id0]; r.dz1 = Z[id1] - Z[id0]; r.dz2 = Z[id2] - Z[id0];
    r.zoff = 0.0f;
    if (a.use_offset) {
      double dY1 = (double)(r.Y1 - r.Y0), dY2 = (double)(r.Y2 - r.Y0);
      double dX1 = (double)(r.X1 - r.X0), dX2 = (double)(r.X2 - r.X0);
      double nx = __dsub_rn(__dmul_rn((double)r.dz1, dY2), __dmul_rn((double)r.dz2, dY1));
      double ny = __dsub_rn(__dmul_rn((double)r.dz2, dX1), __dmul_rn((double)r.dz1, dX2));
      double dzdx = __dmul_rn(__ddiv_rn(nx, (double)area2), (double)SGI_SUBPIX);
      double dzdy = __dmul_rn(__ddiv_rn(ny, (double)area2), (double)SGI_SUBPIX);
      float m = (float)fmax(fabs(dzdx), fabs(dzdy));
      float zmax = fmaxf(Z[id0], fmaxf(Z[id1], Z[id2]));
      float rr = 0.0f;
      if (zmax > 0.0f) {
        int eb = (int)((__float_as_uint(zmax) >> 23) & 255u);
        if (eb > 23 && eb < 255) rr = __uint_as_float((uint32_t)(eb - 23) << 23);
      }
      r.zoff = a.factor * m + a.units * rr;
    }
    int px0 = (mnx - SGI_SUBPIX / 2 + (SGI_SUBPIX - 1)) >> 8, px1 = (mxx - SGI_SUBPIX / 2) >> 8;
    int py0 = (mny - SGI_SUBPIX / 2 + (SGI_SUBPIX - 1)) >> 8, py1 = (mxy - SGI_SUBPIX / 2) >> 8;
    px0 = max(px0, 0); py0 = max(py0, 0); px1 = min(px1, a.W - 1); py1 = min(py1, a.H - 1);
    if (px0 > px1 || py0 > py1) { invalidate(&a.rec[slot]); continue; }
    r.px0 = (int16_t)px0; r.py0 = (int16_t)py0; r.px1 = (int16_t)px1; r.py1 = (int16_t)py1;
    r.prim_front = ((t * 8 + (f - 1)) << 1) | front;
    r.pad0 = was_clipped ? 1 : 0;      // 0: attributes come straight from the source vertices (order in pad1)
    r.pad1 = (id1 == 1) ? 0 : 1;        // unclipped only: 1 = vertices 1 and 2 were swapped to make the record CCW
    a.rec[slot] = r;
    if (f == 1) first = r;
    if (a.attr) {
      SgiRecAttr q;
      const int ids[3] = {id0, id1, id2};
      const int src[3] = {i0, i1, i2};
      q.pad = 0.0f; q.pad2 = 0.0f;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        q.iw[k] = IW[ids[k]];
        const float b0 = poly[ids[k]].b0, b1 = poly[ids[k]].b1, b2 = poly[ids[k]].b2;
#pragma unroll
        for (int c = 0; c < 9; c++) {
          const float* arr = (c < 3) ? a.xyz : (c < 6 ? a.nrm : a.rgb);
          const int cc = c % 3;
          float v = 0.0f;
          if (arr) {
            if (!was_clipped) v = arr[3 * (size_t)src[ids[k]] + cc];       // the source vertex itself (ids[] = 0,1,2 or 0,2,1)
            else {
              const float s0 = arr[3 * (size_t)i0 + cc], s1 = arr[3 * (size_t)i1 + cc], s2 = arr[3 * (size_t)i2 + cc];
              v = (b0 * s0 + b1 * s1) + b2 * s2;
            }
          }
          if (c < 6) q.A[k][c] = v; else q.C[k][cc] = v;
        }
      }
      a.attr[slot] = q;
      if (a.uvrec) {
        SgiRecUV w;
        w.pad[0] = w.pad[1] = w.pad[2] = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float b0 = poly[ids[k]].b0, b1 = poly[ids[k]].b1, b2 = poly[ids[k]].b2;
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            float v;
            if (!was_clipped) v = a.uv[3 * (size_t)src[ids[k]] + cc];
            else {
              const float s0 = a.uv[3 * (size_t)i0 + cc], s1 = a.uv[3 * (size_t)i1 + cc], s2 = a.uv[3 * (size_t)i2 + cc];
              v = (b0 * s0 + b1 * s1) + b2 * s2;
            }
            w.U[k][cc] = v;
          }
        }
        a.uvrec[slot] = w;
      }
    }
  }
  return n - 2;
}

// the clipping route of setup_triangle, kept out of line: its polygon arrays live in local memory
__device__ __noinline__ int setup_triangle_clipped(const SetupBinArgs& a, int t, int i0, int i1, int i2, CV v0, CV v1, CV v2, SgiRec& first, int& base) {
  CV poly[10];
  poly[0] = v0; poly[1] = v1; poly[2] = v2;
  bool was_clipped;
  int n = clip_polygon(poly, 3, was_clipped, a.no_far_clip);
  if (n < 3) { invalidate(&a.rec[t]); return 0; }
  return emit_records<10>(a, t, i0, i1, i2, poly, n, was_clipped, first, base);
}

// Set-up of source triangle t: writes its record (slot t) and the records of the extra fan triangles clipping produced
// (consecutive slots from `base`), valid or invalidated.  Returns the number of fan triangles (0 = rejected; `first` is then
// invalid) and the record of slot t in registers for the binning that follows.
__device__ __forceinline__ int setup_triangle(const SetupBinArgs& a, int t, SgiRec& first, int& base) {
  first.prim_front = -1;
  base = -1;
  a.ovf_base[t] = -1;
  int i0 = a.idx[3 * t], i1 = a.idx[3 * t + 1], i2 = a.idx[3 * t + 2];
  // a repeated index is a zero-area triangle whatever the matrix (the silhouette pass zeroes the index triples of the quads it
  // drops: two thirds of a shadow-volume pass's slots): nothing to set up
  if (i0 == i1 || i1 == i2 || i0 == i2) return 0;
  CV poly[3];
  poly[0] = xform(a.mvp, a.xyz[3 * (size_t)i0], a.xyz[3 * (size_t)i0 + 1], a.xyz[3 * (size_t)i0 + 2]);
  poly[1] = xform(a.mvp, a.xyz[3 * (size_t)i1], a.xyz[3 * (size_t)i1 + 1], a.xyz[3 * (size_t)i1 + 2]);
  poly[2] = xform(a.mvp, a.xyz[3 * (size_t)i2], a.xyz[3 * (size_t)i2 + 1], a.xyz[3 * (size_t)i2 + 2]);
  poly[0].b0 = 1; poly[0].b1 = 0; poly[0].b2 = 0;
  poly[1].b0 = 0; poly[1].b1 = 1; poly[1].b2 = 0;
  poly[2].b0 = 0; poly[2].b1 = 0; poly[2].b2 = 1;
  {  // trivial reject against the true frustum
    int o0 = 0, o1 = 0, o2 = 0, o3 = 0, o4 = 0, o5 = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const CV& v = poly[k];
      if (!(v.w + v.z >= 0.0f)) o0++;
      if (!(v.w - v.z >= 0.0f)) o1++;
      if (!(v.w + v.x >= 0.0f)) o2++;
      if (!(v.w - v.x >= 0.0f)) o3++;
      if (!(v.w + v.y >= 0.0f)) o4++;
      if (!(v.w - v.y >= 0.0f)) o5++;
    }
    if (a.no_far_clip) o1 = 0;
    if (o0 == 3 || o1 == 3 || o2 == 3 || o3 == 3 || o4 == 3 || o5 == 3) { invalidate(&a.rec[t]); return 0; }
  }
  // all three vertices inside the near / far planes and the guard band: nothing to clip (clip_polygon would return the triangle as it is)
  bool inside = true;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const CV& v = poly[k];
    inside = inside && (v.w + v.z >= 0.0f) && (a.no_far_clip || (v.w - v.z >= 0.0f)) && (SGI_GUARD * v.w + v.x >= 0.0f) && (SGI_GUARD * v.w - v.x >= 0.0f) &&
             (SGI_GUARD * v.w + v.y >= 0.0f) && (SGI_GUARD * v.w - v.y >= 0.0f);
  }
  if (inside) return emit_records<3>(a, t, i0, i1, i2, poly, 3, false, first, base);
  return setup_triangle_clipped(a, t, i0, i1, i2, poly[0], poly[1], poly[2], first, base);
}


// ---- binning -------------------------------------------------------------------------------------------------
// conservative triangle / tile overlap: for each edge evaluate at the tile corner that maximises it
__device__ __forceinline__ bool tile_overlaps(const SgiRec& r, int tx, int ty, int W, int H) {
  long long cx0 = (long long)(tx << SGI_TILE_LOG2) * SGI_SUBPIX + SGI_SUBPIX / 2;
  long long cy0 = (long long)(ty << SGI_TILE_LOG2) * SGI_SUBPIX + SGI_SUBPIX / 2;
  long long cx1 = cx0 + (long long)(SGI_TILE - 1) * SGI_SUBPIX, cy1 = cy0 + (long long)(SGI_TILE - 1) * SGI_SUBPIX;
  const int XA[3] = {r.X1, r.X2, r.X0}, YA[3] = {r.Y1, r.Y2, r.Y0};
  const int XB[3] = {r.X2, r.X0, r.X1}, YB[3] = {r.Y2, r.Y0, r.Y1};
#pragma unroll
  for (int e = 0; e < 3; e++) {
    long long dx = (long long)XB[e] - XA[e], dy = (long long)YB[e] - YA[e];
    long long px = (dy < 0) ? cx1 : cx0;      // coefficient of px is -dy
    long long py = (dx > 0) ? cy1 : cy0;      // coefficient of py is  dx
    long long v = dx * (py - YA[e]) - dy * (px - XA[e]);
    if (v < 0) return false;
  }
  return true;
}

// Shadow volumes: a prism triangle whose depth over the whole tile is behind the tile's farthest scene depth cannot
// pass the depth test anywhere in it (z-pass counting), so the pair is not listed at all.  The interpolated depth is
// affine in the pixel position: over the tile it is smallest at one of the four corner pixels (clamped to the viewport);
// evaluated with the fragment formula and lowered by a slack that covers its fp32 rounding (the terms b*dz can be large
// at corners outside the triangle).  Conservative: never removes a fragment that would have passed.  Deterministic, so
// the counting and the filling walk of the binner agree.
__device__ __forceinline__ bool tile_behind_scene(const SgiRec& r, int tx, int ty, int W, int H, const unsigned int* __restrict__ tile_zmax, int tiles_x) {
  if (!tile_zmax) return false;
  const unsigned int bound = __ldg(&tile_zmax[ty * tiles_x + tx]);
  if (bound >= 0x3F800000u) return false;                       // background in the tile: everything in front of 1.0 counts
  const int x0 = tx << SGI_TILE_LOG2, y0 = ty << SGI_TILE_LOG2;
  const int x1 = min(x0 + SGI_TILE - 1, W - 1), y1 = min(y0 + SGI_TILE - 1, H - 1);
  float zmin = 2.0f;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const int PX = ((c & 1) ? x1 : x0) * SGI_SUBPIX + SGI_SUBPIX / 2, PY = ((c & 2) ? y1 : y0) * SGI_SUBPIX + SGI_SUBPIX / 2;
    const long long E1 = (long long)(r.X0 - r.X2) * (long long)(PY - r.Y2) - (long long)(r.Y0 - r.Y2) * (long long)(PX - r.X2);
    const long long E2 = (long long)(r.X1 - r.X0) * (long long)(PY - r.Y0) - (long long)(r.Y1 - r.Y0) * (long long)(PX - r.X0);
    const float t1 = ((float)E1 * r.ia) * r.dz1, t2 = ((float)E2 * r.ia) * r.dz2;
    const float zc = ((r.z0 + t1) + t2) + r.zoff - (1.0e-6f + 5.0e-7f * (fabsf(t1) + fabsf(t2) + fabsf(r.zoff)));
    zmin = fminf(zmin, zc);
  }
  if (!(zmin > 0.0f)) return false;                             // NaN / negative bounds never cull
  return __float_as_uint(fminf(zmin, 1.0f)) > bound;
}

// stencil = count mod 256 (an 8-bit stencil buffer with GL_INCR_WRAP / GL_DECR_WRAP), after all list segments have been added
__global__ void __launch_bounds__(256) k_sv_stencil(const int32_t* __restrict__ count, uint8_t* __restrict__ stencil, int W, int rx0, int ry0, int rx1) {
  const int x = rx0 + blockIdx.x * 256 + threadIdx.x, y = ry0 + blockIdx.y;
  if (x >= rx1) return;
  const size_t o = (size_t)y * W + x;
  stencil[o] = (uint8_t)((unsigned int)count[o] & 255u);
}

// largest scene depth of every 64x64 tile (pixels outside the viewport ignored)
__global__ void __launch_bounds__(256) k_tile_zmax(const float* __restrict__ depth, int W, int H, int tiles_x, unsigned int* __restrict__ out) {
  const int tx = blockIdx.x, ty = blockIdx.y;
  float m = 0.0f;
  for (int q = threadIdx.x; q < SGI_TILE * SGI_TILE; q += 256) {
    const int x = (tx << SGI_TILE_LOG2) + (q & (SGI_TILE - 1)), y = (ty << SGI_TILE_LOG2) + (q >> SGI_TILE_LOG2);
    if (x < W && y < H) m = fmaxf(m, __ldg(&depth[(size_t)y * W + x]));
  }
  unsigned int v = __float_as_uint(fminf(fmaxf(m, 0.0f), 1.0f));
  v = __reduce_max_sync(0xffffffffu, v);
  __shared__ unsigned int red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) v = max(v, red[k]);
    out[ty * tiles_x + tx] = v;
  }
}

#ifndef SGI_BIG_TILES
#define SGI_BIG_TILES 256
#endif

// Set-up and binning in one pass, one thread per source triangle, triangles in mesh order (all loads and stores coalesced,
// neighbouring triangles share their vertices in L1).  A record overlapping <= 4 tiles is appended by its thread with all (up
// to 4) atomics in flight before the first dependent store.  Larger records are parked in shared memory and, after a barrier,
// flattened into one (record, tile) index space that the CTA's 128 threads walk together, 4 pairs per thread per trip: meshes
// list their large triangles (floors, walls) consecutively, and left to their own threads a few warps would walk thousands
// of pairs while the others idle.  The extra fan triangles of a clipped source triangle (rare) take the same two routes,
// re-read from the records just written.
// Lists are fixed-capacity segments: position = atomicAdd(cursor of the tile).  Entries beyond the capacity go to a shared
// spill list of (tile, record) pairs that the tile kernel scans for the tiles that overflowed, so a frame whose triangles
// gather in one tile (an object crossing a tile corner changes the longest list by 4x) still renders completely; only
// when the spill list itself is full is the frame reported (SGI_ERR_OVERFLOW) and re-run with larger lists.
#define SGI_SB_THREADS 128
#define SGI_SB_QCAP 256
struct BinRec {                  // what the tile walk of a large record needs (64 B)
  int X0, Y0, X1, Y1, X2, Y2;
  float z0, dz1, dz2, ia, zoff;
  int slot, bx0, by0, bw, nt;
};
__device__ __forceinline__ void list_append(const SetupBinArgs& a, int tile, int pos, int slot) {
  if (pos < a.cap) a.pairs[(size_t)tile * a.cap + pos] = slot;
  else {
    const int o = atomicAdd(&a.counters[5], 1);
    if (o < a.spill_cap) a.spill[o] = make_int2(tile, slot);
  }
}

__global__ void __launch_bounds__(SGI_SB_THREADS) k_setup_bin(const SetupBinArgs a) {
  __shared__ BinRec q[SGI_SB_QCAP];
  __shared__ int q_pref[SGI_SB_QCAP + 1];
  __shared__ int q_n, warp_sum[SGI_SB_THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int W = a.W, H = a.H;
  if (tid == 0) q_n = 0;
  __syncthreads();
  const int t = blockIdx.x * SGI_SB_THREADS + tid;
  SgiRec r;
  r.prim_front = -1;
  int base = -1, nf = 0;
  if (t < a.T) nf = setup_triangle(a, t, r, base);
  // one record: un-binned big list / shared queue of large records / this thread alone, 4 tiles at a time with plain atomics
  auto bin_plain = [&](const SgiRec& rr, int slot, bool small_done) {
    if (rr.prim_front < 0) return;
    const int bx0 = max((int)rr.px0 >> SGI_TILE_LOG2, a.tx0), by0 = max((int)rr.py0 >> SGI_TILE_LOG2, a.ty0);
    const int bx1 = min((int)rr.px1 >> SGI_TILE_LOG2, a.tx1), by1 = min((int)rr.py1 >> SGI_TILE_LOG2, a.ty1);
    if (bx0 > bx1 || by0 > by1) return;
    const int bw = bx1 - bx0 + 1, nt = bw * (by1 - by0 + 1);
    if (nt <= 4 && small_done) return;                   // appended by the aggregated path below
    if (nt > SGI_BIG_TILES) atomicAdd(&a.counters[7], 1);        // statistic for the host: is k_bin_big worth launching for passes like this one
    if (nt > a.big_tiles) {            // e.g. the floor: listing it in thousands of tiles is not this thread's job (tile kernel or k_bin_big)
      if (nt > a.huge_tiles) a.big_list[a.big_cap - 1 - atomicAdd(&a.counters[6], 1)] = slot;
      else a.big_list[atomicAdd(&a.counters[3], 1)] = slot;
      return;
    }
    int k = -1;
    if (nt > 4) { k = atomicAdd(&q_n, 1); if (k >= SGI_SB_QCAP) k = -1; }
    if (k >= 0) {
      BinRec b;
      b.X0 = rr.X0; b.Y0 = rr.Y0; b.X1 = rr.X1; b.Y1 = rr.Y1; b.X2 = rr.X2; b.Y2 = rr.Y2;
      b.z0 = rr.z0; b.dz1 = rr.dz1; b.dz2 = rr.dz2; b.ia = rr.ia; b.zoff = rr.zoff;
      b.slot = slot; b.bx0 = bx0; b.by0 = by0; b.bw = bw; b.nt = nt;
      q[k] = b;
      return;
    }
    for (int k0 = 0; k0 < nt; k0 += 4) {
      int tiles[4], pos[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        tiles[u] = -1;
        if (k0 + u < nt) {
          const int ty = by0 + (k0 + u) / bw, tx = bx0 + (k0 + u) % bw;
          if ((nt == 1 || tile_overlaps(rr, tx, ty, W, H)) && !tile_behind_scene(rr, tx, ty, W, H, a.tile_zmax, a.tiles_x)) tiles[u] = ty * a.tiles_x + tx;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (tiles[u] >= 0) pos[u] = atomicAdd(&a.tile_cnt[tiles[u]], 1);
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (tiles[u] >= 0) list_append(a, tiles[u], pos[u], slot);
    }
  };
  // ---- the triangle's own record when it overlaps <= 4 tiles (the common case): neighbouring triangles of a mesh land in the
  //      same tile, so the lanes of a warp that append to the same list share ONE atomic (match_any), which takes the
  //      serialisation of same-address atomics on the lists of dense tiles off the critical path
  {
    int tiles[4] = {-1, -1, -1, -1};
    if (r.prim_front >= 0) {
      const int bx0 = max((int)r.px0 >> SGI_TILE_LOG2, a.tx0), by0 = max((int)r.py0 >> SGI_TILE_LOG2, a.ty0);
      const int bx1 = min((int)r.px1 >> SGI_TILE_LOG2, a.tx1), by1 = min((int)r.py1 >> SGI_TILE_LOG2, a.ty1);
      const int bw = bx1 - bx0 + 1, nt = (bx0 <= bx1 && by0 <= by1) ? bw * (by1 - by0 + 1) : 0;
      if (nt >= 1 && nt <= 4) {
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (u < nt) {
            const int ty = by0 + u / bw, tx = bx0 + u % bw;
            if ((nt == 1 || tile_overlaps(r, tx, ty, W, H)) && !tile_behind_scene(r, tx, ty, W, H, a.tile_zmax, a.tiles_x)) tiles[u] = ty * a.tiles_x + tx;
          }
      }
    }
    unsigned grp[4];
    int pos[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {                        // all four atomics of the warp in flight before the first dependent store
      grp[u] = __match_any_sync(0xffffffffu, tiles[u]);
      pos[u] = 0;
      if (tiles[u] >= 0 && lane == __ffs(grp[u]) - 1) pos[u] = atomicAdd(&a.tile_cnt[tiles[u]], __popc(grp[u]));
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      pos[u] = __shfl_sync(0xffffffffu, pos[u], __ffs(grp[u]) - 1);
      if (tiles[u] >= 0) list_append(a, tiles[u], pos[u] + __popc(grp[u] & ((1u << lane) - 1u)), t);
    }
  }
  bin_plain(r, t, true);                                 // big / large: un-binned list or shared queue
  for (int f = 1; f < nf; f++) { const int slot = base + f - 1; const SgiRec rr = a.rec[slot]; bin_plain(rr, slot, false); }
  __syncthreads();
  // ---- larger records of this CTA, flattened: inclusive prefix of their tile counts, then every thread takes pairs
  const int nq = min(q_n, SGI_SB_QCAP);
  if (nq == 0) return;
  if (a.few_tiles) {
    // A pass of few tiles (shadow volumes at 640x480: 80 tiles, 1.5 M pairs): the appends of the whole GPU meet on a few dozen
    // cursors, and the pair walk below spends its time waiting for contended atomics (45 % of the kernel's stall samples).  Here
    // the walk is tile-major - a lane per record, all lanes on the same tile - so a warp reserves its entries of a tile with ONE
    // atomic.  More (cheap) overlap tests, a thirtieth of the atomics.  Four tiles in flight per trip.
    const int gx = a.tx1 - a.tx0 + 1, ntile = gx * (a.ty1 - a.ty0 + 1);
    for (int rb = (tid >> 5) * 32; rb < nq; rb += SGI_SB_THREADS) {
      const int ri = rb + lane;
      const bool live = ri < nq;
      const BinRec b = q[live ? ri : 0];
      SgiRec rr;
      rr.X0 = b.X0; rr.Y0 = b.Y0; rr.X1 = b.X1; rr.Y1 = b.Y1; rr.X2 = b.X2; rr.Y2 = b.Y2;
      rr.z0 = b.z0; rr.dz1 = b.dz1; rr.dz2 = b.dz2; rr.ia = b.ia; rr.zoff = b.zoff;
      const int bx1 = b.bx0 + b.bw - 1, by1 = b.by0 + b.nt / b.bw - 1;
      int cx = a.tx0, cy = a.ty0;                                  // the tile of step t, advanced without a division
      for (int t0 = 0; t0 < ntile; t0 += 4) {
        unsigned hits[4]; int pos[4], tl[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int t = t0 + u, ty = cy, tx = cx;
          if (++cx > a.tx1) { cx = a.tx0; cy++; }
          const bool hit = live && t < ntile && tx >= b.bx0 && tx <= bx1 && ty >= b.by0 && ty <= by1 && tile_overlaps(rr, tx, ty, W, H) &&
                           !tile_behind_scene(rr, tx, ty, W, H, a.tile_zmax, a.tiles_x);
          hits[u] = __ballot_sync(0xffffffffu, hit);
          tl[u] = ty * a.tiles_x + tx;
          pos[u] = 0;
          if (hits[u] && lane == __ffs(hits[u]) - 1) pos[u] = atomicAdd(&a.tile_cnt[tl[u]], __popc(hits[u]));
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (hits[u]) {
            const int p0 = __shfl_sync(0xffffffffu, pos[u], __ffs(hits[u]) - 1);
            if ((hits[u] >> lane) & 1u) list_append(a, tl[u], p0 + __popc(hits[u] & ((1u << lane) - 1u)), b.slot);
          }
      }
    }
    return;
  }
  {
    const int i0 = 2 * tid, i1 = 2 * tid + 1;
    const int c0 = i0 < nq ? q[i0].nt : 0, c1 = i1 < nq ? q[i1].nt : 0;
    int incl = c0 + c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
    if (lane == 31) warp_sum[tid >> 5] = incl;
    __syncthreads();
    int off = 0;
    for (int w = 0; w < (tid >> 5); w++) off += warp_sum[w];
    incl += off;
    q_pref[i0 + 1] = incl - c1; q_pref[i1 + 1] = incl;       // q_pref[i + 1] = pairs of records 0..i
    if (tid == 0) q_pref[0] = 0;
    __syncthreads();
  }
  const int total = q_pref[nq];
  for (int p0 = 0; p0 < total; p0 += 4 * SGI_SB_THREADS) {
    int tiles[4], pos[4], slots[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int p = p0 + u * SGI_SB_THREADS + tid;
      tiles[u] = -1;
      if (p < total) {
        int lo = 0, hi = nq;                                 // owner = the record i with q_pref[i] <= p < q_pref[i + 1]
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (q_pref[mid] <= p) lo = mid; else hi = mid; }
        const BinRec& b = q[lo];
        const int k = p - q_pref[lo];
        const int ty = b.by0 + k / b.bw, tx = b.bx0 + k % b.bw;
        SgiRec rr;
        rr.X0 = b.X0; rr.Y0 = b.Y0; rr.X1 = b.X1; rr.Y1 = b.Y1; rr.X2 = b.X2; rr.Y2 = b.Y2;
        rr.z0 = b.z0; rr.dz1 = b.dz1; rr.dz2 = b.dz2; rr.ia = b.ia; rr.zoff = b.zoff;
        slots[u] = b.slot;
        if (tile_overlaps(rr, tx, ty, W, H) && !tile_behind_scene(rr, tx, ty, W, H, a.tile_zmax, a.tiles_x)) tiles[u] = ty * a.tiles_x + tx;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (tiles[u] >= 0) pos[u] = atomicAdd(&a.tile_cnt[tiles[u]], 1);
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (tiles[u] >= 0) list_append(a, tiles[u], pos[u], slots[u]);
  }
}

// The records the set-up kernel did not bin (floors, walls, at an 8192^2 map every triangle of some size), binned after all.
// With thousands of tiles per pass, every tile CTA testing every big record is the larger cost (a city under an 8192^2 map: 394
// records beyond 256 tiles x 16384 tiles = 6.5 M record tests and 25 KB of record reads per tile, where 96 % of the tiles hold
// nothing else), and the set-up kernel's own walk of the mid-sized ones leaves a few of its CTAs with most of the work.  Here the
// work is spread over the whole GPU: records up to `huge_tiles` tiles are dealt to the warps round robin (a lane per tile of the
// bounding box), the larger ones are walked by all CTAs together in chunks of 256 tiles.  Exact triangle / tile test as in the
// set-up kernel; k_order then reports no big records to the tile kernel.  Launched only for passes of many tiles (option
// "tile_bin_big"); small passes keep the per-tile test.
__global__ void __launch_bounds__(256) k_bin_big(const SetupBinArgs a) {
  SGI_GRID_DEP_WAIT();                                          // the set-up kernel's big lists and cursors
  const int nmid = a.counters[3], nhuge = a.counters[6];
  const int lane = threadIdx.x & 31, warp = (blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = gridDim.x * 8;
  auto bin_range = [&](const SgiRec& rr, int slot, int bx0, int by0, int bw, int k0, int k1, int step, int first) {
    for (int k = k0 + first; k < k1; k += step) {
      const int ty = by0 + k / bw, tx = bx0 + k % bw;
      if (tile_overlaps(rr, tx, ty, a.W, a.H) && !tile_behind_scene(rr, tx, ty, a.W, a.H, a.tile_zmax, a.tiles_x)) {
        const int tile = ty * a.tiles_x + tx;
        list_append(a, tile, atomicAdd(&a.tile_cnt[tile], 1), slot);
      }
    }
  };
  for (int b = warp; b < nmid; b += nwarps) {
    const int slot = a.big_list[b];
    const SgiRec rr = a.rec[slot];
    const int bx0 = max((int)rr.px0 >> SGI_TILE_LOG2, a.tx0), by0 = max((int)rr.py0 >> SGI_TILE_LOG2, a.ty0);
    const int bx1 = min((int)rr.px1 >> SGI_TILE_LOG2, a.tx1), by1 = min((int)rr.py1 >> SGI_TILE_LOG2, a.ty1);
    if (bx0 > bx1 || by0 > by1) continue;
    const int bw = bx1 - bx0 + 1;
    bin_range(rr, slot, bx0, by0, bw, 0, bw * (by1 - by0 + 1), 32, lane);
  }
  for (int h = 0; h < nhuge; h++) {
    const int slot = a.big_list[a.big_cap - 1 - h];
    const SgiRec rr = a.rec[slot];
    const int bx0 = max((int)rr.px0 >> SGI_TILE_LOG2, a.tx0), by0 = max((int)rr.py0 >> SGI_TILE_LOG2, a.ty0);
    const int bx1 = min((int)rr.px1 >> SGI_TILE_LOG2, a.tx1), by1 = min((int)rr.py1 >> SGI_TILE_LOG2, a.ty1);
    if (bx0 > bx1 || by0 > by1) continue;
    const int bw = bx1 - bx0 + 1, nt = bw * (by1 - by0 + 1);
    // chunk c of this record belongs to CTA (c + h) mod grid: the first chunks of successive records go to different CTAs
    for (int c = (int)((blockIdx.x + gridDim.x - h % gridDim.x) % gridDim.x); c * 256 < nt; c += gridDim.x)
      bin_range(rr, slot, bx0, by0, bw, c * 256, min(nt, c * 256 + 256), 256, threadIdx.x);
  }
}

// After the binner: one CTA turns the list lengths into the work items of the tile kernel (one CTA each).
//  * list length the tile kernel reads = min(cursor, capacity); the longest list ever wanted goes to the host-mapped flag word
//    (the host sizes the capacity from it), a cursor beyond the capacity raises the sticky overflow flag;
//  * the binner's live counters (append cursors, clipped-extra slots, big-triangle count) are snapshotted and zeroed here,
//    so the next pass on this scratch set needs no memset;
//  * adaptive subdivision: a tile whose list is much longer than the even share of the pass (total / (8 x SM count CTAs))
//    is split into 4 sub-tiles of 32x32 or 16 of 16x16 pixels, each rasterised by its own CTA from the same list, so
//    that hot tiles (a dense object in a few tiles, shadow-volume prisms at low resolution) do not serialise the pass
//    on one SM.  The item count is capped by the launched grid (max_items): the threshold doubles until it fits.
//  * launch order: busiest first (counting sort on a half-octave bucket of the item's weight), so that the long items
//    start early and the short ones fill the tail (longest-processing-time-first; the order is irrelevant to the result).
// item = tile | level << 20 | sub << 22, level 0/1/2 = 64/32/16-pixel region, sub = sy * (1 << level) + sx.
// Histogram and scatter use one shared-memory atomic per distinct key per warp (__match_any_sync): most tiles of a pass fall
// into two or three buckets, and same-address shared atomics serialise (the previous form of this kernel, one atomic per
// tile, took 36 us on the 16 384 tiles of an 8192^2 map).
__device__ __forceinline__ int split_level(int c, int w, int max_level = 2) {
  if (max_level >= 3 && c > 32 * (long long)w) return 3;
  return c > 8 * (long long)w ? 2 : (c > 2 * (long long)w ? 1 : 0);
}
__device__ __forceinline__ int weight_bucket(int c) {
  const int l = 31 - __clz(c | 1);
  return c < 2 ? c : 2 * l + ((c >> (l - 1)) & 1);
}
struct OrderArgs {
  int32_t* tile_cnt; int n_tiles; int cap; int spill_cap;
  int32_t* counters; int32_t* snap; volatile int32_t* h_flags; int32_t* d_sticky; int size_class;
  int2* order; int tiles_x, tx0, ty0, gx, gy, busiest_first, max_items, split_floor, n_sm;
  unsigned int* mm_min; unsigned int* mm_max; int mm_n;     // depth pass feeding the min-max cull: block extrema reset here (1.0 / 0.0)
  int big_binned;                                            // k_bin_big ran: the tile kernel has no big list to test
  int big_work;                                              // records x tiles from which k_bin_big pays (option "tile_bin_big_work")
  int seg_split, max_level;                                  // stencil pass: hot tiles shared by list segment, up to 4^3 per tile, against a finer even share
};
#define SGI_ORDER_KEYS 256      // 64 weight buckets x 4 (3 levels used)
#define SGI_ORDER_REG 16        // tiles per thread held in registers (grids up to 16 384 tiles: an 8192^2 map); larger grids re-read
// (one global round trip on the critical path: the cursors and the binner's counters are all loaded up front; everything
//  else happens in registers and shared memory; the work items are fire-and-forget stores)
// SEG: the stencil pass's list-segment sharing (4^3 per tile, finer share) - a template parameter, and the big-record statistic is
// handled by a thread of another warp: with them in the main path the kernel spilled (64 registers are all 1024 threads get) and
// took 16 instead of 12 us on the c2 passes
template <bool SEG>
__global__ void __launch_bounds__(1024) k_order(const OrderArgs a) {
  constexpr int MAXLV = SEG ? 3 : 2;
  __shared__ int hist[SGI_ORDER_KEYS];
  __shared__ int red_sum[32], red_max[32];
  __shared__ int s_items, s_w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nl = a.gx * a.gy;
  const bool in_regs = nl <= 1024 * SGI_ORDER_REG;
  for (int k = tid; k < SGI_ORDER_KEYS; k += 1024) hist[k] = 0;
  SGI_GRID_DEP_WAIT();                                 // the binner's cursors and counters from here on
  if (a.mm_min) for (int i = tid; i < a.mm_n; i += 1024) { a.mm_min[i] = 0x3F800000u; a.mm_max[i] = 0u; }
  int c0 = 0, c3 = 0, c5 = 0, st_long = 0, st_tot = 0;
  if (tid == 0) { c0 = a.counters[0]; c3 = a.counters[3]; c5 = a.counters[5]; st_long = a.d_sticky[a.size_class]; st_tot = a.d_sticky[4 + a.size_class]; }
  if (tid == 32) {
    // is k_bin_big worth launching for passes like this one (records beyond SGI_BIG_TILES x tiles)?  Only a CHANGE of the answer is
    // written to the host-mapped word
    const int c7 = a.counters[7], st_big = a.d_sticky[8 + a.size_class];
    const int want_big = ((long long)c7 * nl >= (long long)a.big_work) ? 1 : 0;
    a.counters[7] = 0;
    if (want_big != st_big) { a.d_sticky[8 + a.size_class] = want_big; a.h_flags[8 + a.size_class] = want_big; }
  }
  // this thread's tiles of the job rectangle, i = tid + 1024 k: tile index and cursor
  int til[SGI_ORDER_REG], cnt[SGI_ORDER_REG];
  {
    // (x, y) of tile i = tid + 1024 k in the rectangle, stepped without a division per tile
    const int qy = 1024 / a.gx, qx = 1024 - qy * a.gx;
    int y = tid / a.gx, x = tid - y * a.gx;
#pragma unroll
    for (int k = 0; k < SGI_ORDER_REG; k++) {
      til[k] = -1; cnt[k] = 0;
      if (k * 1024 < nl) {                             // (warp-uniform)
        if (k * 1024 + tid < nl) {
          til[k] = (a.ty0 + y) * a.tiles_x + a.tx0 + x;
          cnt[k] = a.tile_cnt[til[k]];
        }
        y += qy; x += qx;
        if (x >= a.gx) { x -= a.gx; y++; }
      }
    }
  }
  auto tile_of = [&](int i, int& c) -> int {          // grids beyond the register window: re-read (the cursors stay until the end)
    const int y = i / a.gx, x = i - y * a.gx;
    const int tl = (a.ty0 + y) * a.tiles_x + a.tx0 + x;
    c = a.tile_cnt[tl];
    return tl;
  };
  // ---- totals (tiles outside the job rectangle are never listed: their cursors are zero)
  int sum = 0, mx = 0;
#pragma unroll
  for (int k = 0; k < SGI_ORDER_REG; k++) if (k * 1024 < nl) { sum += min(cnt[k], a.cap); mx = max(mx, cnt[k]); }
  if (!in_regs) for (int i = 1024 * SGI_ORDER_REG + tid; i < nl; i += 1024) { int c; tile_of(i, c); sum += min(c, a.cap); mx = max(mx, c); }
  sum = __reduce_add_sync(0xffffffffu, sum); mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) { red_sum[warp] = sum; red_max[warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    int s2 = __reduce_add_sync(0xffffffffu, red_sum[lane]), m2 = __reduce_max_sync(0xffffffffu, red_max[lane]);
    if (lane == 0) {
      a.snap[0] = c0; a.snap[3] = a.big_binned ? 0 : c3; a.snap[2] = s2; a.snap[5] = min(c5, a.spill_cap);
      a.counters[0] = 0; a.counters[3] = 0; a.counters[5] = 0; a.counters[6] = 0;
      // longest list / largest pair total ever wanted (the host sizes the lists from them).  The running maxima live in device
      // memory and the host-mapped words are only ever WRITTEN: a read of host memory from this single-CTA kernel waits behind
      // whatever DMA traffic is on PCIe at the time (measured: +0.03 ms per pass while a frame is being copied out)
      if (m2 > st_long) { a.d_sticky[a.size_class] = m2; a.h_flags[1 + a.size_class] = m2; }
      if (s2 + c5 > st_tot) { a.d_sticky[4 + a.size_class] = s2 + c5; a.h_flags[4 + a.size_class] = s2 + c5; }
      if (c5 > a.spill_cap) a.h_flags[0] = 1;                   // entries were dropped: the frame is incomplete
      s_w = a.split_floor > 0 ? max(SEG ? a.split_floor / 2 : a.split_floor, s2 / ((SEG ? 32 : 8) * a.n_sm)) : 0x7FFFFFF;
    }
  }
  // list length per tile (what the tile kernel reads) | bit 30: the tile has further entries in the spill list
#pragma unroll
  for (int k = 0; k < SGI_ORDER_REG; k++) if (k * 1024 < nl) cnt[k] = cnt[k] > a.cap ? (a.cap | 0x40000000) : cnt[k];
  auto len_of = [&](int c) -> int { return c > a.cap ? (a.cap | 0x40000000) : c; };
  for (;;) {                                          // largest subdivision that fits the launched grid
    __syncthreads();
    if (tid == 0) s_items = 0;
    __syncthreads();
    const int w = s_w;
    int local = 0;
#pragma unroll
    for (int k = 0; k < SGI_ORDER_REG; k++) if (k * 1024 < nl && til[k] >= 0) local += 1 << (2 * split_level(cnt[k] & 0x3FFFFFFF, w, MAXLV));
    if (!in_regs) for (int i = 1024 * SGI_ORDER_REG + tid; i < nl; i += 1024) { int c; tile_of(i, c); local += 1 << (2 * split_level(min(c, a.cap), w, MAXLV)); }
    local = __reduce_add_sync(0xffffffffu, local);
    if (lane == 0 && local) atomicAdd(&s_items, local);
    __syncthreads();
    if (s_items <= a.max_items || w >= 0x7FFFFFF) break;
    __syncthreads();
    if (tid == 0) s_w = w >= 0x3FFFFFF ? 0x7FFFFFF : 2 * w;
  }
  const int w = s_w;
  // ---- histogram of the items over (weight bucket, level): one shared atomic per distinct key per warp
  auto key_of = [&](int len) -> int { const int c = len & 0x3FFFFFFF; const int lv = split_level(c, w, MAXLV); return (a.busiest_first ? weight_bucket(c >> (SEG ? 2 * lv : lv)) : 0) * 4 + lv; };
  auto hist_step = [&](int key) {
    const unsigned grp = __match_any_sync(0xffffffffu, key);
    if (key >= 0 && lane == __ffs(grp) - 1) atomicAdd(&hist[key], __popc(grp) << (2 * (key & 3)));
  };
#pragma unroll
  for (int k = 0; k < SGI_ORDER_REG; k++)
    if (k * 1024 < nl) hist_step(til[k] >= 0 ? key_of(cnt[k]) : -1);          // (warp-uniform condition)
  for (int i0 = 1024 * SGI_ORDER_REG; i0 < nl; i0 += 1024) {
    int key = -1;
    if (i0 + tid < nl) { int c; tile_of(i0 + tid, c); key = key_of(len_of(c)); }
    hist_step(key);
  }
  __syncthreads();
  if (warp == 0) {                                    // exclusive prefix in descending key order: lane l owns keys 255 - 8 l .. 248 - 8 l
    int v[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) { v[j] = hist[SGI_ORDER_KEYS - 1 - (8 * lane + j)]; tot += v[j]; }
    int incl = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
    int run = incl - tot;
#pragma unroll
    for (int j = 0; j < 8; j++) { hist[SGI_ORDER_KEYS - 1 - (8 * lane + j)] = run; run += v[j]; }
    if (lane == 31) a.snap[4] = min(incl, a.max_items);
  }
  __syncthreads();
  // ---- scatter the work items; re-zero the cursors for the next pass on this scratch set (no memset between passes)
  auto scatter_step = [&](int key, int tile, int len) {
    const unsigned grp = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(grp) - 1;
    int at = 0;
    const int lv = key & 3, nsub = 1 << (2 * lv);
    if (key >= 0 && lane == leader) at = atomicAdd(&hist[key], __popc(grp) * nsub);
    at = __shfl_sync(0xffffffffu, at, leader);
    if (key >= 0) {
      at += __popc(grp & ((1u << lane) - 1u)) * nsub;
      for (int sidx = 0; sidx < nsub; sidx++)
        if (at + sidx < a.max_items) a.order[at + sidx] = make_int2(tile | (lv << 20) | (sidx << 22), len);
      if (len) a.tile_cnt[tile] = 0;
    }
  };
#pragma unroll
  for (int k = 0; k < SGI_ORDER_REG; k++)
    if (k * 1024 < nl) scatter_step(til[k] >= 0 ? key_of(cnt[k]) : -1, til[k], cnt[k]);
  for (int i0 = 1024 * SGI_ORDER_REG; i0 < nl; i0 += 1024) {
    int key = -1, tile = 0, len = 0;
    if (i0 + tid < nl) { int c; tile = tile_of(i0 + tid, c); len = len_of(c); key = key_of(len); }
    scatter_step(key, tile, len);
  }
}

// ---- per-tile rasterisation ----------------------------------------------------------------------------------
struct TileArgs {
  const SgiRec* rec; const SgiRecAttr* attr; const int32_t* ovf_base;
  const int2* tile_order; const int32_t* pairs; int cap;   // work items (item, list length); list of tile t: pairs[t * cap .. + length)
  const int32_t* big_list; const int32_t* counters;      // k_order's snapshot: [3] = number of un-binned big triangles, [4] = work items, [5] = spill entries
  const int2* spill;
  int bulk_flush;                                        // depth tiles: cp.async.bulk row copies (option "tile_bulk_flush")
  int tiles_x, tx0, ty0;
  int W, H, rx0, ry0, rx1, ry1;
  const float* xyz; const float* nrm; const int32_t* idx;
  float* depth; float4* pos4; float4* nrm4;
  const float* rgb; float4* albedo4;
  const float* scene_depth; int depth_func; int32_t* count; uint8_t* stencil;
  float4* mom4; int mom_tech, z_near, z_far; float mq[16], mqt[4];      // MOMENTS
  unsigned int* ids;                                                   // IDS
  int sv_zfail, sv_caps; unsigned long long* frag_counter;             // SVCOUNT: depth-fail mode, capped volumes, optional fragment tally
  int sv_split_lists;                                                  // SVCOUNT: hot tiles are shared by list segment, counts added atomically
  unsigned int* mm_min; unsigned int* mm_max; int mm_w;                // DEPTH: per 32x32-texel block extrema of the map (float bits), or null
  const SgiRecUV* uvrec; SgiTex tex[3];                                 // GBUFFER_RGB: texture select (useTextureForColoring), or null
  int direct_max;                                                       // DEPTH: lists up to this length take the register path (option "tile_direct"), 0 = never
  int static_items, refresh_full_only;                                  // scheduling switches of the work-item loop (options "tile_static_items", "tile_refresh_full")
};

#define ONE_BITS 0x3F800000u

__device__ __forceinline__ bool edge_in(long long e, int dx, int dy) {
  return e > 0 || (e == 0 && (dy < 0 || (dy == 0 && dx < 0)));
}

// coverage + the three edge values of pixel centre (px,py) [pixels]
__device__ __forceinline__ bool cover(int X0, int Y0, int X1, int Y1, int X2, int Y2, int px, int py, long long& E0,
                                      long long& E1, long long& E2) {
  int PX = px * SGI_SUBPIX + SGI_SUBPIX / 2, PY = py * SGI_SUBPIX + SGI_SUBPIX / 2;
  int dx0 = X2 - X1, dy0 = Y2 - Y1;      // edge 0: v1 -> v2 (opposite v0)
  int dx1 = X0 - X2, dy1 = Y0 - Y2;      // edge 1: v2 -> v0
  int dx2 = X1 - X0, dy2 = Y1 - Y0;      // edge 2: v0 -> v1
  E0 = (long long)dx0 * (long long)(PY - Y1) - (long long)dy0 * (long long)(PX - X1);
  if (!edge_in(E0, dx0, dy0)) return false;
  E1 = (long long)dx1 * (long long)(PY - Y2) - (long long)dy1 * (long long)(PX - X2);
  if (!edge_in(E1, dx1, dy1)) return false;
  E2 = (long long)dx2 * (long long)(PY - Y0) - (long long)dy2 * (long long)(PX - X0);
  return edge_in(E2, dx2, dy2);
}

__device__ __forceinline__ float frag_z(float z0, float dz1, float dz2, float ia, float zoff, long long E1, long long E2) {
  float b1 = (float)E1 * ia, b2 = (float)E2 * ia;
  float z = (z0 + b1 * dz1) + b2 * dz2;
  z = z + zoff;
  if (!(z >= 0.0f)) z = 0.0f;
  if (z > 1.0f) z = 1.0f;
  return z;
}

__constant__ int c_rcp16[9] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192};     // ceil(2^16 / n)

// a * b + c with a 32 x 32 -> 64-bit product (IMAD.WIDE with the 64-bit addend).  Inline PTX so that the compiler keeps the
// incremental form: written in C it folds the block step back into the edge function and multiplies 64-bit operands again.
__device__ __forceinline__ long long madw(int a, int b, long long c) {
  long long d;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}

// Edge set-up of one triangle for repeated coverage tests.  With bias_e = (edge e owns its boundary ? 0 : 1) the
// top-left rule `E > 0 || (E == 0 && owns)` is `E - bias >= 0`, and the bias rides for free as the addend of the
// second wide multiply; all three tests collapse to one sign test of e0|e1|e2.  Exactly the integers cover() produces.
struct EdgeSet {
  int X0, Y0, X1, Y1, X2, Y2;
  int dx0, dy0, dx1, dy1, dx2, dy2;
  long long b0, b1, b2;
  __device__ __forceinline__ void init(int x0, int y0, int x1, int y1, int x2, int y2) {
    X0 = x0; Y0 = y0; X1 = x1; Y1 = y1; X2 = x2; Y2 = y2;
    dx0 = X2 - X1; dy0 = Y2 - Y1; dx1 = X0 - X2; dy1 = Y0 - Y2; dx2 = X1 - X0; dy2 = Y1 - Y0;
    b0 = (dy0 < 0 || (dy0 == 0 && dx0 < 0)) ? 0 : 1;
    b1 = (dy1 < 0 || (dy1 == 0 && dx1 < 0)) ? 0 : 1;
    b2 = (dy2 < 0 || (dy2 == 0 && dx2 < 0)) ? 0 : 1;
  }
  // pixel (px,py) [absolute pixels]: true if covered; E1/E2 are the un-biased edge values the depth needs
  __device__ __forceinline__ bool test(int px, int py, long long& E1, long long& E2) const {
    const int PX = px * SGI_SUBPIX + SGI_SUBPIX / 2, PY = py * SGI_SUBPIX + SGI_SUBPIX / 2;
    const long long e0 = (long long)dx0 * (long long)(PY - Y1) - ((long long)dy0 * (long long)(PX - X1) + b0);
    const long long e1 = (long long)dx1 * (long long)(PY - Y2) - ((long long)dy1 * (long long)(PX - X2) + b1);
    const long long e2 = (long long)dx2 * (long long)(PY - Y0) - ((long long)dy2 * (long long)(PX - X0) + b2);
    if ((e0 | e1 | e2) < 0) return false;
    E1 = e1 + b1; E2 = e2 + b2;
    return true;
  }
};

// The tile payload in shared memory uses a row pitch of 72 words, so that the 8x4-pixel blocks the warps
// work on (4 rows 8 banks apart) and full rows (flush) are both free of bank conflicts.
#define SGI_PITCH (SGI_TILE + 8)
#define SGI_BLK_W 8
#define SGI_BLK_H 4

// largest value of edge (a -> b) over the pixel centres of the 8x4 block whose first pixel is (px,py)
__device__ __forceinline__ long long edge_block_max(int Xa, int Ya, int Xb, int Yb, int px, int py) {
  int dx = Xb - Xa, dy = Yb - Ya;
  int sx = (dy < 0) ? px + SGI_BLK_W - 1 : px;      // coefficient of x is -dy
  int sy = (dx > 0) ? py + SGI_BLK_H - 1 : py;      // coefficient of y is  dx
  int PX = sx * SGI_SUBPIX + SGI_SUBPIX / 2, PY = sy * SGI_SUBPIX + SGI_SUBPIX / 2;
  return (long long)dx * (long long)(PY - Ya) - (long long)dy * (long long)(PX - Xa);
}

// Hierarchical depth.  Each 8x4 block of the tile keeps an upper bound `bz` of the depths currently stored in it
// (depths only ever decrease, so any earlier maximum stays an upper bound).  A triangle whose smallest possible depth
// is above that bound cannot change the block and is skipped before any per-pixel work.  The triangle bound is the
// smallest vertex depth plus the polygon offset, lowered by a slack that covers the fp32 rounding of the interpolation
// z = (z0 + b1*dz1) + b2*dz2 (b in [0,1] up to rounding), so culling never removes a fragment that would have passed.
__device__ __forceinline__ unsigned int tri_depth_lower_bound(float z0, float dz1, float dz2, float zoff) {
  float zmin = fminf(z0, fminf(z0 + dz1, z0 + dz2)) + zoff;
  zmin -= 4.0e-6f * (1.0f + fabsf(dz1) + fabsf(dz2) + fabsf(zoff));
  if (!(zmin > 0.0f)) return 0u;                     // also catches NaN
  return __float_as_uint(fminf(zmin, 1.0f));
}
#define SGI_NBLK ((SGI_TILE / SGI_BLK_W) * (SGI_TILE / SGI_BLK_H))     // 8 x 16 = 128 blocks per tile
#define SGI_ZBUCKETS 64          // work items of a chunk are issued nearest-first (counting sort on their depth bound)

// Large triangles of the current chunk, parked in shared memory for the warp-cooperative phase.
template <int NT>
struct TriQueue {
  int X0[NT], Y0[NT], X1[NT], Y1[NT], X2[NT], Y2[NT];
  float z0[NT], dz1[NT], dz2[NT], ia[NT], zoff[NT];
  int meta[NT];
  unsigned int zlo[NT];            // conservative lower bound of every depth this triangle can produce (float bits)
  int box[NT];                     // lx0 | ly0<<8 | lx1<<16 | ly1<<24 (tile-local inclusive bbox)
  int group[4 * NT];               // work items: queue index << 3 | group of 32 blocks (a 64x64 bbox has 128 blocks)
};
#define SGI_SPLIT_EXTRA 1536      // most CTAs a pass may add by subdividing hot tiles
#define SGI_SPLIT_EXTRA_SEG 8192  // ... the stencil pass, whose hot tiles are shared by list segment (up to 64 per tile)
#define SGI_MAX_FULL 8
#define SGI_SMALL_TRI 8           // bbox candidates up to which one thread rasterises the triangle alone

// texture2D on a scene texture: GL_LINEAR, GL_REPEAT, no mipmaps.  The filter's arithmetic is the oracle's definition
// (oracle_raster.c tex_fetch_linear_repeat): weights fract(u*size - 0.5), texels (byte / 255, alpha 1) accumulated 00, 10, 01, 11.
__device__ __forceinline__ float tex_wrap(float f, float size) {
  f = f - floorf(f / size) * size;
  if (!(f < size)) f = 0.0f;
  return f;
}
__device__ __forceinline__ float4 tex_fetch_linear_repeat(const SgiTex& t, float u, float v) {
  const float fw = (float)t.w, fh = (float)t.h;
  const float x = u * fw - 0.5f, y = v * fh - 0.5f;
  const float x0 = floorf(x), y0 = floorf(y), ax = x - x0, ay = y - y0;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float wgt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
    const float fi = tex_wrap(x0 + (float)(k & 1), fw), fj = tex_wrap(y0 + (float)(k >> 1), fh);
    float4 tx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (fi >= 0.0f && fi < fw && fj >= 0.0f && fj < fh) {
      const uchar4 p = __ldg(&t.texels[(size_t)(int)fj * t.w + (size_t)(int)fi]);
      tx = make_float4((float)p.x / 255.0f, (float)p.y / 255.0f, (float)p.z / 255.0f, 1.0f);
    }
    acc.x = acc.x + tx.x * wgt; acc.y = acc.y + tx.y * wgt; acc.z = acc.z + tx.z * wgt; acc.w = acc.w + tx.w * wgt;
  }
  return acc;
}
// GBuffer.frag:11-30 computeFragmentColor with useTextureForColoring == 1
__device__ __forceinline__ float4 fragment_color(const SgiTex (&tex)[3], float u, float v, float b, float r, float g, float bl) {
  int sel = -1;
  if (b > 0.99f && b < 1.001f) sel = 0;
  else if (b > 1.999f && b < 2.001f) sel = 1;
  else if (b > 2.999f && b < 3.001f) sel = 2;
  if (sel >= 0) {
    const SgiTex& t = tex[sel];
    if (t.texels && t.w > 0 && t.h > 0) return tex_fetch_linear_repeat(t, u, v);
    return make_float4(0.f, 0.f, 0.f, 0.f);              // an unbound sampler reads (0,0,0,0)
  }
  return make_float4(r, g, bl, 1.0f);
}

template <int MODE>
struct TileSink {                  // where fragments go: the tile payload in shared memory
  unsigned int* zt; unsigned long long* kt; int* ct; const float* sd; int depth_func;
  int zfail, caps; int* nfrag;     // shadow volumes: depth-fail counting, capped volumes (8 triangles per source triangle), fragment tally
  __device__ __forceinline__ void fragment(int lx, int ly, float z, int meta) const {
    const int p = ly * SGI_PITCH + lx;
    if (MODE == SGI_MODE_DEPTH) {
      const unsigned int zb = __float_as_uint(z);
      if (zb < zt[p]) atomicMin(&zt[p], zb);
    } else if SGI_KEYED(MODE) {
      if (z < 1.0f) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned int)(meta >> 1);
        if (key < kt[p]) atomicMin(&kt[p], key);
      }
    } else {
      // ShadowVolumes/src/main.cpp:160-172: front faces +1 / back faces -1 on fragments that pass the depth test (depth-pass);
      // depth-fail mode: back faces +1 / front faces -1 on fragments that fail it.  Cap triangles (6 and 7 of each group of 8)
      // are tested strictly: a cap fragment coplanar with the visible surface - the surface's own triangle - counts as failing,
      // which is what makes the depth-fail count equal the depth-pass count of the open prisms (shadow volumes)
      (*nfrag)++;
      const float d = sd[p];
      const bool strict = depth_func == SGI_DEPTH_LESS || (caps && ((meta >> 4) & 7) >= 6);
      const bool pass = strict ? (z < d) : (z <= d);
      if (!zfail) { if (pass) atomicAdd(&ct[p], (meta & 1) ? 1 : -1); }
      else if (!pass) atomicAdd(&ct[p], (meta & 1) ? -1 : 1);
    }
  }
};

// Short lists (a city under an 8192^2 map: one to a few dozen triangles per tile, most of them larger than the tile).  The general
// path below pays a fixed price per tile - payload clear, queue, work items, five barriers, flush from shared memory - that is most
// of the kernel where lists are short.  Here every thread owns a patch of 4 x PH texels in registers instead: the list's set-ups
// are parked in shared memory once (one barrier), every thread walks the list, skips triangles whose bounding box or whose edge
// maxima miss its patch, steps the three biased edge values texel by texel (one wide multiply-add each, exact integers) and keeps
// the minimum depth; the patch then goes straight to global memory as 16-byte stores.  No shared-memory payload, no atomics.
// The depths are frag_z() of the same integers in the same order of operations as the general path: identical bits.
template <int NT>
__device__ __forceinline__ void tile_direct_depth(const TileArgs& a, TriQueue<NT>& tq, int ox, int oy, const int32_t* __restrict__ list,
                                                  int nlisted, int nitems) {
  constexpr int PH = SGI_TILE * SGI_TILE / NT / 4;
  const int tid = threadIdx.x;
  uint4 q0, q1, q2;
  int box = 0xFF;                                              // 0xFF: a box no patch overlaps
  unsigned int zlo = 0u;
  if (tid < nitems) {
    const int slot = tid < nlisted ? __ldg(&list[tid]) : __ldg(&a.big_list[tid - nlisted]);
    const uint4* rp = reinterpret_cast<const uint4*>(&a.rec[slot]);
    q0 = __ldg(rp); q1 = __ldg(rp + 1); q2 = __ldg(rp + 2);
    const uint4 q3 = __ldg(rp + 3);
    const int px0 = (int)(short)(q3.x & 0xFFFF), py0 = (int)(short)(q3.x >> 16);
    const int px1 = (int)(short)(q3.y & 0xFFFF), py1 = (int)(short)(q3.y >> 16);
    const int lx0 = max(px0 - ox, 0), ly0 = max(py0 - oy, 0), lx1 = min(px1 - ox, SGI_TILE - 1), ly1 = min(py1 - oy, SGI_TILE - 1);
    if (lx0 <= lx1 && ly0 <= ly1) box = lx0 | (ly0 << 8) | (lx1 << 16) | (ly1 << 24);
    zlo = tri_depth_lower_bound(__uint_as_float(q1.z), __uint_as_float(q1.w), __uint_as_float(q2.x), __uint_as_float(q2.z));
    tq.zlo[tid] = zlo;
  }
  __syncthreads();
  if (tid < nitems) {
    // nearest first (rank by the depth bound, ties by list position): a thread leaves the list at the first triangle that lies
    // wholly behind everything its patch already holds
    int rank = 0;
    for (int j = 0; j < nitems; j++) { const unsigned int o = tq.zlo[j]; rank += (o < zlo || (o == zlo && j < tid)) ? 1 : 0; }
    tq.X0[rank] = (int)q0.x; tq.Y0[rank] = (int)q0.y; tq.X1[rank] = (int)q0.z; tq.Y1[rank] = (int)q0.w; tq.X2[rank] = (int)q1.x; tq.Y2[rank] = (int)q1.y;
    tq.z0[rank] = __uint_as_float(q1.z); tq.dz1[rank] = __uint_as_float(q1.w); tq.dz2[rank] = __uint_as_float(q2.x);
    tq.ia[rank] = __uint_as_float(q2.y); tq.zoff[rank] = __uint_as_float(q2.z);
    tq.box[rank] = box; tq.meta[rank] = (int)zlo;
  }
  __syncthreads();
  const int plx = (tid & 15) << 2, ply = (tid >> 4) * PH;
  unsigned int best[PH][4];
#pragma unroll
  for (int j = 0; j < PH; j++)
#pragma unroll
    for (int i = 0; i < 4; i++) best[j][i] = ONE_BITS;
  const int PX = (ox + plx) * SGI_SUBPIX + SGI_SUBPIX / 2, PY = (oy + ply) * SGI_SUBPIX + SGI_SUBPIX / 2;
  unsigned int pmax = ONE_BITS;                                // largest depth the patch holds
  for (int k = 0; k < nitems; k++) {
    if ((unsigned int)tq.meta[k] > pmax) break;                // this and every later triangle: behind the whole patch
    const int box = tq.box[k];
    if (plx > ((box >> 16) & 0xFF) || plx + 3 < (box & 0xFF) || ply > ((box >> 24) & 0xFF) || ply + PH - 1 < ((box >> 8) & 0xFF)) continue;
    const int X0 = tq.X0[k], Y0 = tq.Y0[k], X1 = tq.X1[k], Y1 = tq.Y1[k], X2 = tq.X2[k], Y2 = tq.Y2[k];
    const int dx0 = X2 - X1, dy0 = Y2 - Y1, dx1 = X0 - X2, dy1 = Y0 - Y2, dx2 = X1 - X0, dy2 = Y1 - Y0;
    const long long b0 = (dy0 < 0 || (dy0 == 0 && dx0 < 0)) ? 0 : 1, b1 = (dy1 < 0 || (dy1 == 0 && dx1 < 0)) ? 0 : 1,
                    b2 = (dy2 < 0 || (dy2 == 0 && dx2 < 0)) ? 0 : 1;
    // biased edge values at the patch's first texel (EdgeSet::test): covered <=> all three >= 0
    long long r0 = (long long)dx0 * (long long)(PY - Y1) - ((long long)dy0 * (long long)(PX - X1) + b0);
    long long r1 = (long long)dx1 * (long long)(PY - Y2) - ((long long)dy1 * (long long)(PX - X2) + b1);
    long long r2 = (long long)dx2 * (long long)(PY - Y0) - ((long long)dy2 * (long long)(PX - X0) + b2);
    {  // the largest value of each edge over the patch (affine: at the corner its coefficients point to) - negative: no texel inside
      const long long m0 = madw(dy0 < 0 ? -dy0 : 0, 3 * SGI_SUBPIX, madw(dx0 > 0 ? dx0 : 0, (PH - 1) * SGI_SUBPIX, r0));
      const long long m1 = madw(dy1 < 0 ? -dy1 : 0, 3 * SGI_SUBPIX, madw(dx1 > 0 ? dx1 : 0, (PH - 1) * SGI_SUBPIX, r1));
      const long long m2 = madw(dy2 < 0 ? -dy2 : 0, 3 * SGI_SUBPIX, madw(dx2 > 0 ? dx2 : 0, (PH - 1) * SGI_SUBPIX, r2));
      if ((m0 | m1 | m2) < 0) continue;
    }
    const float z0 = tq.z0[k], dz1 = tq.dz1[k], dz2 = tq.dz2[k], ia = tq.ia[k], zoff = tq.zoff[k];
#pragma unroll
    for (int j = 0; j < PH; j++) {
      long long c0 = r0, c1 = r1, c2 = r2;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if ((c0 | c1 | c2) >= 0) best[j][i] = min(best[j][i], __float_as_uint(frag_z(z0, dz1, dz2, ia, zoff, c1 + b1, c2 + b2)));
        if (i < 3) { c0 = madw(dy0, -SGI_SUBPIX, c0); c1 = madw(dy1, -SGI_SUBPIX, c1); c2 = madw(dy2, -SGI_SUBPIX, c2); }
      }
      if (j < PH - 1) { r0 = madw(dx0, SGI_SUBPIX, r0); r1 = madw(dx1, SGI_SUBPIX, r1); r2 = madw(dx2, SGI_SUBPIX, r2); }
    }
    pmax = 0u;
#pragma unroll
    for (int j = 0; j < PH; j++)
#pragma unroll
      for (int i = 0; i < 4; i++) pmax = max(pmax, best[j][i]);
  }
  const int x = ox + plx;
  const bool vec_ok = (a.W & 3) == 0 && x >= a.rx0 && x + 3 < a.rx1;
  if (x + 3 < a.rx0 || x >= a.rx1) return;
#pragma unroll
  for (int j = 0; j < PH; j++) {
    const int y = oy + ply + j;
    if (y < a.ry0 || y >= a.ry1) continue;
    float* dst = a.depth + (size_t)y * a.W + x;
    if (vec_ok) *reinterpret_cast<uint4*>(dst) = make_uint4(best[j][0], best[j][1], best[j][2], best[j][3]);
    else {
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (x + i >= a.rx0 && x + i < a.rx1) dst[i] = __uint_as_float(best[j][i]);
    }
  }
}

// One CTA per 64x64 tile.  The tile's triangle list is consumed in chunks of 256: every thread fetches one
// record (all loads of a chunk in flight together); triangles whose bounding box holds <= 16 pixel centres are
// rasterised by that thread on the spot, the others are parked in shared memory and then rasterised
// warp-cooperatively: the bounding box is walked in 8x4 blocks, 32 blocks conservatively tested at once (one
// per lane), surviving blocks rasterised one per trip with one lane per pixel.
// (min blocks = 1024 / NT: 64 registers per thread, i.e. 1024 resident threads per SM whatever the CTA size; left to itself
//  the compiler took 80-94 registers for the small-CTA depth variants and occupancy fell to 768 threads)
template <int MODE, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) k_tile(const TileArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned int* zt = reinterpret_cast<unsigned int*>(smem_raw);
  unsigned long long* kt = reinterpret_cast<unsigned long long*>(smem_raw);
  int* ct = reinterpret_cast<int*>(smem_raw);
  constexpr int NCELL = SGI_TILE * SGI_PITCH;
  constexpr size_t PAYLOAD = SGI_KEYED(MODE) ? (size_t)NCELL * 8 : (size_t)NCELL * 4;
  constexpr size_t SDBYTES = (MODE == SGI_MODE_SVCOUNT) ? (size_t)NCELL * 4 : 0;
  float* sd = reinterpret_cast<float*>(smem_raw + PAYLOAD);
  TriQueue<NT>& tq = *reinterpret_cast<TriQueue<NT>*>(smem_raw + PAYLOAD + SDBYTES);
  __shared__ int next_item, q_count, g_count;
  __shared__ int fc_count, fc_list[SGI_MAX_FULL];   // queued triangles that cover this CTA's whole region (no block walk needed)
  __shared__ unsigned int bz[SGI_NBLK];          // per 8x4 block: upper bound of the stored depths (SV: of the scene depths)
  __shared__ unsigned int zq_min, zq_max;        // depth range of the queued triangles of the chunk
  __shared__ int bucket_cnt[SGI_ZBUCKETS], bucket_pos[SGI_ZBUCKETS];

  const int tid = threadIdx.x, lane = tid & 31;
  SGI_GRID_DEP_WAIT();                                           // k_order's work items (programmatic dependent launch)
  if ((int)blockIdx.x >= a.counters[4]) return;                  // the grid is an upper bound of the item count
  const int2 item2 = a.tile_order[blockIdx.x];                   // work items of k_order, busiest first: (item, list length | spill flag)
  const int item = item2.x;
  const int tile = item & 0xFFFFF, split = (item >> 20) & 3, sub = item >> 22;
  // A hot tile is shared by 4^split CTAs.  Depth and G-buffer passes give each a quarter / a sixteenth of the tile's pixels (all
  // of them scan the whole list).  Stencil counts are sums, so there the LIST is cut instead: every CTA counts its segment of the
  // list over the whole tile and adds its counts to the target (option "sv_split_lists") - no record is scanned twice.
  const bool seg_split = MODE == SGI_MODE_SVCOUNT && a.sv_split_lists && split > 0;
  const int level = seg_split ? 0 : split, rsub = seg_split ? 0 : sub;
  const int rs_log2 = SGI_TILE_LOG2 - level, rs = 1 << rs_log2;  // this CTA's region of the tile: rs x rs pixels at (qx0,qy0)
  const int qx0 = (rsub & ((1 << level) - 1)) << rs_log2, qy0 = (rsub >> level) << rs_log2;
  const int tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  const int ox = tx << SGI_TILE_LOG2, oy = ty << SGI_TILE_LOG2;

  const int32_t* __restrict__ list = a.pairs + (size_t)tile * a.cap;
  const int tn = item2.y;
  const int nlisted = tn & 0x3FFFFFFF, nbig = a.counters[3];
  const int nspill = (tn & 0x40000000) ? a.counters[5] : 0;    // the list was full: this tile's further entries are somewhere in the spill list
  const int nall = nlisted + nbig + nspill;                    // this tile's list, then the un-binned big triangles, then the spill list (all tiles')
  int it_lo = 0, nitems = nall;                                // the CTA's share of that index space: [it_lo, nitems)
  if (seg_split) { const int per = (nall + (1 << (2 * split)) - 1) >> (2 * split); it_lo = min(nall, sub * per); nitems = min(nall, it_lo + per); }
  const bool empty = nitems == it_lo;                            // nothing can touch the tile: the flush writes the clear values
  if (MODE == SGI_MODE_DEPTH && NT >= 256 && level == 0 && !empty && nitems <= a.direct_max && !a.mm_min) {
    tile_direct_depth<NT>(a, tq, ox, oy, list, nlisted, nitems);        // (nitems <= NT; a spilled list is never short)
    return;
  }

  if (!empty) {
    if (MODE == SGI_MODE_DEPTH) {                              // 4 cells per store
      const int rq_log2 = rs_log2 - 2;
      for (int q = tid; q < (rs * rs) >> 2; q += NT) {
        const int lx = qx0 + ((q & ((1 << rq_log2) - 1)) << 2), ly = qy0 + (q >> rq_log2);
        *reinterpret_cast<uint4*>(&zt[ly * SGI_PITCH + lx]) = make_uint4(ONE_BITS, ONE_BITS, ONE_BITS, ONE_BITS);
      }
    } else if SGI_KEYED(MODE) {   // 2 cells per store
      const int rq_log2 = rs_log2 - 1;
      const unsigned long long clr = ((unsigned long long)ONE_BITS << 32) | 0xFFFFFFFFull;
      for (int q = tid; q < (rs * rs) >> 1; q += NT) {
        const int lx = qx0 + ((q & ((1 << rq_log2) - 1)) << 1), ly = qy0 + (q >> rq_log2);
        *reinterpret_cast<ulonglong2*>(&kt[ly * SGI_PITCH + lx]) = make_ulonglong2(clr, clr);
      }
    } else {
      for (int q = tid; q < rs * rs; q += NT) {
        const int lx = qx0 + (q & (rs - 1)), ly = qy0 + (q >> rs_log2);
        const int p = ly * SGI_PITCH + lx;
        ct[p] = 0;
        const int x = ox + lx, y = oy + ly;
        sd[p] = (x < a.W && y < a.H) ? a.scene_depth[(size_t)y * a.W + x] : 0.0f;
      }
    }
  }
  if (MODE != SGI_MODE_SVCOUNT) { if (tid < SGI_NBLK) bz[tid] = ONE_BITS; }
  else {
    // shadow volumes test against a fixed scene depth: the block bound is its largest scene depth (LEQUAL / LESS both
    // fail for fragments above it)
    __syncthreads();
    if (tid < SGI_NBLK) {
      const int bx = tid % (SGI_TILE / SGI_BLK_W), by = tid / (SGI_TILE / SGI_BLK_W);
      const bool mine = bx * SGI_BLK_W >= qx0 && bx * SGI_BLK_W < qx0 + rs && by * SGI_BLK_H >= qy0 && by * SGI_BLK_H < qy0 + rs;
      float m = 0.0f;
      if (mine)
        for (int j = 0; j < SGI_BLK_H; j++)
          for (int i = 0; i < SGI_BLK_W; i++) m = fmaxf(m, sd[(by * SGI_BLK_H + j) * SGI_PITCH + bx * SGI_BLK_W + i]);
      // (depth-fail counting tallies the fragments BEHIND the scene: nothing may be culled on that side)
      bz[tid] = a.sv_zfail ? ONE_BITS : __float_as_uint(fminf(fmaxf(m, 0.0f), 1.0f));
    }
  }
  int nfrag = 0;
  const TileSink<MODE> sink = {zt, kt, ct, sd, a.depth_func, a.sv_zfail, a.sv_caps, &nfrag};
  // the nearest-first order only pays where hierarchical depth has something to cull: lists of a few dozen triangles skip it
  const bool sort_items = nitems - it_lo > 48;
  // scheduling of the work-item loop, 2 = by list length (measured, profiles/r2_experiments.txt #6): long lists (the teapot's 270 per
  // tile) want the shared cursor and a bound refresh after every block, short ones (a city's few dozen) static dealing and a refresh
  // of fully covered blocks only; stencil counting has no bounds to refresh and prefers static dealing
  const bool static_items = a.static_items == 2 ? (MODE == SGI_MODE_SVCOUNT || !sort_items) : a.static_items != 0;
  const bool refresh_full_only = a.refresh_full_only == 2 ? !sort_items : a.refresh_full_only != 0;

  for (int base = it_lo; base < nitems; base += NT) {
    if (tid == 0) { next_item = 0; q_count = 0; g_count = 0; fc_count = 0; zq_min = 0xFFFFFFFFu; zq_max = 0u; }
    if (tid < SGI_ZBUCKETS) bucket_cnt[tid] = 0;
    int my_k = -1, my_ng = 0;
    unsigned int my_zlo = 0u;
    __syncthreads();                                           // payload initialised / previous chunk drained
    int my_slot = -1;
    if (base + tid < nitems) {
      const int it = base + tid;
      if (it < nlisted) my_slot = __ldg(&list[it]);
      else if (it < nlisted + nbig) my_slot = __ldg(&a.big_list[it - nlisted]);
      else { const int2 e = __ldg(&a.spill[it - nlisted - nbig]); if (e.x == tile) my_slot = e.y; }
    }
    if (my_slot >= 0) {
      const SgiRec* rp = &a.rec[my_slot];
      const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(rp) + 3);
      const int px0 = (int)(short)(q3.x & 0xFFFF), py0 = (int)(short)(q3.x >> 16);
      const int px1 = (int)(short)(q3.y & 0xFFFF), py1 = (int)(short)(q3.y >> 16);
      const int lx0 = max(px0 - ox, qx0), ly0 = max(py0 - oy, qy0);
      const int lx1 = min(px1 - ox, qx0 + rs - 1), ly1 = min(py1 - oy, qy0 + rs - 1);
      const int w = lx1 - lx0 + 1, h = ly1 - ly0 + 1;
      // whole-tile items fetch the full record at once (nearly every listed triangle touches the tile); sub-tile items
      // look at the bounding box first, most of the tile's list misses their region
      uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0, q2 = q0;
      if (level == 0 || (w > 0 && h > 0)) {
        q0 = __ldg(reinterpret_cast<const uint4*>(rp));
        q1 = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
        q2 = __ldg(reinterpret_cast<const uint4*>(rp) + 2);
      }
      if (w > 0 && h > 0) {
        const int X0 = (int)q0.x, Y0 = (int)q0.y, X1 = (int)q0.z, Y1 = (int)q0.w, X2 = (int)q1.x, Y2 = (int)q1.y;
        if (w * h <= SGI_SMALL_TRI) {
          const float z0 = __uint_as_float(q1.z), dz1 = __uint_as_float(q1.w), dz2 = __uint_as_float(q2.x);
          const float ia = __uint_as_float(q2.y), zoff = __uint_as_float(q2.z);
          EdgeSet es;
          es.init(X0, Y0, X1, Y1, X2, Y2);
          for (int ly = ly0; ly <= ly1; ly++)
            for (int lx = lx0; lx <= lx1; lx++) {
              long long E1, E2;
              if (!es.test(ox + lx, oy + ly, E1, E2)) continue;
              sink.fragment(lx, ly, frag_z(z0, dz1, dz2, ia, zoff, E1, E2), (int)q2.w);
            }
        } else {
          const int k = atomicAdd(&q_count, 1);
          tq.X0[k] = X0; tq.Y0[k] = Y0; tq.X1[k] = X1; tq.Y1[k] = Y1; tq.X2[k] = X2; tq.Y2[k] = Y2;
          tq.z0[k] = __uint_as_float(q1.z); tq.dz1[k] = __uint_as_float(q1.w); tq.dz2[k] = __uint_as_float(q2.x);
          tq.ia[k] = __uint_as_float(q2.y); tq.zoff[k] = __uint_as_float(q2.z);
          tq.meta[k] = (int)q2.w;
          my_zlo = tri_depth_lower_bound(__uint_as_float(q1.z), __uint_as_float(q1.w), __uint_as_float(q2.x), __uint_as_float(q2.z));
          tq.zlo[k] = my_zlo;
          tq.box[k] = lx0 | (ly0 << 8) | (lx1 << 16) | (ly1 << 24);
          // one work item per group of 32 blocks, so a triangle covering the tile is shared by 4 warps
          const int nb = (lx1 / SGI_BLK_W - lx0 / SGI_BLK_W + 1) * (ly1 / SGI_BLK_H - ly0 / SGI_BLK_H + 1);
          my_k = k; my_ng = (nb + 31) >> 5;
          if (w == rs && h == rs) {
            // the bounding box spans the whole region: if every edge function is inside at the region corner that
            // minimises it (edge functions are affine), all rs x rs pixels are covered and the triangle is drawn by the
            // whole CTA in one coalesced sweep, without block tests and per-pixel edge tests (floors, walls, SV prisms)
            const int x0c = (ox + qx0) * SGI_SUBPIX + SGI_SUBPIX / 2, x1c = x0c + (rs - 1) * SGI_SUBPIX;
            const int y0c = (oy + qy0) * SGI_SUBPIX + SGI_SUBPIX / 2, y1c = y0c + (rs - 1) * SGI_SUBPIX;
            const int XA[3] = {X1, X2, X0}, YA[3] = {Y1, Y2, Y0}, XB[3] = {X2, X0, X1}, YB[3] = {Y2, Y0, Y1};
            bool full = true;
#pragma unroll
            for (int e = 0; e < 3; e++) {
              const int dx = XB[e] - XA[e], dy = YB[e] - YA[e];
              const int cx = (dy < 0) ? x0c : x1c, cy = (dx > 0) ? y0c : y1c;       // corner minimising this edge
              const long long bias = (dy < 0 || (dy == 0 && dx < 0)) ? 0 : 1;
              const long long v = (long long)dx * (long long)(cy - YA[e]) - ((long long)dy * (long long)(cx - XA[e]) + bias);
              full = full && v >= 0;
            }
            if (full) {
              const int f = atomicAdd(&fc_count, 1);
              if (f < SGI_MAX_FULL) { fc_list[f] = k; my_ng = 0; }
            }
          }
          if (sort_items) { atomicMin(&zq_min, my_zlo); atomicMax(&zq_max, my_zlo); }
          else if (my_ng > 0) {          // short list: work items in arrival order, no sort phases (three barriers less per chunk)
            const int g0 = atomicAdd(&g_count, my_ng);
            for (int g = 0; g < my_ng; g++) tq.group[g0 + g] = (k << 3) | g;
          }
        }
      }
    }
    __syncthreads();
    // nearest-first issue order: counting sort of the work items on their depth bound, so that the block bounds
    // tighten early and the triangles behind them are culled
    int my_b = 0;
    if (sort_items) {
    if (my_k >= 0) {
      const unsigned int lo = zq_min, span = zq_max - lo + 1u;
      my_b = (int)(((unsigned long long)(my_zlo - lo) * SGI_ZBUCKETS) / span);
      atomicAdd(&bucket_cnt[my_b], my_ng);
    }
    __syncthreads();
    if (tid < 32) {                                            // exclusive prefix over the 64 buckets
      int c0 = bucket_cnt[2 * tid], c1 = bucket_cnt[2 * tid + 1], incl = c0 + c1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, d); if (tid >= d) incl += v; }
      bucket_pos[2 * tid] = incl - c0 - c1; bucket_pos[2 * tid + 1] = incl - c1;
      if (tid == 31) g_count = incl;
    }
    __syncthreads();
    if (my_k >= 0) {
      const int g0 = atomicAdd(&bucket_pos[my_b], my_ng);
      for (int g = 0; g < my_ng; g++) tq.group[g0 + g] = (my_k << 3) | g;
    }
    __syncthreads();
    }
    const int ngroups = g_count;
    {  // region-covering triangles first: every thread takes its pixels of the region
      const int nfull = min(fc_count, SGI_MAX_FULL);
      for (int f = 0; f < nfull; f++) {
        const int qi = fc_list[f];
        const int X0 = tq.X0[qi], Y0 = tq.Y0[qi], X1 = tq.X1[qi], Y1 = tq.Y1[qi], X2 = tq.X2[qi], Y2 = tq.Y2[qi];
        const float z0 = tq.z0[qi], dz1 = tq.dz1[qi], dz2 = tq.dz2[qi], ia = tq.ia[qi], zoff = tq.zoff[qi];
        const int meta = tq.meta[qi];
        const int dx1 = X0 - X2, dy1 = Y0 - Y2, dx2 = X1 - X0, dy2 = Y1 - Y0;
        // The edge values the depth needs are affine in the pixel position with integer coefficients: one pixel to the right
        // is -256 dy, one row up +256 dx.  Where every value over the region stays below 2^52 in magnitude (affine: checked
        // with the value at the region origin plus the largest possible excursion) they are exact in double precision, so a
        // thread steps from row to row with one double add per edge instead of two wide multiplies, and (float) of the exact
        // double rounds once, to the same float as (float) of the 64-bit integer: identical depths, a quarter of the instructions.
        const int PXo = (ox + qx0) * SGI_SUBPIX + SGI_SUBPIX / 2, PYo = (oy + qy0) * SGI_SUBPIX + SGI_SUBPIX / 2;
        const long long E1o = (long long)dx1 * (long long)(PYo - Y2) - (long long)dy1 * (long long)(PXo - X2);
        const long long E2o = (long long)dx2 * (long long)(PYo - Y0) - (long long)dy2 * (long long)(PXo - X0);
        const long long span = (long long)(SGI_TILE - 1) * SGI_SUBPIX;
        const long long m1 = llabs(E1o) + span * (llabs((long long)dx1) + llabs((long long)dy1));
        const long long m2 = llabs(E2o) + span * (llabs((long long)dx2) + llabs((long long)dy2));
        if (m1 < (1LL << 52) && m2 < (1LL << 52)) {
          // NT is a multiple of the region width: a thread keeps its column and moves up NT / rs rows per step
          const int lxr = tid & (rs - 1), lyr0 = tid >> rs_log2, ystep = NT >> rs_log2;
          double e1 = (double)(E1o - (long long)lxr * (256LL * dy1) + (long long)lyr0 * (256LL * dx1));
          double e2 = (double)(E2o - (long long)lxr * (256LL * dy2) + (long long)lyr0 * (256LL * dx2));
          const double s1 = (double)((long long)ystep * (256LL * dx1)), s2 = (double)((long long)ystep * (256LL * dx2));
          const int lx = qx0 + lxr;
          for (int lyr = lyr0; lyr < rs; lyr += ystep) {
            const float b1 = (float)e1 * ia, b2 = (float)e2 * ia;
            float z = (z0 + b1 * dz1) + b2 * dz2;
            z = z + zoff;
            if (!(z >= 0.0f)) z = 0.0f;
            if (z > 1.0f) z = 1.0f;
            sink.fragment(lx, qy0 + lyr, z, meta);
            e1 += s1; e2 += s2;
          }
        } else {
          for (int q = tid; q < rs * rs; q += NT) {
            const int lx = qx0 + (q & (rs - 1)), ly = qy0 + (q >> rs_log2);
            const int PX = (ox + lx) * SGI_SUBPIX + SGI_SUBPIX / 2, PY = (oy + ly) * SGI_SUBPIX + SGI_SUBPIX / 2;
            const long long E1 = (long long)dx1 * (long long)(PY - Y2) - (long long)dy1 * (long long)(PX - X2);
            const long long E2 = (long long)dx2 * (long long)(PY - Y0) - (long long)dy2 * (long long)(PX - X0);
            sink.fragment(lx, ly, frag_z(z0, dz1, dz2, ia, zoff, E1, E2), meta);
          }
        }
      }
    }
    // the per-block depth bounds only pay off when there is something to cull: refresh them only in busy chunks
    const bool refresh_bounds = ngroups >= 8;
    // work items are dealt to the warps round robin (they are in nearest-first order: every warp starts near the front); a shared
    // cursor balanced them slightly better but cost an atomic and a shuffle per item, 6 % of the depth kernel's instructions
    for (int item_s = tid >> 5;; item_s += NT / 32) {
      int item = item_s;
      if (!static_items) {                                     // shared cursor: balances warps over items of very different cost
        if (lane == 0) item = atomicAdd(&next_item, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
      }
      if (item >= ngroups) break;
      const int grp = tq.group[item];
      const int qi = grp >> 3, b0 = (grp & 7) << 5;
      const int X0 = tq.X0[qi], Y0 = tq.Y0[qi], X1 = tq.X1[qi], Y1 = tq.Y1[qi], X2 = tq.X2[qi], Y2 = tq.Y2[qi];
      const int box = tq.box[qi];
      const int lx0 = box & 0xFF, ly0 = (box >> 8) & 0xFF, lx1 = (box >> 16) & 0xFF, ly1 = (box >> 24) & 0xFF;
      const int bx0 = lx0 / SGI_BLK_W, bx1 = lx1 / SGI_BLK_W, by0 = ly0 / SGI_BLK_H, by1 = ly1 / SGI_BLK_H;
      const int nbx = bx1 - bx0 + 1, nb = nbx * (by1 - by0 + 1);
      const int sub_x = lane & (SGI_BLK_W - 1), sub_y = lane >> 3;
      const int b = b0 + lane;
      // b / nbx and b % nbx for b < 256, nbx <= 8 without a division: (b * ceil(2^16 / nbx)) >> 16 is exact in that range
      const int bq = (b * c_rcp16[nbx]) >> 16;
      const int bx = bx0 + (b - bq * nbx), by = by0 + bq;
      const unsigned int zlo = tq.zlo[qi];
      bool keep = b < nb;
      const unsigned int bound = keep ? bz[by * (SGI_TILE / SGI_BLK_W) + bx] : 0u;
      if (keep) keep = zlo <= bound;                           // hierarchical depth, triangle-wide bound
      if (keep && nb > 1) {
        const int gx = ox + bx * SGI_BLK_W, gy = oy + by * SGI_BLK_H;
        keep = edge_block_max(X1, Y1, X2, Y2, gx, gy) >= 0 && edge_block_max(X2, Y2, X0, Y0, gx, gy) >= 0 &&
               edge_block_max(X0, Y0, X1, Y1, gx, gy) >= 0;
        if (keep && bound < ONE_BITS && MODE == SGI_MODE_SVCOUNT) {
          // (only for shadow volumes and for triangles spanning many blocks, where the triangle-wide bound is loose;
          //  for small ones the extra test costs more than it culls: measured, profiles/r1_hiz.txt)
          // hierarchical depth, per-block bound: the interpolated depth is affine in the pixel position, so over the
          // block it is smallest at one of the four corner pixels; evaluate those with the fragment formula and lower
          // the result by a slack covering fp32 rounding (terms t = b*dz can be large at corners outside the triangle)
          const float tz0 = tq.z0[qi], tdz1 = tq.dz1[qi], tdz2 = tq.dz2[qi], tia = tq.ia[qi], tzoff = tq.zoff[qi];
          float zmin_blk = 2.0f;
#pragma unroll
          for (int cnr = 0; cnr < 4; cnr++) {
            const int PX = (gx + ((cnr & 1) ? SGI_BLK_W - 1 : 0)) * SGI_SUBPIX + SGI_SUBPIX / 2;
            const int PY = (gy + ((cnr & 2) ? SGI_BLK_H - 1 : 0)) * SGI_SUBPIX + SGI_SUBPIX / 2;
            const long long E1 = (long long)(X0 - X2) * (long long)(PY - Y2) - (long long)(Y0 - Y2) * (long long)(PX - X2);
            const long long E2 = (long long)(X1 - X0) * (long long)(PY - Y0) - (long long)(Y1 - Y0) * (long long)(PX - X0);
            const float t1 = ((float)E1 * tia) * tdz1, t2 = ((float)E2 * tia) * tdz2;
            const float zc = ((tz0 + t1) + t2) + tzoff - (1.0e-6f + 5.0e-7f * (fabsf(t1) + fabsf(t2) + fabsf(tzoff)));
            zmin_blk = fminf(zmin_blk, zc);
          }
          if (zmin_blk > 1.0f) zmin_blk = 1.0f;
          keep = !(zmin_blk > 0.0f) || __float_as_uint(zmin_blk) <= bound;     // NaN / negative bounds never cull
        }
      }
      unsigned int mask = __ballot_sync(0xffffffffu, keep);
      if (!mask) continue;
      const float z0 = tq.z0[qi], dz1 = tq.dz1[qi], dz2 = tq.dz2[qi], ia = tq.ia[qi], zoff = tq.zoff[qi];
      const int meta = tq.meta[qi];
      EdgeSet es;
      es.init(X0, Y0, X1, Y1, X2, Y2);
      // No bounding-box test per pixel: a pixel outside the box cannot pass the edge tests, and pixels beyond the
      // viewport (the box is clamped to it) land in tile cells that are never flushed.
      //
      // The edge functions are affine in the pixel position with integer coefficients: the lane's (biased) values at its pixel
      // of the tile's first block are formed once per work item; a block then adds (block column) x (8 px step) and (block row) x
      // (4 px step), which is two wide multiply-adds per edge with the 64-bit lane value as the addend - six IMAD.WIDE and one
      // sign test per pixel instead of six wide multiplies, their operand differences and three 64-bit subtractions.
      // Exactly the integers EdgeSet::test produces.
      const int PXl = (ox + sub_x) * SGI_SUBPIX + SGI_SUBPIX / 2, PYl = (oy + sub_y) * SGI_SUBPIX + SGI_SUBPIX / 2;
      const long long l0 = (long long)es.dx0 * (long long)(PYl - Y1) - ((long long)es.dy0 * (long long)(PXl - X1) + es.b0);
      const long long l1 = (long long)es.dx1 * (long long)(PYl - Y2) - ((long long)es.dy1 * (long long)(PXl - X2) + es.b1);
      const long long l2 = (long long)es.dx2 * (long long)(PYl - Y0) - ((long long)es.dy2 * (long long)(PXl - X0) + es.b2);
      while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const int kbx = __shfl_sync(0xffffffffu, bx, k), kby = __shfl_sync(0xffffffffu, by, k);
        const int lx = kbx * SGI_BLK_W + sub_x, ly = kby * SGI_BLK_H + sub_y;
        const int cx = -(kbx * (SGI_BLK_W * SGI_SUBPIX)), cy = kby * (SGI_BLK_H * SGI_SUBPIX);
        const long long e0 = madw(es.dy0, cx, madw(es.dx0, cy, l0));
        const long long e1 = madw(es.dy1, cx, madw(es.dx1, cy, l1));
        const long long e2 = madw(es.dy2, cx, madw(es.dx2, cy, l2));
        const bool covered = (e0 | e1 | e2) >= 0;
        if (covered) sink.fragment(lx, ly, frag_z(z0, dz1, dz2, ia, zoff, e1 + es.b1, e2 + es.b2), meta);
        // (a block the triangle covers only in part keeps texels at their old depth: its bound would not move)
        if (MODE != SGI_MODE_SVCOUNT && refresh_bounds && (!refresh_full_only || __all_sync(0xffffffffu, covered))) {
          // refresh the block's bound from what is stored now (one warp-wide max; other warps can only lower it further)
          const int p = ly * SGI_PITCH + lx;
          const unsigned int cur = (MODE == SGI_MODE_DEPTH) ? zt[p] : (unsigned int)(kt[p] >> 32);
          const unsigned int wmax = __reduce_max_sync(0xffffffffu, cur);
          const int bi = kby * (SGI_TILE / SGI_BLK_W) + kbx;
          if (lane == 0 && wmax < bz[bi]) atomicMin(&bz[bi], wmax);
        }
      }
    }
    if (base + NT < nitems) __syncthreads();                   // every warp is done with this chunk's queue (the last chunk runs into the flush barrier)
  }
  if (MODE == SGI_MODE_DEPTH) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile was written through the generic proxy
  __syncthreads();

  // ---- write the tile to HBM exactly once ------------------------------------------------------------------
  if (MODE == SGI_MODE_DEPTH && a.mm_min) {
    // Extrema of the finished depths per 32x32-texel block of the map (the shadow pass culls whole tap windows with them: a
    // pixel nearer than the smallest depth around its footprint has no blocker, one beyond the largest has no lit tap).
    // A warp's 32 lanes read one row segment inside ONE block: two warp reductions, then one global atomic pair per warp and row
    // group.  Texels beyond the map edge still hold the clear value 1.0 (they only make the maximum more conservative).
    const int bxg = (ox + qx0) >> 5, byg = (oy + qy0) >> 5;
    if (empty) {
      const int nb = (rs + 31) >> 5;
      if (tid < nb * nb) {
        const int o = (byg + tid / nb) * a.mm_w + bxg + tid % nb;
        atomicMax(&a.mm_max[o], ONE_BITS);                    // (the minimum is already 1.0)
      }
    } else {
      unsigned int mn = 0xFFFFFFFFu, mx = 0u;
      int cur = -1;
      for (int q = tid; q < rs * rs; q += NT) {
        const int lxr = q & (rs - 1), lyr = q >> rs_log2;
        const int blk = (lyr >> 5) * 2 + (lxr >> 5);          // block of the region this texel belongs to (warp-uniform)
        if (blk != cur && cur >= 0) {
          mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
          if (lane == 0) { const int o = (byg + (cur >> 1)) * a.mm_w + bxg + (cur & 1); atomicMin(&a.mm_min[o], mn); atomicMax(&a.mm_max[o], mx); }
          mn = 0xFFFFFFFFu; mx = 0u;
        }
        cur = blk;
        const unsigned int v = zt[(qy0 + lyr) * SGI_PITCH + qx0 + lxr];
        mn = min(mn, v); mx = max(mx, v);
      }
      if (cur >= 0) {
        mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) { const int o = (byg + (cur >> 1)) * a.mm_w + bxg + (cur & 1); atomicMin(&a.mm_min[o], mn); atomicMax(&a.mm_max[o], mx); }
      }
    }
  }
  if (MODE == SGI_MODE_DEPTH) {
    const int gx0 = ox + qx0, gy0 = oy + qy0;
    const int wv = min(rs, a.W - gx0);                          // columns of the region inside the map
    // Whole rows of the region inside the job rectangle, 16-byte aligned: one bulk copy per row, shared -> global, issued by
    // one thread per row (cp.async.bulk, the TMA engine's linear form: the row pitch of 72 words that keeps the raster free of
    // bank conflicts rules out a 2-D tensor box).  The SM's load/store path carries no flush traffic and the other threads are done.
    const bool bulk = a.bulk_flush && !empty && (a.W & 3) == 0 && (wv & 3) == 0 && wv > 0 && gx0 >= a.rx0 && gx0 + wv <= a.rx1;
    if (bulk) {
      if (tid < rs) {
        const int y = gy0 + tid;
        if (y >= a.ry0 && y < a.ry1) {
          const unsigned src = (unsigned)__cvta_generic_to_shared(&zt[(qy0 + tid) * SGI_PITCH + qx0]);
          float* dst = a.depth + (size_t)y * a.W + gx0;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(wv * 4) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // shared memory must outlive the copy's reads
        }
      }
      return;
    }
    // four texels per thread and store: 16-byte stores when the row pitch allows it
    const int rq_log2 = rs_log2 - 2;
    const bool vec_ok = (a.W & 3) == 0;
    for (int q = tid; q < (rs * rs) >> 2; q += NT) {
      const int lx = qx0 + ((q & ((1 << rq_log2) - 1)) << 2), ly = qy0 + (q >> rq_log2);
      const int x = ox + lx, y = oy + ly;
      if (y < a.ry0 || y >= a.ry1 || x + 3 < a.rx0 || x >= a.rx1) continue;
      const uint4 v = empty ? make_uint4(ONE_BITS, ONE_BITS, ONE_BITS, ONE_BITS)
                            : *reinterpret_cast<const uint4*>(&zt[ly * SGI_PITCH + lx]);
      float* dst = a.depth + (size_t)y * a.W + x;
      if (vec_ok && x >= a.rx0 && x + 3 < a.rx1) {
        *reinterpret_cast<uint4*>(dst) = v;
      } else {
        const unsigned int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (x + i >= a.rx0 && x + i < a.rx1) dst[i] = __uint_as_float(vv[i]);
      }
    }
  }
  if (MODE == SGI_MODE_SVCOUNT && a.frag_counter && nfrag) atomicAdd(a.frag_counter, (unsigned long long)nfrag);
  for (int q = tid; MODE != SGI_MODE_DEPTH && q < rs * rs; q += NT) {
    const int lx = qx0 + (q & (rs - 1)), ly = qy0 + (q >> rs_log2);
    const int p = ly * SGI_PITCH + lx;
    const int x = ox + lx, y = oy + ly;
    if (x < a.rx0 || x >= a.rx1 || y < a.ry0 || y >= a.ry1) continue;
    const size_t o = (size_t)y * a.W + x;
    if (MODE == SGI_MODE_SVCOUNT) {
      const int c = empty ? 0 : ct[p];
      if (a.sv_split_lists) { if (c) atomicAdd(&a.count[o], c); }       // (target zeroed before the pass; k_sv_stencil derives the stencil)
      else { a.count[o] = c; a.stencil[o] = (uint8_t)((unsigned int)c & 255u); }
    } else {
      const unsigned long long key = empty ? 0xFFFFFFFFull : kt[p];
      const unsigned int lo32 = (unsigned int)(key & 0xFFFFFFFFull);
      if (MODE == SGI_MODE_IDS) { a.ids[o] = lo32; continue; }     // visibility only: the winning primitive (0xFFFFFFFF = background)
      if (MODE == SGI_MODE_MOMENTS) {
        // Moments.frag / Exponential.frag / ExponentialMoments.frag on the winning fragment: a function of the un-offset depth
        // plane at this texel and at its two quad partners (dFdx / dFdy), fused into the flush - the moment target crosses
        // HBM once and no depth map is written at all
        if (lo32 == 0xFFFFFFFFu) { a.mom4[o] = make_float4(0.f, 0.f, 0.f, 1.f); continue; }      // glClearColor(0,0,0,1), main.cpp:356
        const int prim = (int)lo32, t = prim >> 3, sub = prim & 7;
        const int slot = (sub == 0) ? t : __ldg(&a.ovf_base[t]) + sub - 1;
        SgiRec r;
        {
          const uint4* rq = reinterpret_cast<const uint4*>(&a.rec[slot]);
          uint4* rd = reinterpret_cast<uint4*>(&r);
          rd[0] = __ldg(rq); rd[1] = __ldg(rq + 1); rd[2] = __ldg(rq + 2); rd[3] = __ldg(rq + 3);
        }
        const int dx1 = r.X0 - r.X2, dy1 = r.Y0 - r.Y2, dx2 = r.X1 - r.X0, dy2 = r.Y1 - r.Y0;
        auto plane = [&](int px, int py) -> float {
          const int PX = px * SGI_SUBPIX + SGI_SUBPIX / 2, PY = py * SGI_SUBPIX + SGI_SUBPIX / 2;
          const long long E1 = (long long)dx1 * (long long)(PY - r.Y2) - (long long)dy1 * (long long)(PX - r.X2);
          const long long E2 = (long long)dx2 * (long long)(PY - r.Y0) - (long long)dy2 * (long long)(PX - r.X0);
          const float b1 = (float)E1 * r.ia, b2 = (float)E2 * r.ia;
          return (r.z0 + b1 * r.dz1) + b2 * r.dz2;
        };
        a.mom4[o] = mom_texel(a.mom_tech, plane(x, y), plane(x ^ 1, y), plane(x, y ^ 1), x & 1, y & 1, a.z_near, a.z_far, a.mq, a.mqt);
        continue;
      }
      if (lo32 == 0xFFFFFFFFu) {
        a.depth[o] = 1.0f;
        a.pos4[o] = make_float4(0.f, 0.f, 0.f, 1.f);
        a.nrm4[o] = make_float4(0.f, 0.f, 0.f, 1.f);
        if (MODE == SGI_MODE_GBUFFER_RGB) a.albedo4[o] = make_float4(0.f, 0.f, 0.f, 1.f);
        continue;
      }
      const int prim = (int)lo32, t = prim >> 3, sub = prim & 7;
      const int slot = (sub == 0) ? t : __ldg(&a.ovf_base[t]) + sub - 1;
      // read-only path for every gather of the resolve: lets the loads of a pixel issue together instead of in
      // program order behind the G-buffer stores
      SgiRec r; SgiRecAttr at;
      {
        const uint4* rq = reinterpret_cast<const uint4*>(&a.rec[slot]);
        const uint4* aq = reinterpret_cast<const uint4*>(&a.attr[slot]);
        uint4* rd = reinterpret_cast<uint4*>(&r); uint4* ad = reinterpret_cast<uint4*>(&at);
        rd[0] = __ldg(rq); rd[1] = __ldg(rq + 1); rd[2] = __ldg(rq + 2); rd[3] = __ldg(rq + 3);
        ad[0] = __ldg(aq); ad[1] = __ldg(aq + 1); ad[2] = __ldg(aq + 2); ad[3] = __ldg(aq + 3); ad[4] = __ldg(aq + 4); ad[5] = __ldg(aq + 5);
        if (MODE == SGI_MODE_GBUFFER_RGB) { ad[6] = __ldg(aq + 6); ad[7] = __ldg(aq + 7); }
      }
      long long E0, E1, E2;
      cover(r.X0, r.Y0, r.X1, r.Y1, r.X2, r.Y2, x, y, E0, E1, E2);
      const float q0 = ((float)E0 * r.ia) * at.iw[0];
      const float q1 = ((float)E1 * r.ia) * at.iw[1];
      const float q2 = ((float)E2 * r.ia) * at.iw[2];
      const float iq = 1.0f / ((q0 + q1) + q2);
      float outv[6];
#pragma unroll
      for (int c = 0; c < 6; c++) outv[c] = ((q0 * at.A[0][c] + q1 * at.A[1][c]) + q2 * at.A[2][c]) * iq;
      a.depth[o] = __uint_as_float((unsigned int)(key >> 32));
      a.pos4[o] = make_float4(outv[0], outv[1], outv[2], 1.0f);
      a.nrm4[o] = make_float4(outv[3], outv[4], outv[5], (r.prim_front & 1) ? 1.0f : 0.0f);
      if (MODE == SGI_MODE_GBUFFER_RGB) {                // third target of GBuffer.frag: the interpolated vertex colour (own instantiation:
                                                         // carrying this code in the colour-less kernel cost 8 % of its time)
        float col[3];
#pragma unroll
        for (int cc = 0; cc < 3; cc++) col[cc] = ((q0 * at.C[0][cc] + q1 * at.C[1][cc]) + q2 * at.C[2][cc]) * iq;
        if (a.uvrec) {                                   // useTextureForColoring: select on the interpolated (u, v, texture id)
          SgiRecUV ur;
          const uint4* uq = reinterpret_cast<const uint4*>(&a.uvrec[slot]);
          uint4* ud = reinterpret_cast<uint4*>(&ur);
          ud[0] = __ldg(uq); ud[1] = __ldg(uq + 1); ud[2] = __ldg(uq + 2);
          float uvw[3];
#pragma unroll
          for (int cc = 0; cc < 3; cc++) uvw[cc] = ((q0 * ur.U[0][cc] + q1 * ur.U[1][cc]) + q2 * ur.U[2][cc]) * iq;
          a.albedo4[o] = fragment_color(a.tex, uvw[0], uvw[1], uvw[2], col[0], col[1], col[2]);
        } else
        a.albedo4[o] = make_float4(col[0], col[1], col[2], 1.0f);
      }
    }
  }
}

template <int MODE, int NT>
constexpr size_t tile_smem_bytes() {
  return (size_t)SGI_TILE * SGI_PITCH * (SGI_KEYED(MODE) ? 8 : 4) + (MODE == SGI_MODE_SVCOUNT ? (size_t)SGI_TILE * SGI_PITCH * 4 : 0) +
         sizeof(TriQueue<NT>);
}

// ---- shadow-volume extrusion: ShadowVolumes/src/ShadowVolume.cpp:116-195 -----------------------------------------
// `per` = 6: the three side quads of the reference's open prism; 8: + near cap (the triangle itself, first vertex first so that
// its raster record - and with it every fragment depth - is the scene triangle's own) + far cap, closing the volume for
// depth-fail counting.  cls[t] = the orientation class (dot(average normal, light position) >= 0) the silhouette pass pairs on.
__global__ void __launch_bounds__(128) k_sv_extrude(const float* __restrict__ xyz, const float* __restrict__ nrm,
                                                    const int32_t* __restrict__ idx, int T, float lx, float ly, float lz,
                                                    int infinity, int per, float* __restrict__ pxyz, int32_t* __restrict__ pidx,
                                                    unsigned char* __restrict__ cls) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float L[3] = {lx, ly, lz};
  int v[3] = {idx[3 * t], idx[3 * t + 1], idx[3 * t + 2]};
  float* q = pxyz + (size_t)t * 18;
  float n[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float p = xyz[3 * (size_t)v[k] + a];
      q[k * 3 + a] = p;
      q[(3 + k) * 3 + a] = (p - L[a]) * (float)infinity;           // :39-41 / :140-142
    }
    s = nrm[3 * (size_t)v[0] + a] + nrm[3 * (size_t)v[1] + a] + nrm[3 * (size_t)v[2] + a];
    n[a] = s / 3.0f;
  }
  float d = n[0] * L[0] + n[1] * L[1] + n[2] * L[2];
  // index order flips on dot(avg normal, light POSITION) >= 0  (:61 / :162)
  const int ordA[24] = {1, 0, 3, 1, 3, 4, 2, 1, 4, 2, 4, 5, 0, 2, 5, 0, 5, 3, 0, 1, 2, 4, 3, 5};
  const int ordB[24] = {4, 3, 0, 4, 0, 1, 5, 4, 1, 5, 1, 2, 3, 5, 2, 3, 2, 0, 0, 2, 1, 3, 4, 5};
  const bool A = d >= 0.0f;
  if (cls) cls[t] = A ? 1 : 0;
#pragma unroll
  for (int k = 0; k < 24; k++)
    if (k < per * 3) pidx[(size_t)t * per * 3 + k] = t * 6 + (A ? ordA[k] : ordB[k]);
}

// Silhouette form: one thread per undirected edge of the mesh (groups prepared once per mesh on the host: the directed edges
// (triangle, local edge) that share a vertex-index pair, backward ones - first index > second - first, each side by triangle).
// Per orientation class, min(#backward, #forward) quads of either direction cancel in pairs, lowest triangles first; the
// index triples of a dropped quad become (0,0,0), a degenerate triangle the rasteriser discards at set-up.
__global__ void __launch_bounds__(128) k_sv_silhouette(const int32_t* __restrict__ grp_start, const int32_t* __restrict__ grp_ent, int G,
                                                       const unsigned char* __restrict__ cls, int per, int32_t* __restrict__ pidx) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int e0 = grp_start[g], e1 = grp_start[g + 1];
  if (e1 - e0 < 2) return;
#pragma unroll
  for (int c = 0; c < 2; c++) {
    int nb = 0, nf = 0;
    for (int e = e0; e < e1; e++) {
      const int ent = grp_ent[e];                        // (3 t + local edge) << 1 | forward
      if (cls[(ent >> 1) / 3] == c) { if (ent & 1) nf++; else nb++; }
    }
    int kb = min(nb, nf), kf = kb;
    if (!kb) continue;
    for (int e = e0; e < e1; e++) {
      const int ent = grp_ent[e];
      const int te = ent >> 1, t = te / 3, le = te - 3 * t;
      if (cls[t] != c) continue;
      int& k = (ent & 1) ? kf : kb;
      if (k > 0) {
        k--;
        int32_t* o = pidx + ((size_t)t * per + 2 * le) * 3;
#pragma unroll
        for (int j = 0; j < 6; j++) o[j] = 0;
      }
    }
  }
}

}  // namespace

// ================================================ host side =====================================================
static int grow(sgi_ctx* ctx, void** p, size_t bytes) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return SGI_ERR_NOMEM; }
  return SGI_OK;
}

static int sgi_raster_reserve(sgi_ctx* ctx, SgiScratch& sc, int max_tris, int W, int H, cudaStream_t stream) {
  int rc;
  if (max_tris > sc.rec_cap_tris) {
    SGI_CUDA(ctx, cudaStreamSynchronize(stream));
    size_t n = (size_t)max_tris * 7 + 16;
    if ((rc = grow(ctx, (void**)&sc.d_rec, n * sizeof(SgiRec)))) return rc;
    if ((rc = grow(ctx, (void**)&sc.d_attr, n * sizeof(SgiRecAttr)))) return rc;
    if ((rc = grow(ctx, (void**)&sc.d_uvrec, n * sizeof(SgiRecUV)))) return rc;
    if ((rc = grow(ctx, (void**)&sc.d_ovf_base, (size_t)max_tris * 4 + 16))) return rc;
    if ((rc = grow(ctx, (void**)&sc.d_big, n * 4))) return rc;
    sc.rec_cap_tris = max_tris;
  }
  if (!sc.h_flags) {
    SGI_CUDA(ctx, cudaHostAlloc((void**)&sc.h_flags, 64, cudaHostAllocMapped));
    for (int k = 0; k < 16; k++) sc.h_flags[k] = 0;
    SGI_CUDA(ctx, cudaMalloc((void**)&sc.d_sticky, 64));
    SGI_CUDA(ctx, cudaMemset(sc.d_sticky, 0, 64));
  }
  int tiles = ((W + SGI_TILE - 1) >> SGI_TILE_LOG2) * ((H + SGI_TILE - 1) >> SGI_TILE_LOG2);
  if (tiles + 1 > sc.tile_cap) {
    int cap = tiles + 1 + 64;
    SGI_CUDA(ctx, cudaStreamSynchronize(stream));
    if ((rc = grow(ctx, (void**)&sc.d_counters, (size_t)(32 + cap) * 4))) return rc;   // counters | snapshot | tile cursors
    sc.d_snap = sc.d_counters + 16;
    sc.d_tile_cnt = sc.d_counters + 32;
    if ((rc = grow(ctx, (void**)&sc.d_tile_order, ((size_t)cap + SGI_SPLIT_EXTRA_SEG) * 8))) return rc;   // work items
    if ((rc = grow(ctx, (void**)&sc.d_tile_zmax, (size_t)cap * 4))) return rc;
    sc.tile_cap = cap;
    sc.needs_clear = true;
  }
  return SGI_OK;
}

// room for `cap` list entries per tile and `spill` (tile, record) pairs
#define SGI_LIST_BUDGET ((size_t)1 << 28)       // entries (1 GiB) in per-tile lists; beyond that the lists are capped and the spill list carries the rest
static int sgi_raster_reserve_lists(sgi_ctx* ctx, SgiScratch& sc, int cap, int n_tiles, int spill, cudaStream_t stream) {
  int rc = SGI_OK;
  const size_t want = (size_t)cap * (size_t)n_tiles + 16;
  if (want > sc.pair_alloc) {
    SGI_CUDA(ctx, cudaStreamSynchronize(stream));
    rc = grow(ctx, (void**)&sc.d_pairs, want * 4);
    sc.pair_alloc = rc ? 0 : want;
    if (rc) return rc;
  }
  if (spill > sc.spill_cap) {
    SGI_CUDA(ctx, cudaStreamSynchronize(stream));
    const int cap2 = spill + spill / 4;
    rc = grow(ctx, (void**)&sc.d_spill, (size_t)cap2 * 8);
    sc.spill_cap = rc ? 0 : cap2;
  }
  return rc;
}
// list capacity per tile for a longest list of `longest` entries: 2x headroom, within the memory budget
static int list_capacity(int longest, int n_tiles) {
  long long cap = (long long)longest * 2 + 64;
  const long long lim = (long long)(SGI_LIST_BUDGET / (size_t)(n_tiles > 0 ? n_tiles : 1));
  if (cap > lim) cap = lim < 64 ? 64 : lim;
  return (int)cap;
}

// Programmatic dependent launch: the kernel may be scheduled while its predecessor on the stream is still draining; it waits
// (griddepcontrol.wait, first instruction that touches the predecessor's output) until that grid has completed and flushed.
// What overlaps is the launch latency and CTA start-up of k_order / k_tile, 2-3 us each on the latency chain of every pass.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
// Per-device function attributes (the opt-in to > 48 KB of dynamic shared memory applies to the CURRENT device only): kept per
// context, so a second context on another GPU of the same process configures its own device.
template <int MODE, int NT>
static int launch_tile_nt(sgi_ctx* ctx, const TileArgs& ta, dim3 grid, cudaStream_t stream) {
  constexpr size_t smem = tile_smem_bytes<MODE, NT>();
  constexpr int cfg_bit = MODE * 4 + (NT == 256 ? 0 : (NT == 512 ? 1 : (NT == 1024 ? 2 : 3)));
  if (!(ctx->func_cfg & (1ull << cfg_bit))) {
    SGI_CUDA(ctx, cudaFuncSetAttribute(k_tile<MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // ask for the largest shared-memory carve-out so that two 1024-thread CTAs (or more of the smaller ones) fit an SM;
    // with the default carve-out ncu showed occupancy limited to ONE CTA by shared memory
    SGI_CUDA(ctx, cudaFuncSetAttribute(k_tile<MODE, NT>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    ctx->func_cfg |= 1ull << cfg_bit;
  }
  const int pass = (MODE == SGI_MODE_DEPTH || MODE == SGI_MODE_MOMENTS) ? SGI_PASS_TILE_DEPTH : (SGI_KEYED(MODE) ? SGI_PASS_TILE_GBUFFER : SGI_PASS_TILE_SV);
  int tslot = sgi_timing_begin(ctx, pass, stream);
  SGI_CUDA(ctx, launch_pdl(k_tile<MODE, NT>, grid, dim3(NT), smem, stream, ctx->pdl, ta));
  sgi_timing_end(ctx, pass, tslot, stream);
  ctx->launches++;
  SGI_CUDA(ctx, cudaGetLastError());
  return SGI_OK;
}

// CTA size of the tile kernel.  Registers cap an SM at 1024 resident threads of this kernel whatever the CTA size, so the
// choice trades per-item latency (a hot item ends sooner with more warps) against barrier / init / flush overhead (cheaper
// with small CTAs, and four small CTAs hide each other's barrier waits).  Few-tile passes (<= 300 tiles: up to 720p /
// 1024^2) and shadow volumes (long lists, fill bound) get 32 warps per item, mid-size passes (<= 1200 tiles: 1080p / 2048^2)
// 16, many-tile passes 8.  Measured on B200 with busiest-first order and hot-tile subdivision on
// (profiles/r1_tile_cta_size.txt).  SGI_TILE_THREADS / "tile_threads" override for experiments.
static int tile_threads(const sgi_ctx* ctx, int n_tiles, int mode) {
  if (ctx->tile_threads) return ctx->tile_threads;
  if (mode == SGI_MODE_SVCOUNT) return 1024;
  // (round 2: with the default light of the c2 scene the 2048^2 depth pass prefers 256-thread CTAs, 75 vs 87 us, but over the
  //  bench's animated light it is 95 vs 59 us the other way: hot tiles want the larger CTA; the rule stays as measured in round 1)
  return n_tiles <= 300 ? 1024 : (n_tiles <= 1200 ? 512 : 256);
}

template <int MODE>
static int launch_tile(sgi_ctx* ctx, const TileArgs& ta, dim3 grid, int n_tiles, cudaStream_t stream) {
  switch (tile_threads(ctx, n_tiles, MODE)) {
    case 128: if (MODE == SGI_MODE_DEPTH) return launch_tile_nt<SGI_MODE_DEPTH, 128>(ctx, ta, grid, stream);   // (experiments; depth pass only)
    case 256: return launch_tile_nt<MODE, 256>(ctx, ta, grid, stream);
    case 512: return launch_tile_nt<MODE, 512>(ctx, ta, grid, stream);
    default: return launch_tile_nt<MODE, 1024>(ctx, ta, grid, stream);
  }
}

int sgi_raster_run(sgi_ctx* ctx, const SgiRasterJob& job, int scratch_set, cudaStream_t stream) {
  SgiScratch& sc = ctx->scratch[scratch_set];
  int rc = sgi_raster_reserve(ctx, sc, job.T, job.W, job.H, stream);
  if (rc) return rc;
  cudaStream_t st = stream;
  const int tiles_x = (job.W + SGI_TILE - 1) >> SGI_TILE_LOG2, tiles_y = (job.H + SGI_TILE - 1) >> SGI_TILE_LOG2;
  const int n_tiles = tiles_x * tiles_y;
  int rx0 = job.rx0, ry0 = job.ry0, rx1 = job.rx1, ry1 = job.ry1;
  if (rx1 <= rx0 || ry1 <= ry0) { rx0 = 0; ry0 = 0; rx1 = job.W; ry1 = job.H; }
  rx0 = rx0 < 0 ? 0 : rx0; ry0 = ry0 < 0 ? 0 : ry0; rx1 = rx1 > job.W ? job.W : rx1; ry1 = ry1 > job.H ? job.H : ry1;
  const int tx0 = rx0 >> SGI_TILE_LOG2, ty0 = ry0 >> SGI_TILE_LOG2;
  const int tx1 = (rx1 - 1) >> SGI_TILE_LOG2, ty1 = (ry1 - 1) >> SGI_TILE_LOG2;
  // size class of the pass: the moment and id passes bin exactly like the depth / G-buffer passes
  const int size_class = (job.mode == SGI_MODE_MOMENTS) ? SGI_MODE_DEPTH : (job.mode == SGI_MODE_SVCOUNT ? SGI_MODE_SVCOUNT : (job.mode == SGI_MODE_DEPTH ? SGI_MODE_DEPTH : SGI_MODE_GBUFFER));

  // a previous frame wanted longer lists than we had: grow before running again (2x headroom over the longest list seen;
  // the spill list takes a whole frame's pairs, so a frame is only ever incomplete when its pair total grows past that)
  int cap = sc.cap_of[size_class];
  if (sc.h_flags[1 + size_class] > cap) cap = list_capacity(sc.h_flags[1 + size_class], n_tiles);
  {
    const int total_seen = sc.h_flags[4 + size_class];
    if ((rc = sgi_raster_reserve_lists(ctx, sc, cap, n_tiles, total_seen > (1 << 16) ? total_seen : (1 << 16), st))) return rc;
  }
  sc.cap_of[size_class] = cap;

  if (sc.needs_clear) {     // live counters | snapshot | cursors in one allocation: cleared once; k_order re-zeroes what the binner dirtied
    SGI_CUDA(ctx, cudaMemsetAsync(sc.d_counters, 0, (size_t)(32 + sc.tile_cap) * 4, st));
    sc.needs_clear = false;
  }

  // shadow volumes with list-segment splitting accumulate their counts: zero the rectangle first (ahead of the binning chain, so
  // that the tile kernel still launches programmatically behind k_order)
  if (job.mode == SGI_MODE_SVCOUNT && ctx->sv_split_lists)
    SGI_CUDA(ctx, cudaMemset2DAsync(job.count + (size_t)ry0 * job.W + rx0, (size_t)job.W * 4, 0, (size_t)(rx1 - rx0) * 4, ry1 - ry0, st));

  // shadow volumes: per-tile farthest scene depth, so that the binner drops (prism, tile) pairs that lie behind the scene
  unsigned int* tile_zmax = nullptr;
  if (job.mode == SGI_MODE_SVCOUNT && job.scene_depth && ctx->sv_tile_cull && !job.sv_zfail) {
    tile_zmax = sc.d_tile_zmax;
    k_tile_zmax<<<dim3(tiles_x, tiles_y), 256, 0, st>>>(job.scene_depth, job.W, job.H, tiles_x, tile_zmax);
    ctx->launches++;
  }

  SetupBinArgs sa;
  sa.xyz = job.xyz; sa.nrm = job.nrm; sa.rgb = (job.rgb && job.albedo4) ? job.rgb : nullptr; sa.idx = job.idx; sa.T = job.T;
  const bool with_tex = job.mode == SGI_MODE_GBUFFER && job.uv && job.albedo4;
  sa.uv = with_tex ? job.uv : nullptr; sa.uvrec = with_tex ? sc.d_uvrec : nullptr;
  for (int k = 0; k < 16; k++) sa.mvp[k] = job.mvp[k];
  sa.W = job.W; sa.H = job.H; sa.use_offset = job.use_offset; sa.factor = job.factor; sa.units = job.units;
  sa.no_far_clip = job.no_far_clip;
  sa.rec = sc.d_rec; sa.attr = (job.mode == SGI_MODE_GBUFFER || job.mode == SGI_MODE_IDS) ? sc.d_attr : nullptr;
  sa.ovf_base = sc.d_ovf_base; sa.counters = sc.d_counters;
  sa.tiles_x = tiles_x; sa.tx0 = tx0; sa.ty0 = ty0; sa.tx1 = tx1; sa.ty1 = ty1;
  sa.tile_cnt = sc.d_tile_cnt; sa.pairs = sc.d_pairs; sa.cap = cap; sa.big_list = sc.d_big; sa.tile_zmax = tile_zmax;
  sa.spill = sc.d_spill; sa.spill_cap = sc.spill_cap;
  int sb_blocks = (job.T + SGI_SB_THREADS - 1) / SGI_SB_THREADS;
  if (sb_blocks < 1) sb_blocks = 1;

  // grid of the tile kernel = upper bound of its work items: every tile of the rectangle + room for subdivided hot tiles
  const int n_rect_tiles = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
  const bool seg_split = job.mode == SGI_MODE_SVCOUNT && ctx->sv_split_lists;
  const int max_items = seg_split ? n_rect_tiles + (63 * n_rect_tiles < SGI_SPLIT_EXTRA_SEG ? 63 * n_rect_tiles : SGI_SPLIT_EXTRA_SEG)
                                  : n_rect_tiles + (15 * n_rect_tiles < SGI_SPLIT_EXTRA ? 15 * n_rect_tiles : SGI_SPLIT_EXTRA);
  OrderArgs oa;
  oa.tile_cnt = sc.d_tile_cnt; oa.n_tiles = n_tiles; oa.cap = cap; oa.spill_cap = sc.spill_cap;
  oa.counters = sc.d_counters; oa.snap = sc.d_snap; oa.h_flags = sc.h_flags; oa.d_sticky = sc.d_sticky; oa.size_class = size_class;
  oa.order = sc.d_tile_order; oa.tiles_x = tiles_x; oa.tx0 = tx0; oa.ty0 = ty0; oa.gx = tx1 - tx0 + 1; oa.gy = ty1 - ty0 + 1;
  oa.busiest_first = ctx->tile_order; oa.max_items = max_items; oa.split_floor = ctx->tile_split; oa.n_sm = ctx->n_sm;
  oa.seg_split = seg_split ? 1 : 0; oa.max_level = seg_split ? 3 : 2;
  oa.mm_min = job.mm_min; oa.mm_max = job.mm_max; oa.mm_n = job.mm_min ? job.mm_w * (2 * tiles_y) : 0;   // (mm_w = 2 * tiles_x)

  // passes of many tiles bin their big records too (k_bin_big)
  // (worth it where the tile CTAs would otherwise do many record tests: big records seen in the last pass of this kind x tiles)
  const bool bin_big = ctx->tile_bin_big > 0 && n_rect_tiles >= ctx->tile_bin_big && (sc.h_flags[8 + size_class] != 0 || ctx->tile_bin_big_work <= 0);
  oa.big_work = ctx->tile_bin_big_work > 0 ? ctx->tile_bin_big_work : 1;
  oa.big_binned = bin_big ? 1 : 0;
  // with k_bin_big the set-up kernel keeps only the records of up to 16 tiles for its own walk; 17 .. 2048 tiles: a warp each,
  // beyond: all CTAs together.  Without it: records beyond SGI_BIG_TILES are tested by every tile CTA.
  sa.few_tiles = (ctx->tile_few_walk && n_rect_tiles <= 128) ? 1 : 0;
  sa.big_tiles = bin_big ? 16 : SGI_BIG_TILES; sa.huge_tiles = bin_big ? 2048 : 0x7FFFFFFF; sa.big_cap = sc.rec_cap_tris * 7 + 16;
  sc.needs_clear = true;                 // until k_order has been queued behind the binner
  k_setup_bin<<<sb_blocks, SGI_SB_THREADS, 0, st>>>(sa);
  if (bin_big) { SGI_CUDA(ctx, launch_pdl(k_bin_big, dim3(2 * ctx->n_sm), dim3(256), 0, st, ctx->pdl, sa)); ctx->launches++; }
  SGI_CUDA(ctx, seg_split ? launch_pdl(k_order<true>, dim3(1), dim3(1024), 0, st, ctx->pdl, oa) : launch_pdl(k_order<false>, dim3(1), dim3(1024), 0, st, ctx->pdl, oa));
  ctx->launches += 2;
  SGI_CUDA(ctx, cudaGetLastError());
  sc.needs_clear = false;
  if (!sc.sized[size_class]) {
    // first pass of this kind on this context: size the tile lists from the measured frame (one sync, once) and bin again
    SGI_CUDA(ctx, cudaStreamSynchronize(st));
    sc.sized[size_class] = true;
    const int longest = sc.h_flags[1 + size_class], total = sc.h_flags[4 + size_class];
    if (longest > cap || sc.h_flags[0]) {
      if (longest > cap) cap = list_capacity(longest, n_tiles);
      if ((rc = sgi_raster_reserve_lists(ctx, sc, cap, n_tiles, total > (1 << 16) ? total : (1 << 16), st))) return rc;
      sc.cap_of[size_class] = cap;
      sc.h_flags[0] = 0;                 // raised by the measuring run; the stream is idle
      sa.pairs = sc.d_pairs; sa.cap = cap; oa.cap = cap;
      sa.spill = sc.d_spill; sa.spill_cap = sc.spill_cap; oa.spill_cap = sc.spill_cap;
      sc.needs_clear = true;
      k_setup_bin<<<sb_blocks, SGI_SB_THREADS, 0, st>>>(sa);
      if (bin_big) { SGI_CUDA(ctx, launch_pdl(k_bin_big, dim3(2 * ctx->n_sm), dim3(256), 0, st, ctx->pdl, sa)); ctx->launches++; }
      SGI_CUDA(ctx, seg_split ? launch_pdl(k_order<true>, dim3(1), dim3(1024), 0, st, ctx->pdl, oa) : launch_pdl(k_order<false>, dim3(1), dim3(1024), 0, st, ctx->pdl, oa));
      ctx->launches += 2;
      SGI_CUDA(ctx, cudaGetLastError());
      sc.needs_clear = false;
    }
  }
  sc.overflow_pending = true;

  TileArgs ta;
  ta.rec = sc.d_rec; ta.attr = sc.d_attr; ta.ovf_base = sc.d_ovf_base;
  ta.tile_order = sc.d_tile_order; ta.pairs = sc.d_pairs; ta.cap = cap;
  ta.big_list = sc.d_big; ta.counters = sc.d_snap; ta.spill = sc.d_spill; ta.bulk_flush = ctx->tile_bulk_flush;
  ta.tiles_x = tiles_x; ta.tx0 = tx0; ta.ty0 = ty0;
  ta.W = job.W; ta.H = job.H; ta.rx0 = rx0; ta.ry0 = ry0; ta.rx1 = rx1; ta.ry1 = ry1;
  ta.xyz = job.xyz; ta.nrm = job.nrm; ta.idx = job.idx;
  ta.depth = job.depth; ta.pos4 = job.pos4; ta.nrm4 = job.nrm4;
  ta.rgb = job.rgb; ta.albedo4 = job.albedo4;
  ta.scene_depth = job.scene_depth; ta.depth_func = job.depth_func; ta.count = job.count; ta.stencil = job.stencil;
  ta.mom4 = job.mom4; ta.mom_tech = job.mom_tech; ta.z_near = job.z_near; ta.z_far = job.z_far; ta.ids = job.ids;
  ta.sv_zfail = job.sv_zfail; ta.sv_caps = job.sv_caps; ta.frag_counter = job.frag_counter;
  ta.sv_split_lists = (job.mode == SGI_MODE_SVCOUNT && ctx->sv_split_lists) ? 1 : 0;
  ta.mm_min = job.mm_min; ta.mm_max = job.mm_max; ta.mm_w = job.mm_w;
  ta.direct_max = ctx->tile_direct < 0 ? 0 : (ctx->tile_direct > 128 ? 128 : ctx->tile_direct);
  ta.static_items = ctx->tile_static_items; ta.refresh_full_only = ctx->tile_refresh_full;
  ta.uvrec = with_tex ? sc.d_uvrec : nullptr;
  for (int k = 0; k < 3; k++) ta.tex[k] = job.tex[k];
  for (int k = 0; k < 16; k++) ta.mq[k] = job.mq[k];
  for (int k = 0; k < 4; k++) ta.mqt[k] = job.mqt[k];
  dim3 grid(max_items);
  if (job.mode == SGI_MODE_DEPTH) rc = launch_tile<SGI_MODE_DEPTH>(ctx, ta, grid, n_rect_tiles, st);
  else if (job.mode == SGI_MODE_GBUFFER) rc = ((job.rgb || with_tex) && job.albedo4) ? launch_tile<SGI_MODE_GBUFFER_RGB>(ctx, ta, grid, n_rect_tiles, st) : launch_tile<SGI_MODE_GBUFFER>(ctx, ta, grid, n_rect_tiles, st);
  else if (job.mode == SGI_MODE_MOMENTS) rc = launch_tile<SGI_MODE_MOMENTS>(ctx, ta, grid, n_rect_tiles, st);
  else if (job.mode == SGI_MODE_IDS) rc = launch_tile<SGI_MODE_IDS>(ctx, ta, grid, n_rect_tiles, st);
  else {
    rc = launch_tile<SGI_MODE_SVCOUNT>(ctx, ta, grid, n_rect_tiles, st);
    if (rc == SGI_OK && ta.sv_split_lists) {
      k_sv_stencil<<<dim3((rx1 - rx0 + 255) / 256, ry1 - ry0), 256, 0, st>>>(job.count, job.stencil, job.W, rx0, ry0, rx1);
      ctx->launches++;
      SGI_CUDA(ctx, cudaGetLastError());
    }
  }
  return rc;
}

// Dilation of the block extrema over the reach of the shadow pass's tap window: out[b] = min / max over the (2R+1)^2 blocks around
// b, blocks outside the map counting as depth 0 for the minimum (CLAMP_TO_BORDER: a tap outside the map reads 0).  A pixel whose
// centre texel lies in block b has its whole window (reach <= 32 R - 1 texels) inside that neighbourhood.
namespace {
__global__ void __launch_bounds__(256) k_mm_dilate(const unsigned int* __restrict__ mn, const unsigned int* __restrict__ mx, int w, int h, int vw, int vh, int R,
                                                   float* __restrict__ dmin, float* __restrict__ dmax) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= w * h) return;
  const int bx = i % w, by = i / w;
  unsigned int lo = 0x3F800000u, hi = 0u;
  for (int dy = -R; dy <= R; dy++)
    for (int dx = -R; dx <= R; dx++) {
      const int x = bx + dx, y = by + dy;
      if (x < 0 || y < 0 || x >= vw || y >= vh) { lo = 0u; continue; }       // beyond the map: border depth 0
      lo = min(lo, mn[y * w + x]); hi = max(hi, mx[y * w + x]);
    }
  dmin[i] = __uint_as_float(lo); dmax[i] = __uint_as_float(hi);
}
}  // namespace

int sgi_minmax_dilate(sgi_ctx* ctx, int set, int R, cudaStream_t st) {
  const int n = ctx->mm_w * ctx->mm_h;
  unsigned int* base = ctx->d_mm;
  // blocks that hold map texels: ceil(S / 32) per axis (the arrays are padded to whole 64x64 tiles)
  k_mm_dilate<<<(n + 255) / 256, 256, 0, st>>>(base, base + n, ctx->mm_w, ctx->mm_h, (ctx->SW + 31) >> 5, (ctx->SH + 31) >> 5, R,
                                               (float*)(base + (size_t)(2 + 2 * set) * n), (float*)(base + (size_t)(3 + 2 * set) * n));
  ctx->launches++;
  SGI_CUDA(ctx, cudaGetLastError());
  return SGI_OK;
}

void sgi_raster_free(SgiScratch& sc) {
  void* ptrs[] = {sc.d_uvrec, sc.d_rec, sc.d_attr, sc.d_ovf_base, sc.d_big, sc.d_counters, sc.d_tile_order, sc.d_pairs, sc.d_tile_zmax, sc.d_spill};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (sc.h_flags) cudaFreeHost(sc.h_flags);
  if (sc.d_sticky) cudaFree(sc.d_sticky);
  sc = SgiScratch();
}

// Edge groups of the current mesh for the silhouette pass, built on the host once per index buffer (the topology does not
// depend on the light): entries sorted by (edge key, direction, triangle).
static int sv_prepare_edges(sgi_ctx* ctx) {
  if (ctx->sv_edges_T == ctx->T && ctx->sv_edges_valid) return SGI_OK;
  const int T = ctx->T;
  // the host copy of the index buffer: from now on sgi_set_mesh keeps it and only invalidates the groups when the indices change
  // (a caller that re-uploads the same mesh every frame, as the reference's loadVBOs does, pays one memcmp)
  std::vector<int32_t>& idx = ctx->h_idx_copy;
  if (!ctx->sv_track || (int)idx.size() != T * 3) {
    idx.resize((size_t)T * 3);
    SGI_CUDA(ctx, cudaStreamSynchronize(ctx->upload_stream));
    SGI_CUDA(ctx, cudaMemcpyAsync(idx.data(), ctx->d_idx, idx.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->sv_track = true;
  }
  struct Ent { long long key; int32_t ent; };
  std::vector<Ent> e;
  e.reserve((size_t)T * 3);
  for (int t = 0; t < T; t++)
    for (int k = 0; k < 3; k++) {
      const int a = idx[3 * (size_t)t + k], b = idx[3 * (size_t)t + (k + 1) % 3];
      if (a == b) continue;                                  // a degenerate edge extrudes a degenerate quad: nothing to pair
      const long long key = a < b ? ((long long)a << 32) | (unsigned)b : ((long long)b << 32) | (unsigned)a;
      e.push_back({key, ((3 * t + k) << 1) | (a < b ? 1 : 0)});
    }
  std::sort(e.begin(), e.end(), [](const Ent& x, const Ent& y) {
    if (x.key != y.key) return x.key < y.key;
    if ((x.ent & 1) != (y.ent & 1)) return (x.ent & 1) < (y.ent & 1);
    return x.ent < y.ent;
  });
  std::vector<int32_t> start, ent(e.size());
  for (size_t i = 0; i < e.size(); i++) {
    if (i == 0 || e[i].key != e[i - 1].key) start.push_back((int32_t)i);
    ent[i] = e[i].ent;
  }
  const int G = (int)start.size();
  start.push_back((int32_t)e.size());
  if (ctx->d_sv_grp_start) cudaFree(ctx->d_sv_grp_start);
  if (ctx->d_sv_grp_ent) cudaFree(ctx->d_sv_grp_ent);
  ctx->d_sv_grp_start = ctx->d_sv_grp_ent = nullptr;
  SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_sv_grp_start, start.size() * 4));
  SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_sv_grp_ent, (ent.size() + 1) * 4));
  SGI_CUDA(ctx, cudaMemcpy(ctx->d_sv_grp_start, start.data(), start.size() * 4, cudaMemcpyHostToDevice));
  if (!ent.empty()) SGI_CUDA(ctx, cudaMemcpy(ctx->d_sv_grp_ent, ent.data(), ent.size() * 4, cudaMemcpyHostToDevice));
  ctx->sv_groups = G; ctx->sv_edges_T = T; ctx->sv_edges_valid = true;
  return SGI_OK;
}

int sgi_sv_extrude_run(sgi_ctx* ctx, const float light[3], float* prism_xyz, int32_t* prism_idx, int per, int silhouette) {
  if (ctx->T <= 0) return SGI_OK;
  if (silhouette) {
    int rc = sv_prepare_edges(ctx);
    if (rc) return rc;
    if (ctx->sv_cls_cap < ctx->T) {
      SGI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (ctx->d_sv_cls) cudaFree(ctx->d_sv_cls);
      ctx->d_sv_cls = nullptr; ctx->sv_cls_cap = 0;
      SGI_CUDA(ctx, cudaMalloc((void**)&ctx->d_sv_cls, (size_t)ctx->T));
      ctx->sv_cls_cap = ctx->T;
    }
  }
  k_sv_extrude<<<(ctx->T + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_xyz, ctx->d_nrm, ctx->d_idx, ctx->T, light[0], light[1],
                                                             light[2], ctx->params.sv_infinity, per, prism_xyz, prism_idx,
                                                             silhouette ? ctx->d_sv_cls : nullptr);
  ctx->launches++;
  if (silhouette && ctx->sv_groups > 0) {
    k_sv_silhouette<<<(ctx->sv_groups + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_sv_grp_start, ctx->d_sv_grp_ent, ctx->sv_groups, ctx->d_sv_cls, per, prism_idx);
    ctx->launches++;
  }
  SGI_CUDA(ctx, cudaGetLastError());
  return SGI_OK;
}
