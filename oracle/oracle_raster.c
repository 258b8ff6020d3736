/*
 * oracle_raster.c — CPU definition of the fixed-function rasteriser the reference relies on
 * (OpenGL driver; no source in /root/reference) + the passes built on it.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h).  Rules = DESIGN.md §3; callers in the reference: glDrawElements at
 * ShadowMapping/src/Viewers/MyGLGeometryViewer.cpp:347, glPolygonOffset ShadowMapping/src/main.cpp:246,
 * glStencilOpSeparate ShadowVolumes/src/main.cpp:167-168.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GUARD 16.0f          /* guard-band factor for x/y clipping                    */
#define SUBPIX 256           /* 8 sub-pixel bits                                      */
#define MAXPOLY 10

typedef struct { float x, y, z, w; float b[3]; } CV;   /* clip vertex + barycentrics wrt source tri */

typedef struct {
  int32_t X[3], Y[3];        /* snapped window coords (1/256 px), CCW order           */
  float z0, dz1, dz2;        /* window depth at v0 and deltas                          */
  float ia;                  /* 1/(float)area2                                         */
  float iw[3];               /* 1/w_clip per vertex                                    */
  float bary[3][3];          /* barycentrics of the 3 vertices wrt the source triangle */
  int64_t area2;
  float zoff;                /* polygon offset (0 if disabled)                         */
  int32_t front;             /* gl_FrontFacing                                         */
  int32_t clipped;           /* 0: the source triangle itself (attributes taken from its vertices as they are) */
  int32_t prim;              /* source triangle * 8 + fan index                        */
  int32_t px0, py0, px1, py1;/* inclusive pixel bbox, clamped to the viewport          */
} SubTri;

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static inline float plane_dist(const CV* v, int p) {
  switch (p) {
    case 0: return v->w + v->z;
    case 1: return v->w - v->z;
    case 2: return GUARD * v->w + v->x;
    case 3: return GUARD * v->w - v->x;
    case 4: return GUARD * v->w + v->y;
    default: return GUARD * v->w - v->y;
  }
}

/* a is inside (da>=0), b outside (db<0): point on the plane, always evaluated inside -> outside */
static inline CV clip_lerp(const CV* a, const CV* b, float da, float db) {
  float t = da / (da - db);
  CV r;
  r.x = a->x + t * (b->x - a->x);
  r.y = a->y + t * (b->y - a->y);
  r.z = a->z + t * (b->z - a->z);
  r.w = a->w + t * (b->w - a->w);
  for (int k = 0; k < 3; k++) r.b[k] = a->b[k] + t * (b->b[k] - a->b[k]);
  return r;
}

/* depth clamp (GL_DEPTH_CLAMP semantics for the far plane only): set around the depth-fail shadow-volume pass; the far plane then
 * neither rejects nor clips and fragment depths saturate at 1 (frag_z clamps to [0,1]) */
static int g_no_far_clip = 0;

static int clip_polygon(CV* poly, int n, int* clipped) {
  CV tmp[MAXPOLY];
  *clipped = 0;
  for (int p = 0; p < 6; p++) {
    if (p == 1 && g_no_far_clip) continue;
    int any_out = 0;
    float d[MAXPOLY];
    for (int i = 0; i < n; i++) { d[i] = plane_dist(&poly[i], p); if (!(d[i] >= 0.0f)) any_out = 1; }
    if (!any_out) continue;
    *clipped = 1;
    int m = 0;
    for (int i = 0; i < n; i++) {
      int j = (i + 1 == n) ? 0 : i + 1;
      int in_i = d[i] >= 0.0f, in_j = d[j] >= 0.0f;
      if (in_i) {
        tmp[m++] = poly[i];
        if (!in_j) tmp[m++] = clip_lerp(&poly[i], &poly[j], d[i], d[j]);
      } else if (in_j) {
        tmp[m++] = clip_lerp(&poly[j], &poly[i], d[j], d[i]);
      }
    }
    n = m;
    if (n < 3) return 0;
    memcpy(poly, tmp, sizeof(CV) * n);
  }
  return n;
}

static inline void xform(const float* m, const float* v, CV* o) {
  float x = v[0], y = v[1], z = v[2];
  o->x = ((m[0] * x + m[4] * y) + m[8] * z) + m[12];
  o->y = ((m[1] * x + m[5] * y) + m[9] * z) + m[13];
  o->z = ((m[2] * x + m[6] * y) + m[10] * z) + m[14];
  o->w = ((m[3] * x + m[7] * y) + m[11] * z) + m[15];
}

/* Build the sub-triangle records of one source triangle.  Returns count (0..7). */
static int setup_triangle(const float* mvp, const float* p0, const float* p1, const float* p2, int tri,
                          int W, int H, int use_offset, float factor, float units, SubTri* out) {
  CV poly[MAXPOLY];
  xform(mvp, p0, &poly[0]); xform(mvp, p1, &poly[1]); xform(mvp, p2, &poly[2]);
  poly[0].b[0] = 1; poly[0].b[1] = 0; poly[0].b[2] = 0;
  poly[1].b[0] = 0; poly[1].b[1] = 1; poly[1].b[2] = 0;
  poly[2].b[0] = 0; poly[2].b[1] = 0; poly[2].b[2] = 1;
  /* trivial reject against the true frustum */
  {
    int o[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 3; k++) {
      const CV* v = &poly[k];
      if (!(v->w + v->z >= 0.0f)) o[0]++;
      if (!(v->w - v->z >= 0.0f)) o[1]++;
      if (!(v->w + v->x >= 0.0f)) o[2]++;
      if (!(v->w - v->x >= 0.0f)) o[3]++;
      if (!(v->w + v->y >= 0.0f)) o[4]++;
      if (!(v->w - v->y >= 0.0f)) o[5]++;
    }
    if (g_no_far_clip) o[1] = 0;
    for (int p = 0; p < 6; p++) if (o[p] == 3) return 0;
  }
  int was_clipped;
  int n = clip_polygon(poly, 3, &was_clipped);
  if (n < 3) return 0;
  float hw = (float)W * 0.5f, hh = (float)H * 0.5f;
  int32_t X[MAXPOLY], Y[MAXPOLY];
  float Z[MAXPOLY], IW[MAXPOLY];
  for (int k = 0; k < n; k++) {
    float nx = poly[k].x / poly[k].w, ny = poly[k].y / poly[k].w, nz = poly[k].z / poly[k].w;
    float xw = nx * hw + hw, yw = ny * hh + hh;
    Z[k] = nz * 0.5f + 0.5f;
    IW[k] = 1.0f / poly[k].w;
    float sx = xw * (float)SUBPIX, sy = yw * (float)SUBPIX;
    if (!(fabsf(sx) < 1.0e9f) || !(fabsf(sy) < 1.0e9f) || !(fabsf(Z[k]) < 1.0e9f)) return 0;
    X[k] = (int32_t)lrintf(sx);
    Y[k] = (int32_t)lrintf(sy);
  }
  int cnt = 0;
  for (int f = 1; f + 1 < n; f++) {
    int id[3] = {0, f, f + 1};
    int64_t area2 = (int64_t)(X[id[1]] - X[id[0]]) * (int64_t)(Y[id[2]] - Y[id[0]]) -
                    (int64_t)(X[id[2]] - X[id[0]]) * (int64_t)(Y[id[1]] - Y[id[0]]);
    if (area2 == 0) continue;
    SubTri* s = &out[cnt];
    s->front = area2 > 0;
    if (area2 < 0) { int t = id[1]; id[1] = id[2]; id[2] = t; area2 = -area2; }
    int32_t mnx = X[id[0]], mxx = mnx, mny = Y[id[0]], mxy = mny;
    for (int k = 0; k < 3; k++) {
      s->X[k] = X[id[k]]; s->Y[k] = Y[id[k]]; s->iw[k] = IW[id[k]];
      for (int q = 0; q < 3; q++) s->bary[k][q] = poly[id[k]].b[q];
      if (s->X[k] < mnx) mnx = s->X[k];
      if (s->X[k] > mxx) mxx = s->X[k];
      if (s->Y[k] < mny) mny = s->Y[k];
      if (s->Y[k] > mxy) mxy = s->Y[k];
    }
    s->area2 = area2;
    s->ia = 1.0f / (float)area2;
    s->z0 = Z[id[0]]; s->dz1 = Z[id[1]] - Z[id[0]]; s->dz2 = Z[id[2]] - Z[id[0]];
    s->prim = tri * 8 + (f - 1);
    s->clipped = was_clipped;
    s->zoff = 0.0f;
    if (use_offset) {
      double dY1 = (double)(s->Y[1] - s->Y[0]), dY2 = (double)(s->Y[2] - s->Y[0]);
      double dX1 = (double)(s->X[1] - s->X[0]), dX2 = (double)(s->X[2] - s->X[0]);
      double nx = (double)s->dz1 * dY2 - (double)s->dz2 * dY1;
      double ny = (double)s->dz2 * dX1 - (double)s->dz1 * dX2;
      double dzdx = nx / (double)area2 * (double)SUBPIX;
      double dzdy = ny / (double)area2 * (double)SUBPIX;
      float m = (float)fmax(fabs(dzdx), fabs(dzdy));
      float zmax = Z[id[0]];
      if (Z[id[1]] > zmax) zmax = Z[id[1]];
      if (Z[id[2]] > zmax) zmax = Z[id[2]];
      float r = 0.0f;
      if (zmax > 0.0f) {
        int eb = (int)((f2u(zmax) >> 23) & 255u);
        if (eb > 23 && eb < 255) r = u2f((uint32_t)(eb - 23) << 23);
      }
      s->zoff = factor * m + units * r;
    }
    s->px0 = (mnx - SUBPIX / 2 + (SUBPIX - 1)) >> 8;
    s->px1 = (mxx - SUBPIX / 2) >> 8;
    s->py0 = (mny - SUBPIX / 2 + (SUBPIX - 1)) >> 8;
    s->py1 = (mxy - SUBPIX / 2) >> 8;
    if (s->px0 < 0) s->px0 = 0;
    if (s->py0 < 0) s->py0 = 0;
    if (s->px1 > W - 1) s->px1 = W - 1;
    if (s->py1 > H - 1) s->py1 = H - 1;
    if (s->px0 > s->px1 || s->py0 > s->py1) continue;
    cnt++;
  }
  return cnt;
}

/* coverage + edge values at pixel (i,j); returns 1 if covered */
static inline int cover(const SubTri* s, int i, int j, int64_t E[3]) {
  int64_t px = (int64_t)i * SUBPIX + SUBPIX / 2, py = (int64_t)j * SUBPIX + SUBPIX / 2;
  for (int e = 0; e < 3; e++) {
    int a = (e + 1) % 3, b = (e + 2) % 3;
    int64_t dx = (int64_t)s->X[b] - s->X[a], dy = (int64_t)s->Y[b] - s->Y[a];
    int64_t v = dx * (py - s->Y[a]) - dy * (px - s->X[a]);
    if (v < 0) return 0;
    if (v == 0 && !(dy < 0 || (dy == 0 && dx < 0))) return 0;
    E[e] = v;
  }
  return 1;
}

static inline float frag_z(const SubTri* s, const int64_t E[3]) {
  float b1 = (float)E[1] * s->ia, b2 = (float)E[2] * s->ia;
  float z = (s->z0 + b1 * s->dz1) + b2 * s->dz2;
  z = z + s->zoff;
  if (!(z >= 0.0f)) z = 0.0f;
  if (z > 1.0f) z = 1.0f;
  return z;
}

static SubTri* build_records(const float* xyz, const int32_t* idx, int T, const float* mvp, int W, int H,
                             int use_offset, float factor, float units, int64_t* n_out) {
  SubTri* rec = (SubTri*)malloc(sizeof(SubTri) * (size_t)(T > 0 ? T : 1) * 7);
  int* cnt = (int*)malloc(sizeof(int) * (size_t)(T > 0 ? T : 1));
  if (!rec || !cnt) { free(rec); free(cnt); return NULL; }
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T; t++)
    cnt[t] = setup_triangle(mvp, xyz + 3 * (size_t)idx[3 * t], xyz + 3 * (size_t)idx[3 * t + 1],
                            xyz + 3 * (size_t)idx[3 * t + 2], t, W, H, use_offset, factor, units,
                            rec + (size_t)t * 7);
  int64_t n = 0;
  for (int t = 0; t < T; t++) {          /* compact, keeping draw order */
    for (int k = 0; k < cnt[t]; k++) { if (n != (int64_t)t * 7 + k) rec[n] = rec[(size_t)t * 7 + k]; n++; }
  }
  free(cnt);
  *n_out = n;
  return rec;
}

int orc_raster_depth(const float* xyz, int V, const int32_t* idx, int T, const float mvp[16], int W, int H,
                     float factor, float units, float* depth) {
  (void)V;
  int64_t n;
  SubTri* rec = build_records(xyz, idx, T, mvp, W, H, 1, factor, units, &n);
  if (!rec) return -1;
  for (size_t i = 0; i < (size_t)W * H; i++) depth[i] = 1.0f;
  int bands = H < 64 ? 1 : 64;
#pragma omp parallel for schedule(dynamic, 1)
  for (int bnd = 0; bnd < bands; bnd++) {
    int r0 = (int)((int64_t)H * bnd / bands), r1 = (int)((int64_t)H * (bnd + 1) / bands) - 1;
    for (int64_t k = 0; k < n; k++) {
      const SubTri* s = &rec[k];
      int y0 = s->py0 > r0 ? s->py0 : r0, y1 = s->py1 < r1 ? s->py1 : r1;
      for (int j = y0; j <= y1; j++)
        for (int i = s->px0; i <= s->px1; i++) {
          int64_t E[3];
          if (!cover(s, i, j, E)) continue;
          float z = frag_z(s, E);
          float* d = &depth[(size_t)j * W + i];
          if (z < *d) *d = z;
        }
    }
  }
  free(rec);
  return 0;
}

int orc_raster_gbuffer(const float* xyz, const float* nrm, int V, const int32_t* idx, int T, const float mvp[16],
                       int W, int H, float* pos4, float* nrm4, float* depth) {
  return orc_raster_gbuffer_ex(xyz, nrm, NULL, V, idx, T, mvp, W, H, pos4, nrm4, NULL, depth);
}

int orc_raster_gbuffer_ex(const float* xyz, const float* nrm, const float* rgb, int V, const int32_t* idx, int T,
                          const float mvp[16], int W, int H, float* pos4, float* nrm4, float* albedo4, float* depth) {
  return orc_raster_gbuffer_tex(xyz, nrm, rgb, NULL, V, idx, T, mvp, W, H, NULL, pos4, nrm4, albedo4, depth);
}

/* ---- GBuffer.frag:11-30 computeFragmentColor with useTextureForColoring == 1 -------------------------------------------
 * texture2D on the scene textures of loadRGBTexture (MyGLTextureViewer.cpp:45-56): RGB8, GL_LINEAR, GL_REPEAT, no mipmaps.  GL
 * leaves the filter's arithmetic to the implementation; DEFINED here as for the other bilinear lookups (oracle_moments_impl.h):
 * weights fract(u*size - 0.5), texels (byte / 255.0f, alpha 1) accumulated in the order 00, 10, 01, 11, indices wrapped. */
static inline float wrap_index(float f, float size) {
  f = f - floorf(f / size) * size;
  if (!(f < size)) f = 0.0f;
  return f;
}
static void tex_fetch_linear_repeat(const orc_texture* t, float u, float v, float out[4]) {
  float fw = (float)t->w, fh = (float)t->h;
  float x = u * fw - 0.5f, y = v * fh - 0.5f;
  float x0 = floorf(x), y0 = floorf(y), ax = x - x0, ay = y - y0;
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  for (int k = 0; k < 4; k++) {
    float wgt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
    float fi = wrap_index(x0 + (float)(k & 1), fw), fj = wrap_index(y0 + (float)(k >> 1), fh);
    float tx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (fi >= 0.0f && fi < fw && fj >= 0.0f && fj < fh) {      /* (NaN coordinates fetch the border colour 0) */
      const uint8_t* p = t->rgb + 3 * ((size_t)(int)fj * t->w + (size_t)(int)fi);
      tx[0] = (float)p[0] / 255.0f; tx[1] = (float)p[1] / 255.0f; tx[2] = (float)p[2] / 255.0f; tx[3] = 1.0f;
    }
    for (int c = 0; c < 4; c++) out[c] = out[c] + tx[c] * wgt;
  }
}
void orc_fragment_color(const float uvw[3], const float rgb[3], const orc_texture tex[3], float out[4]) {
  float b = uvw[2];
  int sel = -1;
  if (b > 0.99f && b < 1.001f) sel = 0;                  /* :16 */
  else if (b > 1.999f && b < 2.001f) sel = 1;            /* :18 */
  else if (b > 2.999f && b < 3.001f) sel = 2;            /* :20 */
  if (sel >= 0 && tex && tex[sel].rgb && tex[sel].w > 0 && tex[sel].h > 0) { tex_fetch_linear_repeat(&tex[sel], uvw[0], uvw[1], out); return; }
  if (sel >= 0) { out[0] = out[1] = out[2] = out[3] = 0.0f; return; }   /* an unbound sampler reads (0,0,0,0): incomplete texture */
  out[0] = rgb[0]; out[1] = rgb[1]; out[2] = rgb[2]; out[3] = 1.0f;      /* :22-23 */
}

/* The G-buffer with the third MRT of GBuffer.frag:32-38: uv == NULL / tex == NULL: useMeshColor form (interpolated vertex colour, 1);
 * otherwise the texture select above on the perspective-correct (u, v, texture id) varying.  rgb may be NULL (colour 0: a disabled
 * vertex attribute).  Background (0,0,0,1). */
int orc_raster_gbuffer_tex(const float* xyz, const float* nrm, const float* rgb, const float* uv, int V, const int32_t* idx, int T,
                           const float mvp[16], int W, int H, const orc_texture* tex, float* pos4, float* nrm4, float* albedo4, float* depth) {
  (void)V;
  const int with_tex = uv != NULL && tex != NULL && albedo4 != NULL;
  const int with_rgb = (rgb != NULL || with_tex) && albedo4 != NULL;
  int64_t n;
  SubTri* rec = build_records(xyz, idx, T, mvp, W, H, 0, 0.0f, 0.0f, &n);
  if (!rec) return -1;
  for (size_t i = 0; i < (size_t)W * H; i++) {
    depth[i] = 1.0f;
    pos4[4 * i + 0] = 0; pos4[4 * i + 1] = 0; pos4[4 * i + 2] = 0; pos4[4 * i + 3] = 1;
    nrm4[4 * i + 0] = 0; nrm4[4 * i + 1] = 0; nrm4[4 * i + 2] = 0; nrm4[4 * i + 3] = 1;
    if (with_rgb) { albedo4[4 * i + 0] = 0; albedo4[4 * i + 1] = 0; albedo4[4 * i + 2] = 0; albedo4[4 * i + 3] = 1; }
  }
  int bands = H < 64 ? 1 : 64;
#pragma omp parallel for schedule(dynamic, 1)
  for (int bnd = 0; bnd < bands; bnd++) {
    int r0 = (int)((int64_t)H * bnd / bands), r1 = (int)((int64_t)H * (bnd + 1) / bands) - 1;
    for (int64_t k = 0; k < n; k++) {
      const SubTri* s = &rec[k];
      int y0 = s->py0 > r0 ? s->py0 : r0, y1 = s->py1 < r1 ? s->py1 : r1;
      if (y0 > y1) continue;
      int t = s->prim >> 3;
      const int32_t* ix = idx + 3 * (size_t)t;
      /* attributes of the sub-triangle's vertices: the source vertices themselves (in the record's CCW order) when
         nothing was clipped, otherwise their barycentric combination */
      float A[3][12];
      const float* arr[4] = {xyz, nrm, rgb, with_tex ? uv : NULL};
      for (int v = 0; v < 3; v++)
        for (int g = 0; g < 4; g++)
          for (int c = 0; c < 3; c++) {
            const float* a = arr[g];
            float val = 0.0f;
            if (a) {
              if (!s->clipped) {
                int src = s->bary[v][0] == 1.0f ? 0 : (s->bary[v][1] == 1.0f ? 1 : 2);
                val = a[3 * (size_t)ix[src] + c];
              } else {
                val = (s->bary[v][0] * a[3 * (size_t)ix[0] + c] + s->bary[v][1] * a[3 * (size_t)ix[1] + c]) + s->bary[v][2] * a[3 * (size_t)ix[2] + c];
              }
            }
            A[v][3 * g + c] = val;
          }
      for (int j = y0; j <= y1; j++)
        for (int i = s->px0; i <= s->px1; i++) {
          int64_t E[3];
          if (!cover(s, i, j, E)) continue;
          float z = frag_z(s, E);
          size_t o = (size_t)j * W + i;
          if (!(z < depth[o])) continue;
          depth[o] = z;
          float q0 = ((float)E[0] * s->ia) * s->iw[0];
          float q1 = ((float)E[1] * s->ia) * s->iw[1];
          float q2 = ((float)E[2] * s->ia) * s->iw[2];
          float iq = 1.0f / ((q0 + q1) + q2);
          float col[3], uvw[3];
          for (int c = 0; c < 3; c++) {
            pos4[4 * o + c] = ((q0 * A[0][c] + q1 * A[1][c]) + q2 * A[2][c]) * iq;
            nrm4[4 * o + c] = ((q0 * A[0][3 + c] + q1 * A[1][3 + c]) + q2 * A[2][3 + c]) * iq;
            col[c] = ((q0 * A[0][6 + c] + q1 * A[1][6 + c]) + q2 * A[2][6 + c]) * iq;
            uvw[c] = ((q0 * A[0][9 + c] + q1 * A[1][9 + c]) + q2 * A[2][9 + c]) * iq;
          }
          if (with_tex) orc_fragment_color(uvw, col, tex, albedo4 + 4 * o);
          else if (with_rgb) { albedo4[4 * o] = col[0]; albedo4[4 * o + 1] = col[1]; albedo4[4 * o + 2] = col[2]; albedo4[4 * o + 3] = 1.0f; }
          pos4[4 * o + 3] = 1.0f;
          nrm4[4 * o + 3] = s->front ? 1.0f : 0.0f;
        }
    }
  }
  free(rec);
  return 0;
}

/* ---- moment shadow maps: the light-view pass with Moments.frag / Exponential.frag / ExponentialMoments.frag bound
 * (ShadowMapping/src/main.cpp:227-243,350-361).  The depth test uses the polygon-offset depth (GL_LESS, first fragment in
 * draw order wins ties); the colour written is a function of the un-offset plane depth of the winning fragment and of the
 * same plane at its two 2x2-quad partners (oracle_moments_impl.h).  mom4: float4[H][W], cleared to (0,0,0,1) (:356). */
static inline float plane_z(const SubTri* s, int i, int j) {
  int64_t px = (int64_t)i * SUBPIX + SUBPIX / 2, py = (int64_t)j * SUBPIX + SUBPIX / 2;
  int64_t E1 = ((int64_t)s->X[0] - s->X[2]) * (py - s->Y[2]) - ((int64_t)s->Y[0] - s->Y[2]) * (px - s->X[2]);
  int64_t E2 = ((int64_t)s->X[1] - s->X[0]) * (py - s->Y[0]) - ((int64_t)s->Y[1] - s->Y[0]) * (px - s->X[0]);
  float b1 = (float)E1 * s->ia, b2 = (float)E2 * s->ia;
  return (s->z0 + b1 * s->dz1) + b2 * s->dz2;
}

int orc_raster_moments(const float* xyz, int V, const int32_t* idx, int T, const float mvp[16], int W, int H, float factor,
                       float units, int technique, int z_near, int z_far, float* mom4) {
  (void)V;
  int64_t n;
  SubTri* rec = build_records(xyz, idx, T, mvp, W, H, 1, factor, units, &n);
  if (!rec) return -1;
  float* depth = (float*)malloc(sizeof(float) * (size_t)W * H);
  int64_t* win = (int64_t*)malloc(sizeof(int64_t) * (size_t)W * H);
  if (!depth || !win) { free(rec); free(depth); free(win); return -1; }
  for (size_t i = 0; i < (size_t)W * H; i++) { depth[i] = 1.0f; win[i] = -1; }
  int bands = H < 64 ? 1 : 64;
#pragma omp parallel for schedule(dynamic, 1)
  for (int bnd = 0; bnd < bands; bnd++) {
    int r0 = (int)((int64_t)H * bnd / bands), r1 = (int)((int64_t)H * (bnd + 1) / bands) - 1;
    for (int64_t k = 0; k < n; k++) {
      const SubTri* s = &rec[k];
      int y0 = s->py0 > r0 ? s->py0 : r0, y1 = s->py1 < r1 ? s->py1 : r1;
      for (int j = y0; j <= y1; j++)
        for (int i = s->px0; i <= s->px1; i++) {
          int64_t E[3];
          if (!cover(s, i, j, E)) continue;
          float z = frag_z(s, E);
          size_t o = (size_t)j * W + i;
          if (z < depth[o]) { depth[o] = z; win[o] = k; }
        }
    }
  }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      size_t o = (size_t)j * W + i;
      float* m = mom4 + 4 * o;
      if (win[o] < 0) { m[0] = 0.0f; m[1] = 0.0f; m[2] = 0.0f; m[3] = 1.0f; continue; }
      const SubTri* s = &rec[win[o]];
      orc_moment_texel(technique, plane_z(s, i, j), plane_z(s, i ^ 1, j), plane_z(s, i, j ^ 1), i & 1, j & 1, z_near, z_far, m);
    }
  free(rec); free(depth); free(win);
  return 0;
}

/* ---- shadow volumes: ShadowVolumes/src/ShadowVolume.cpp:15-114 (build) == :116-195 (update) ---- */
void orc_sv_build_prisms(const float* xyz, const float* nrm, int V, const int32_t* idx, int T, const float light[3],
                         int infinity, float* prism_xyz, int32_t* prism_idx) {
  (void)V;
  static const int ORD_A[18] = {1, 0, 3, 1, 3, 4, 2, 1, 4, 2, 4, 5, 0, 2, 5, 0, 5, 3};   /* :64-83  */
  static const int ORD_B[18] = {4, 3, 0, 4, 0, 1, 5, 4, 1, 5, 1, 2, 3, 5, 2, 3, 2, 0};   /* :89-108 */
#pragma omp parallel for schedule(static)
  for (int t = 0; t < T; t++) {
    int v[3] = {idx[3 * t], idx[3 * t + 1], idx[3 * t + 2]};
    float* q = prism_xyz + (size_t)t * 18;
    for (int a = 0; a < 3; a++)
      for (int k = 0; k < 3; k++) {
        q[k * 3 + a] = xyz[3 * (size_t)v[k] + a];
        q[(3 + k) * 3 + a] = (xyz[3 * (size_t)v[k] + a] - light[a]) * (float)infinity;   /* :39-41 */
      }
    float n[3];
    for (int a = 0; a < 3; a++) {
      n[a] = nrm[3 * (size_t)v[0] + a] + nrm[3 * (size_t)v[1] + a] + nrm[3 * (size_t)v[2] + a];   /* :58 */
      n[a] = n[a] / 3.0f;                                                                         /* :59 */
    }
    float d = n[0] * light[0] + n[1] * light[1] + n[2] * light[2];                                /* glm::dot */
    const int* ord = (d >= 0.0f) ? ORD_A : ORD_B;
    for (int k = 0; k < 18; k++) prism_idx[(size_t)t * 18 + k] = t * 6 + ord[k];
  }
}

/* Silhouette form (north_star (4); SURVEY F1: an optimisation of the same counts).  The two side quads that two triangles of the
 * same orientation class extrude from a shared edge, walked in opposite directions, cover the same pixels with opposite facing:
 * they are dropped in pairs.  Per undirected edge (vertex-index pair) and class, min(#forward, #backward) pairs go, lowest
 * triangle indices first; what stays are the quads of silhouette edges (classes differ), boundary edges and the unpaired rest of
 * non-manifold edges.  keep[3 t + e] = 1 if the quad of triangle t's edge e (v_e -> v_(e+1)%3) is drawn. */
typedef struct { int64_t key; int32_t te; int32_t cls_dir; } EdgeEnt;
static int edge_cmp(const void* a, const void* b) {
  const EdgeEnt* x = (const EdgeEnt*)a; const EdgeEnt* y = (const EdgeEnt*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  if (x->cls_dir != y->cls_dir) return x->cls_dir < y->cls_dir ? -1 : 1;
  return x->te < y->te ? -1 : (x->te > y->te ? 1 : 0);
}
void orc_sv_silhouette_keep(const float* nrm, const int32_t* idx, int T, const float light[3], uint8_t* keep) {
  EdgeEnt* e = (EdgeEnt*)malloc(sizeof(EdgeEnt) * (size_t)(T > 0 ? T : 1) * 3);
  for (int t = 0; t < T; t++) {
    int v[3] = {idx[3 * t], idx[3 * t + 1], idx[3 * t + 2]};
    float n[3];
    for (int a = 0; a < 3; a++) {
      n[a] = nrm[3 * (size_t)v[0] + a] + nrm[3 * (size_t)v[1] + a] + nrm[3 * (size_t)v[2] + a];
      n[a] = n[a] / 3.0f;
    }
    float d = n[0] * light[0] + n[1] * light[1] + n[2] * light[2];
    int cls = d >= 0.0f ? 1 : 0;                                     /* the orientation rule of ShadowVolume.cpp:55-61 */
    for (int k = 0; k < 3; k++) {
      int a = v[k], b = v[(k + 1) % 3];
      EdgeEnt* q = &e[3 * (size_t)t + k];
      q->key = a < b ? ((int64_t)a << 32) | (uint32_t)b : ((int64_t)b << 32) | (uint32_t)a;
      q->te = 3 * t + k;
      q->cls_dir = cls * 2 + (a < b ? 1 : 0);
      keep[3 * (size_t)t + k] = 1;
      if (a == b) q->key = -1 - (int64_t)q->te;                      /* a degenerate edge (its quad is degenerate too) pairs with nothing */
    }
  }
  qsort(e, (size_t)T * 3, sizeof(EdgeEnt), edge_cmp);
  size_t n = (size_t)T * 3, i = 0;
  while (i < n) {
    size_t j = i;
    while (j < n && e[j].key == e[i].key) j++;
    for (int cls = 0; cls < 2; cls++) {                              /* entries are sorted by (class, direction, triangle) */
      size_t b0 = i; while (b0 < j && e[b0].cls_dir < cls * 2) b0++;
      size_t f0 = b0; while (f0 < j && e[f0].cls_dir < cls * 2 + 1) f0++;
      size_t f1 = f0; while (f1 < j && e[f1].cls_dir < cls * 2 + 2) f1++;
      size_t nb = f0 - b0, nf = f1 - f0, k = nb < nf ? nb : nf;
      for (size_t m = 0; m < k; m++) { keep[e[b0 + m].te] = 0; keep[e[f0 + m].te] = 0; }
    }
    i = j;
  }
  free(e);
}

/* Volumes with 6 (sides) or 8 (sides + near cap + far cap) triangles per source triangle; vertices as orc_sv_build_prisms.
 * Quads with keep == 0 become the degenerate triangle (0,0,0), which the rasteriser discards. */
void orc_sv_build_volumes(const float* xyz, const float* nrm, int V, const int32_t* idx, int T, const float light[3], int infinity,
                          const uint8_t* keep, int caps, float* prism_xyz, int32_t* vol_idx) {
  int32_t* side = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T > 0 ? T : 1) * 18);
  orc_sv_build_prisms(xyz, nrm, V, idx, T, light, infinity, prism_xyz, side);
  const int per = caps ? 8 : 6;
  for (int t = 0; t < T; t++) {
    int32_t* o = vol_idx + (size_t)t * per * 3;
    for (int k = 0; k < 18; k++) o[k] = side[(size_t)t * 18 + k];
    if (keep)
      for (int q = 0; q < 3; q++)
        if (!keep[3 * (size_t)t + q]) for (int k = 0; k < 6; k++) o[q * 6 + k] = 0;
    if (caps) {
      const int ordA = side[(size_t)t * 18] == t * 6 + 1;              /* ORD_A starts 1,0,3 ; ORD_B 4,3,0 */
      const int nearA[3] = {0, 1, 2}, nearB[3] = {0, 2, 1}, farA[3] = {4, 3, 5}, farB[3] = {3, 4, 5};
      for (int k = 0; k < 3; k++) { o[18 + k] = t * 6 + (ordA ? nearA[k] : nearB[k]); o[21 + k] = t * 6 + (ordA ? farA[k] : farB[k]); }
    }
  }
  free(side);
}

/* depth-pass (zfail == 0: the reference's stencil ops, ShadowVolumes/src/main.cpp:160-172) or depth-fail counting (zfail == 1: back
 * faces +1, front faces -1 on fragments that FAIL the depth test; far plane not clipped, depths saturate at 1).  `per` = triangles
 * per source triangle in prism_idx; with per == 8 triangles 6 and 7 of each group are caps, whose depth test is strict (a cap
 * fragment coplanar with the visible surface - the surface's own triangle - counts as failing). */
int orc_sv_count_ex(const float* prism_xyz, int PV, const int32_t* prism_idx, int PT, const float mvp[16], int W, int H,
                    const float* scene_depth, int depth_func, int zfail, int per, int32_t* count, uint8_t* stencil) {
  (void)PV;
  int64_t n;
  g_no_far_clip = zfail ? 1 : 0;
  SubTri* rec = build_records(prism_xyz, prism_idx, PT, mvp, W, H, 0, 0.0f, 0.0f, &n);
  g_no_far_clip = 0;
  if (!rec) return -1;
  memset(count, 0, sizeof(int32_t) * (size_t)W * H);
  int bands = H < 64 ? 1 : 64;
#pragma omp parallel for schedule(dynamic, 1)
  for (int bnd = 0; bnd < bands; bnd++) {
    int r0 = (int)((int64_t)H * bnd / bands), r1 = (int)((int64_t)H * (bnd + 1) / bands) - 1;
    for (int64_t k = 0; k < n; k++) {
      const SubTri* s = &rec[k];
      int y0 = s->py0 > r0 ? s->py0 : r0, y1 = s->py1 < r1 ? s->py1 : r1;
      int inc = s->front ? 1 : -1;            /* GL_FRONT zpass INCR_WRAP / GL_BACK zpass DECR_WRAP */
      int is_cap = per == 8 && ((s->prim >> 3) % 8) >= 6;
      for (int j = y0; j <= y1; j++)
        for (int i = s->px0; i <= s->px1; i++) {
          int64_t E[3];
          if (!cover(s, i, j, E)) continue;
          float z = frag_z(s, E);
          float d = scene_depth[(size_t)j * W + i];
          int pass = (depth_func == ORC_DEPTH_LESS || is_cap) ? (z < d) : (z <= d);
          if (!zfail) { if (pass) count[(size_t)j * W + i] += inc; }
          else if (!pass) count[(size_t)j * W + i] -= inc;
        }
    }
  }
  if (stencil)
    for (size_t i = 0; i < (size_t)W * H; i++) stencil[i] = (uint8_t)((uint32_t)count[i] & 255u);
  free(rec);
  return 0;
}

int orc_sv_count(const float* prism_xyz, int PV, const int32_t* prism_idx, int PT, const float mvp[16], int W, int H,
                 const float* scene_depth, int depth_func, int32_t* count, uint8_t* stencil) {
  return orc_sv_count_ex(prism_xyz, PV, prism_idx, PT, mvp, W, H, scene_depth, depth_func, 0, 6, count, stencil);
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
