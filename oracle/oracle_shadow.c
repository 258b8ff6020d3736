/*
 * oracle_shadow.c — CPU restatement of the reference's matrix set-up and per-pixel shadow passes.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Each function cites the reference lines it follows
 * (paths relative to /root/reference).  fp32, source order, no FMA contraction.
 */
#include "oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ================================ matrices ==================================================== */
void orc_mat4_identity(float m[16]) { memset(m, 0, 64); m[0] = m[5] = m[10] = m[15] = 1.0f; }

/* glm/core/type_mat4x4.inl:757-779: Result[c] = A0*B[c][0] + A1*B[c][1] + A2*B[c][2] + A3*B[c][3] */
void orc_mat4_mul(const float a[16], const float b[16], float out[16]) {
  float r[16];
  for (int c = 0; c < 4; c++)
    for (int k = 0; k < 4; k++)
      r[c * 4 + k] = ((a[0 + k] * b[c * 4 + 0] + a[4 + k] * b[c * 4 + 1]) + a[8 + k] * b[c * 4 + 2]) + a[12 + k] * b[c * 4 + 3];
  memcpy(out, r, 64);
}

static float radians_f(float deg) {                       /* glm/core/func_trigonometric.inl:35-44 */
  const float pi = (float)3.1415926535897932384626433832795;
  return deg * (pi / 180.0f);
}

/* glm/gtc/matrix_transform.inl:223-244 */
void orc_perspective(float fovy, float aspect, float zn, float zf, float out[16]) {
  float range = tanf(radians_f(fovy / 2.0f)) * zn;
  float left = -range * aspect, right = range * aspect, bottom = -range, top = range;
  memset(out, 0, 64);
  out[0] = (2.0f * zn) / (right - left);
  out[5] = (2.0f * zn) / (top - bottom);
  out[10] = -(zf + zn) / (zf - zn);
  out[11] = -1.0f;
  out[14] = -(2.0f * zf * zn) / (zf - zn);
}

static void normalize3(const float v[3], float o[3]) {    /* glm/core/func_geometric.inl:239-248 */
  float sqr = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  float inv = 1.0f / sqrtf(sqr);
  o[0] = v[0] * inv; o[1] = v[1] * inv; o[2] = v[2] * inv;
}
static void cross3(const float x[3], const float y[3], float o[3]) {   /* func_geometric.inl cross */
  float r0 = x[1] * y[2] - y[1] * x[2];
  float r1 = x[2] * y[0] - y[2] * x[0];
  float r2 = x[0] * y[1] - y[0] * x[1];
  o[0] = r0; o[1] = r1; o[2] = r2;
}

/* glm/gtc/matrix_transform.inl:32-43 */
void orc_translate(const float m[16], const float v[3], float out[16]) {
  float r[16];
  memcpy(r, m, 64);
  for (int k = 0; k < 4; k++) r[12 + k] = ((m[0 + k] * v[0] + m[4 + k] * v[1]) + m[8 + k] * v[2]) + m[12 + k];
  memcpy(out, r, 64);
}

/* glm/gtc/matrix_transform.inl:383-411 */
void orc_look_at(const float eye[3], const float at[3], const float up[3], float out[16]) {
  float d[3] = {at[0] - eye[0], at[1] - eye[1], at[2] - eye[2]};
  float f[3], u[3], s[3], c[3];
  normalize3(d, f);
  normalize3(up, u);
  cross3(f, u, c);
  normalize3(c, s);
  cross3(s, f, u);
  float r[16];
  orc_mat4_identity(r);
  r[0] = s[0]; r[4] = s[1]; r[8] = s[2];
  r[1] = u[0]; r[5] = u[1]; r[9] = u[2];
  r[2] = -f[0]; r[6] = -f[1]; r[10] = -f[2];
  float ne[3] = {-eye[0], -eye[1], -eye[2]};
  orc_translate(r, ne, out);
}

/* glm/gtc/matrix_transform.inl:44-78 */
void orc_rotate(const float m[16], float angle, const float v[3], float out[16]) {
  float a = radians_f(angle);
  float c = cosf(a), s = sinf(a);
  float axis[3], temp[3];
  normalize3(v, axis);
  for (int k = 0; k < 3; k++) temp[k] = (1.0f - c) * axis[k];
  float R[3][3];
  R[0][0] = c + temp[0] * axis[0];
  R[0][1] = 0 + temp[0] * axis[1] + s * axis[2];
  R[0][2] = 0 + temp[0] * axis[2] - s * axis[1];
  R[1][0] = 0 + temp[1] * axis[0] - s * axis[2];
  R[1][1] = c + temp[1] * axis[1];
  R[1][2] = 0 + temp[1] * axis[2] + s * axis[0];
  R[2][0] = 0 + temp[2] * axis[0] + s * axis[1];
  R[2][1] = 0 + temp[2] * axis[1] - s * axis[0];
  R[2][2] = c + temp[2] * axis[2];
  float r[16];
  for (int col = 0; col < 3; col++)
    for (int k = 0; k < 4; k++)
      r[col * 4 + k] = (m[0 + k] * R[col][0] + m[4 + k] * R[col][1]) + m[8 + k] * R[col][2];
  for (int k = 0; k < 4; k++) r[12 + k] = m[12 + k];
  memcpy(out, r, 64);
}

/* ShadowMapping/src/Viewers/MyGLGeometryViewer.cpp:139-145 */
void orc_bias_mul(const float light_mvp[16], float out[16]) {
  float bias[16] = {0.5f, 0, 0, 0, 0, 0.5f, 0, 0, 0, 0, 0.5f, 0, 0.5f, 0.5f, 0.5f, 1.0f};
  orc_mat4_mul(bias, light_mvp, out);
}

/* glm/gtc/matrix_inverse.inl:77-100 on mat3(mv) (MyGLGeometryViewer.cpp:115) */
void orc_normal_matrix(const float mv[16], float o[9]) {
#define M(c, r) mv[(c) * 4 + (r)]
  float det = +M(0, 0) * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) - M(0, 1) * (M(1, 0) * M(2, 2) - M(1, 2) * M(2, 0)) +
              M(0, 2) * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
  float inv[9];
  inv[0 * 3 + 0] = +(M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2));
  inv[0 * 3 + 1] = -(M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2));
  inv[0 * 3 + 2] = +(M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1));
  inv[1 * 3 + 0] = -(M(0, 1) * M(2, 2) - M(2, 1) * M(0, 2));
  inv[1 * 3 + 1] = +(M(0, 0) * M(2, 2) - M(2, 0) * M(0, 2));
  inv[1 * 3 + 2] = -(M(0, 0) * M(2, 1) - M(2, 0) * M(0, 1));
  inv[2 * 3 + 0] = +(M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2));
  inv[2 * 3 + 1] = -(M(0, 0) * M(1, 2) - M(1, 0) * M(0, 2));
  inv[2 * 3 + 2] = +(M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1));
#undef M
  for (int k = 0; k < 9; k++) o[k] = inv[k] / det;
}

/* ShadowMapping/src/main.cpp:283: lightEye = mat3(rotate(180, (0,1,0))) * lightEye */
void orc_rotate_light_180(const float eye[3], float out[3]) {
  float id[16], r[16];
  orc_mat4_identity(id);
  const float ax[3] = {0, 1, 0};
  orc_rotate(id, 180.0f, ax, r);
  for (int k = 0; k < 3; k++) out[k] = (r[0 + k] * eye[0] + r[4 + k] * eye[1]) + r[8 + k] * eye[2];   /* mat3*vec3 */
}

/* SoftShadowMapping/src/Scene/LightSource/UniformSampledLightSource.cpp:27-38 */
void orc_uniform_light_sample(const float p[3], int size, int n_lights, int index, float out[3]) {
  float halfSize = (float)((float)size / 2.0);
  float factor = sqrtf((float)n_lights);
  float sampleSize = (float)((factor - 1) / 2.0);
  out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
  out[0] += (((index % (int)factor) - sampleSize) / sampleSize) * halfSize;
  out[1] += (((int)(index / factor) - sampleSize) / sampleSize) * halfSize;
}

/* Shadow.frag:93-98 / NonConservativeSMSR.frag:313-319: `for(float w=-offset; w<offset; w+=stepSize)` */
int orc_pcf_offsets(int kernel_order, int penumbra_size, int inclusive, float* out, int cap) {
  float offset = (float)penumbra_size;
  float stepSize = 2 * offset / (float)kernel_order;
  int n = 0;
  if (!(stepSize > 0.0f)) return -1;
  for (float w = -offset; inclusive ? (w <= offset) : (w < offset); w += stepSize) {
    if (n >= cap) return -1;
    out[n++] = w;
  }
  return n;
}

/* ================================ per-pixel passes ============================================ */
typedef struct { const float* d; int w, h; float fw, fh; } Smap;
typedef struct { float x, y, z, w; } V4;

/* A.3: GL_NEAREST + CLAMP_TO_BORDER(0) on a depth texture: MyGLTextureViewer.cpp:3-28 */
static inline float sm_fetch(const Smap* s, float u, float v) {
  float fu = u * s->fw, fv = v * s->fh;
  float fi = floorf(fu), fj = floorf(fv);
  if (!(fi >= 0.0f && fi < s->fw && fj >= 0.0f && fj < s->fh)) return 0.0f;
  return s->d[(size_t)(int)fj * s->w + (int)fi];
}
static inline float glsl_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline float glsl_max(float a, float b) { return a < b ? b : a; }
static inline float glsl_min(float a, float b) { return b < a ? b : a; }
static inline float glsl_fract(float x) { return x - floorf(x); }

static inline V4 mat4_mul_v4(const float* m, V4 v) {
  V4 r;
  r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
  r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
  r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
  r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
  return r;
}

/* Shadow.frag:222-238 (identical copies in every shadow shader) */
static float pre_evaluation(const orc_camera* cam, float si, V4 vertex, V4 normal) {
  V4 ev = mat4_mul_v4(cam->mv, vertex);
  const float* nm = cam->normal_matrix;
  float n[3], nn[3], L[3], d[3];
  for (int k = 0; k < 3; k++) n[k] = (nm[0 + k] * normal.x + nm[3 + k] * normal.y) + nm[6 + k] * normal.z;
  {
    float sqr = (n[0] * n[0] + n[1] * n[1]) + n[2] * n[2];
    float inv = 1.0f / sqrtf(sqr);
    nn[0] = n[0] * inv; nn[1] = n[1] * inv; nn[2] = n[2] * inv;
  }
  d[0] = cam->light_pos[0] - ev.x; d[1] = cam->light_pos[1] - ev.y; d[2] = cam->light_pos[2] - ev.z;
  {
    float sqr = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    float inv = 1.0f / sqrtf(sqr);
    L[0] = d[0] * inv; L[1] = d[1] * inv; L[2] = d[2] * inv;
  }
  if (!(normal.w != 0.0f)) { nn[0] *= -1.0f; nn[1] *= -1.0f; nn[2] *= -1.0f; }
  float dt = (nn[0] * L[0] + nn[1] * L[1]) + nn[2] * L[2];
  if (glsl_max(dt, 0.0f) == 0.0f) return si;
  return 1.0f;
}

/* ---- Shadow.frag:86-116 ---- */
/* ---- Shadow.frag:41-84: cubic() weights and textureBicubic() on the NEAREST depth texture (4 fetches), .z component ---- */
static void cubic4(float v, float o[4]) {
  float n[4] = {1.0f - v, 2.0f - v, 3.0f - v, 4.0f - v}, s[4];
  for (int k = 0; k < 4; k++) s[k] = n[k] * n[k] * n[k];
  float x = s[0];
  float y = s[1] - 4.0f * s[0];
  float z = s[2] - 4.0f * s[1] + 6.0f * s[0];
  float w = 6.0f - x - y - z;
  const float sixth = 1.0f / 6.0f;
  o[0] = x * sixth; o[1] = y * sixth; o[2] = z * sixth; o[3] = w * sixth;
}
static float texture_bicubic_z(const Smap* sm, float u, float v) {
  float invx = 1.0f / sm->fw, invy = 1.0f / sm->fh;
  float tx = u * sm->fw - 0.5f, ty = v * sm->fh - 0.5f;
  float fx = glsl_fract(tx), fy = glsl_fract(ty);
  tx -= fx; ty -= fy;
  float xc[4], yc[4];
  cubic4(fx, xc); cubic4(fy, yc);
  float c0 = tx + -0.5f, c1 = tx + 1.5f, c2 = ty + -0.5f, c3 = ty + 1.5f;
  float s0 = xc[0] + xc[1], s1 = xc[2] + xc[3], s2 = yc[0] + yc[1], s3 = yc[2] + yc[3];
  float o0 = c0 + xc[1] / s0, o1 = c1 + xc[3] / s1, o2 = c2 + yc[1] / s2, o3 = c3 + yc[3] / s3;
  o0 *= invx; o1 *= invx; o2 *= invy; o3 *= invy;
  float sample0 = sm_fetch(sm, o0, o2), sample1 = sm_fetch(sm, o1, o2), sample2 = sm_fetch(sm, o0, o3), sample3 = sm_fetch(sm, o1, o3);
  float sx = s0 / (s0 + s1), sy = s2 / (s2 + s3);
  return glsl_mix(glsl_mix(sample3, sample2, sx), glsl_mix(sample1, sample0, sx), sy);
}

/* Shadow.frag:86-116 with tricubicPCF == 1, bilinearPCF == 0 (:101-102) */
static float pcf_tricubic(const orc_params* p, const Smap* s, V4 c) {
  float incrWidth = 1.0f / (float)p->shadow_map_width;
  float incrHeight = 1.0f / (float)p->shadow_map_height;
  float illuminationCount = 0;
  float offset = (float)p->penumbra_size;
  float stepSize = 2 * offset / (float)p->kernel_order;
  int count = 0;
  if (!(stepSize > 0.0f)) return 1.0f;
  for (float w = -offset; w < offset; w += stepSize)
    for (float h = -offset; h < offset; h += stepSize) {
      float dfl = texture_bicubic_z(s, c.x + w * incrWidth, c.y + h * incrHeight);
      if (c.z <= dfl) illuminationCount++;
      else illuminationCount += p->shadow_intensity;
      count++;
    }
  return illuminationCount / (float)count;
}

static float pcf(const orc_params* p, const Smap* s, V4 c) {
  float incrWidth = 1.0f / (float)p->shadow_map_width;
  float incrHeight = 1.0f / (float)p->shadow_map_height;
  float illuminationCount = 0;
  float offset = (float)p->penumbra_size;
  float stepSize = 2 * offset / (float)p->kernel_order;
  int count = 0;
  if (!(stepSize > 0.0f)) return 1.0f;     /* guard: the reference would loop forever */
  for (float w = -offset; w < offset; w += stepSize)
    for (float h = -offset; h < offset; h += stepSize) {
      float dfl = sm_fetch(s, c.x + w * incrWidth, c.y + h * incrHeight);
      if (c.z <= dfl) illuminationCount++;
      else illuminationCount += p->shadow_intensity;
      count++;
    }
  return illuminationCount / (float)count;
}

/* ---- PlausibleSoftShadow.frag:166-194, 365-374, 376-398, 556-563 ---- */
static float pcss_penumbra(const orc_params* p, const Smap* s, V4 c) {
  float averageDepth = 0.0f;
  int numberOfBlockers = 0;
  float blockerSearchWidth;
  if ((float)p->shadow_map_width <= 1024.0f) blockerSearchWidth = (float)p->light_source_radius / (float)p->shadow_map_width;
  else blockerSearchWidth = (float)p->light_source_radius / 1024.0f;
  float filterWidth = ((float)p->blocker_search_size - 1.0f) * 0.5f;
  for (int h = (int)(-filterWidth); (float)h <= filterWidth; h++)
    for (int w = (int)(-filterWidth); (float)w <= filterWidth; w++) {
      float u = c.x + ((float)w * blockerSearchWidth) / filterWidth;
      float v = c.y + ((float)h * blockerSearchWidth) / filterWidth;
      float dfl = sm_fetch(s, u, v);
      if (c.z > dfl) { averageDepth += dfl; numberOfBlockers++; }
    }
  if (numberOfBlockers == 0) averageDepth = 1.0f;
  else averageDepth = averageDepth / (float)numberOfBlockers;
  /* computePenumbraWidth */
  float penumbraWidth;
  if (averageDepth < 0.99f) penumbraWidth = 0.0f;
  else {
    float pw = ((c.z - averageDepth) / averageDepth) * (float)p->light_source_radius;
    penumbraWidth = ((float)p->z_near * pw) / c.z;
  }
  return penumbraWidth;
}
static float pcss(const orc_params* p, const Smap* s, V4 c) {
  float penumbraWidth = pcss_penumbra(p, s, c);
  /* PCF */
  float illuminationCount = 0.0f;
  float stepSize = 2.0f * penumbraWidth / (float)p->kernel_size;
  float fw2 = ((float)p->kernel_size - 1.0f) * 0.5f;
  if (stepSize <= 0.0f || stepSize >= 1.0f) return 1.0f;
  for (int h = (int)(-fw2); (float)h <= fw2; h++)
    for (int w = (int)(-fw2); (float)w <= fw2; w++) {
      float u = c.x + ((float)w * penumbraWidth) / fw2;
      float v = c.y + ((float)h * penumbraWidth) / fw2;
      float dfl = sm_fetch(s, u, v);
      if (c.z <= dfl) illuminationCount++;
      else illuminationCount += p->shadow_intensity;
    }
  return illuminationCount / (float)(p->kernel_size * p->kernel_size);
}

/* ================================ RBSM, non-conservative ====================================== */
typedef struct {
  const orc_params* p; const Smap* s;
  float sx, sy;        /* shadowMapStep: MyGLGeometryViewer.cpp:238 (double 1.0/int narrowed) */
  float newDepth;      /* NonConservativeSMSR.frag:21 global                                   */
  int filtered;        /* FilteredRBSM.frag variant                                            */
} Rb;

/* NonConservativeSMSR.frag:23-54 with both break flags false */
static void nc_getdisc4(Rb* r, V4 c, float dir[4]) {
  c.x -= r->sx;
  dir[0] = (c.z <= sm_fetch(r->s, c.x, c.y)) ? 1.0f : 0.0f;
  c.x += 2.0f * r->sx;
  dir[1] = (c.z <= sm_fetch(r->s, c.x, c.y)) ? 1.0f : 0.0f;
  c.x -= r->sx;
  c.y += r->sy;
  dir[2] = (c.z <= sm_fetch(r->s, c.x, c.y)) ? 1.0f : 0.0f;
  c.y -= 2.0f * r->sy;
  dir[3] = (c.z <= sm_fetch(r->s, c.x, c.y)) ? 1.0f : 0.0f;
}

/* :56-94 */
static int nc_getdisc_f(Rb* r, V4 c, float dx, float dy, float discType) {
  if (dx == 0.0f) {
    c.x -= r->sx;
    float left = (c.z <= sm_fetch(r->s, c.x, c.y)) ? 1.0f : 0.0f;
    if (fabsf(left - discType) == 0.0f) return 1;
    c.x += 2.0f * r->sx;
    float right = (c.z <= sm_fetch(r->s, c.x, c.y)) ? 1.0f : 0.0f;
    if (fabsf(right - discType) == 0.0f) return 1;
    c.x -= r->sx;
  }
  if (dy == 0.0f) {
    c.y += r->sy;
    float bottom = (c.z <= sm_fetch(r->s, c.x, c.y)) ? 1.0f : 0.0f;
    if (fabsf(bottom - discType) == 0.0f) return 1;
    c.y -= 2.0f * r->sy;
    float top = (c.z <= sm_fetch(r->s, c.x, c.y)) ? 1.0f : 0.0f;
    if (fabsf(top - discType) == 0.0f) return 1;
  }
  return 0;
}

/* :96-178 ; disc = (r,g,b,a) */
static int nc_getdisc_v(Rb* r, V4 c, float dx, float dy, const float disc[4]) {
  float thr = r->p->depth_threshold;
  V4 rel = c;
  r->newDepth = c.z;
#define NC_SIDE(UX, UY)                                                                   \
  {                                                                                       \
    float dfl = sm_fetch(r->s, (UX), (UY));                                               \
    if (disc[2] == 1.0f) {                                                                \
      if (fabsf(c.z - dfl) < thr) { c.z -= thr; r->newDepth = c.z; }                      \
    }                                                                                     \
    float side = (c.z <= dfl) ? 1.0f : 0.0f;                                              \
    if (fabsf(side - disc[2]) == 0.0f) return 1;                                          \
  }
  if (dx == 0.0f) {
    if (disc[0] == 0.5f || disc[0] == 0.75f) { rel.x = c.x - r->sx; NC_SIDE(rel.x, rel.y) }
    if (disc[0] == 0.75f || disc[0] == 0.25f) { rel.x = c.x + r->sx; NC_SIDE(rel.x, rel.y) }
  }
  if (dy == 0.0f) {
    if (disc[1] == 0.5f || disc[1] == 0.75f) { rel.y = c.y + r->sy; NC_SIDE(rel.x, rel.y) }
    if (disc[1] == 0.75f || disc[1] == 0.25f) { rel.y = c.y - r->sy; NC_SIDE(rel.x, rel.y) }
  }
#undef NC_SIDE
  return 0;
}

/* :180-232 */
static float nc_disc_length(Rb* r, const float disc[4], V4 lightCoord, float dx, float dy, float subCoord) {
  float thr = r->p->depth_threshold;
  V4 c = lightCoord;
  float foundEdgeEnd = 0.0f;
  float dist = 0.0f;
  float stx = dx * r->sx, sty = dy * r->sy;
  c.x += stx; c.y += sty;
  for (int it = 0; it < r->p->max_search; it++) {
    float dfl = sm_fetch(r->s, c.x, c.y);
    if (disc[2] == 0.0f)
      if (fabsf(c.z - dfl) < thr) c.z -= thr;
    float center = (c.z <= dfl) ? 1.0f : 0.0f;
    if (fabsf(center - disc[2]) == 0.0f) {
      int hasDisc = nc_getdisc_f(r, c, 0.0f, 0.0f, disc[2]);
      foundEdgeEnd = hasDisc ? 1.0f : 0.0f;
      break;
    } else {
      int hasDisc = nc_getdisc_v(r, c, dx, dy, disc);
      if (!hasDisc) break;
    }
    dist++;
    c.x += stx; c.y += sty;
    if (disc[2] == 1.0f) c.z = r->newDepth;
  }
  return glsl_mix(-(dist + (1.0f - subCoord)), dist + (1.0f - subCoord), foundEdgeEnd);
}

/* :234-243 (FilteredRBSM.frag:241 uses T instead of max(T, 2*shadow-1)) */
static float nc_rel_pos(Rb* r, float ax, float ay, float shadow) {
  float T = 1;
  if (ax < 0.0f && ay < 0.0f) T = 0;
  if (ax > 0.0f && ay > 0.0f) T = -2;
  float edgeLength = glsl_min(fabsf(ax) + fabsf(ay), (float)r->p->max_search);
  float lead = r->filtered ? T : glsl_max(T, 2 * shadow - 1);
  return (lead * fabsf(glsl_max(T * ax, T * ay))) / edgeLength;
}

/* :266-274 / FilteredRBSM.frag:266-272 */
static float nc_revectorize(Rb* r, float rx, float ry, float shadow) {
  if (r->filtered) {
    if (rx * ry < 0) return (1.0f - shadow) + (2 * shadow - 1) * glsl_max(rx, ry);
    else if (rx * ry == 0.0f) return shadow;
    else {
      float v = (1.0f - shadow) + (2 * shadow - 1) * (rx + ry);
      return glsl_min(glsl_max(v, 0.0f), 1.0f);
    }
  }
  if ((rx * ry == 2 * shadow) ||
      ((fabsf(rx) * fabsf(ry) > 0) && ((1.0f - shadow) + (2 * shadow - 1) * (fabsf(rx) + fabsf(ry)) < 0.5f)))
    return 0.0f;
  return 1.0f;
}

/* :276-286 */
static void nc_compute_disc(Rb* r, V4 c, float dfl, float disc[4]) {
  float center = (c.z <= dfl) ? 1.0f : 0.0f;
  float discType = 1.0f - center;
  float dir[4];
  nc_getdisc4(r, c, dir);
  float d0 = fabsf(dir[0] - center), d1 = fabsf(dir[1] - center), d2 = fabsf(dir[2] - center), d3 = fabsf(dir[3] - center);
  disc[0] = (2.0f * d0 + d1) / 4.0f;   /* (2*disc.xz + disc.yw)/4 */
  disc[1] = (2.0f * d2 + d3) / 4.0f;
  disc[2] = discType;
  disc[3] = 1.0f;
}

/* orientateDS :245-254 + estimateRelativePosition :256-264 */
static void nc_rel(Rb* r, V4 c, const float disc[4], float subx, float suby, float shadow, float* rx, float* ry) {
  float left = nc_disc_length(r, disc, c, -1, 0, (1.0f - subx));
  float right = nc_disc_length(r, disc, c, 1, 0, subx);
  float down = nc_disc_length(r, disc, c, 0, -1, (1.0f - suby));
  float up = nc_disc_length(r, disc, c, 0, 1, suby);
  *rx = nc_rel_pos(r, left, right, shadow);
  *ry = nc_rel_pos(r, down, up, shadow);
}

/* :288-304 */
static float nc_smsr(Rb* r, V4 c) {
  float si = r->p->shadow_intensity;
  float dfl = sm_fetch(r->s, c.x, c.y);
  float disc[4];
  nc_compute_disc(r, c, dfl, disc);
  float subx = glsl_fract(c.x * (float)r->p->shadow_map_width), suby = glsl_fract(c.y * (float)r->p->shadow_map_height);
  float shadow = (c.z <= dfl) ? 1.0f : 0.0f;
  if (disc[0] > 0.0f || disc[1] > 0.0f) {
    if (disc[0] == 0.75f && disc[1] == 0.75f) return glsl_mix(1.0f - shadow, 1.0f, si);
    float rx, ry;
    nc_rel(r, c, disc, subx, suby, shadow, &rx, &ry);
    return glsl_mix(nc_revectorize(r, rx, ry, shadow), 1.0f, si);
  }
  return glsl_mix(shadow, 1.0f, si);
}

/* :306-349 */
static float nc_rpcf(Rb* r, V4 c) {
  const orc_params* p = r->p;
  float si = p->shadow_intensity;
  float incrWidth = 1.0f / (float)p->shadow_map_width, incrHeight = 1.0f / (float)p->shadow_map_height;
  float illuminationCount = 0.0f;
  float offset = (float)p->penumbra_size;
  float stepSize = 2 * offset / (float)p->kernel_order;
  int count = 0;
  if (!(stepSize > 0.0f)) return 1.0f;
  for (float w = -offset; w <= offset; w += stepSize)
    for (float h = -offset; h <= offset; h += stepSize) {
      V4 sc = {c.x + w * incrWidth, c.y + h * incrHeight, c.z, c.w};
      float dfl = sm_fetch(r->s, sc.x, sc.y);
      float shadow = (sc.z <= dfl) ? 1.0f : 0.0f;
      float disc[4];
      nc_compute_disc(r, sc, dfl, disc);
      if (disc[0] > 0.0f || disc[1] > 0.0f) {
        float subx = glsl_fract(sc.x * (float)p->shadow_map_width), suby = glsl_fract(sc.y * (float)p->shadow_map_height);
        float rx, ry;
        nc_rel(r, sc, disc, subx, suby, shadow, &rx, &ry);
        illuminationCount += glsl_mix(nc_revectorize(r, rx, ry, shadow), 1.0f, si);
      } else {
        shadow = glsl_mix(shadow, 1.0f, si);
        illuminationCount += shadow;
      }
      count++;
    }
  return illuminationCount / (float)count;
}

/* ================================ RBSM, conservative ========================================== */
/* ConservativeSMSR.frag:23-48 ; returns abs(dir - 1) */
static void cs_disc(Rb* r, V4 c, float d[4]) {
  float dir[4];
  nc_getdisc4(r, c, dir);       /* identical fetch sequence */
  for (int k = 0; k < 4; k++) d[k] = fabsf(dir[k] - 1.0f);
}

/* :50-86 */
static float cs_rel_distance(Rb* r, V4 sc, float dx, float dy, float cc) {
  float thr = r->p->depth_threshold;
  V4 t = sc;
  float foundSilhouetteEnd = 0.0f;
  float distance = 0.0f;
  float stx = dx * r->sx, sty = dy * r->sy;
  t.x += stx; t.y += sty;
  for (int it = 0; it < r->p->max_search; it++) {
    float dfl = sm_fetch(r->s, t.x, t.y);
    if (fabsf(t.z - dfl) < thr) t.z -= thr;
    float center = (t.z <= dfl) ? 1.0f : 0.0f;
    int isCenterUmbra = !(center != 0.0f);
    if (isCenterUmbra) { foundSilhouetteEnd = 1.0f; break; }
    else {
      float d[4];
      cs_disc(r, t, d);
      if ((d[0] + d[1] + d[2] + d[3]) == 0.0f) break;
    }
    distance++;
    t.x += stx; t.y += sty;
  }
  distance = distance + (1.0f - cc);
  return glsl_mix(-distance, distance, foundSilhouetteEnd);
}

/* :99-108 */
static float cs_norm(Rb* r, float ax, float ay) {
  float T = 1;
  if (ax < 0.0f && ay < 0.0f) T = 0;
  if (ax > 0.0f && ay > 0.0f) T = -2;
  float length = glsl_min(fabsf(ax) + fabsf(ay), (float)r->p->max_search);
  return fabsf(glsl_max(T * ax, T * ay)) / length;
}

/* :88-97, :110-126 */
static float cs_revec(Rb* r, V4 sc, float cx, float cy) {
  float dl = cs_rel_distance(r, sc, -1, 0, (1.0f - cx));
  float dr = cs_rel_distance(r, sc, 1, 0, cx);
  float db = cs_rel_distance(r, sc, 0, -1, (1.0f - cy));
  float dt = cs_rel_distance(r, sc, 0, 1, cy);
  float rx = cs_norm(r, dl, dr), ry = cs_norm(r, db, dt);
  if ((rx * ry > 0) && (1.0f - rx > ry)) return r->p->shadow_intensity;
  return 1.0f;
}

/* :128-144 */
static float cs_smsr(Rb* r, V4 c) {
  float si = r->p->shadow_intensity;
  float dfl = sm_fetch(r->s, c.x, c.y);
  float shadow = (c.z <= dfl) ? 1.0f : 0.0f;
  if (shadow == 0.0f) return si;
  float d[4];
  cs_disc(r, c, d);
  if ((d[0] + d[1] + d[2] + d[3]) == 0.0f) return 1.0f;
  else if ((d[0] + d[1]) == 2.0f || (d[2] + d[3]) == 2.0f) return si;
  float cx = glsl_fract(c.x * (float)r->p->shadow_map_width), cy = glsl_fract(c.y * (float)r->p->shadow_map_height);
  return cs_revec(r, c, cx, cy);
}

/* :146-198 */
static float cs_rpcf(Rb* r, V4 c) {
  const orc_params* p = r->p;
  float si = p->shadow_intensity;
  float incrWidth = 1.0f / (float)p->shadow_map_width, incrHeight = 1.0f / (float)p->shadow_map_height;
  float illuminationCount = 0.0f;
  float offset = (float)p->penumbra_size;
  float stepSize = 2 * offset / (float)p->kernel_order;
  int count = 0;
  if (!(stepSize > 0.0f)) return 1.0f;
  for (float w = -offset; w <= offset; w += stepSize)
    for (float h = -offset; h <= offset; h += stepSize) {
      float dfl = sm_fetch(r->s, c.x + w * incrWidth, c.y + h * incrHeight);
      float shadow = (c.z <= dfl) ? 1.0f : si;
      if (shadow == 1.0f) {
        V4 sc = {c.x + w * incrWidth, c.y + h * incrHeight, c.z, c.w};
        float d[4];
        cs_disc(r, sc, d);
        if (d[0] == 0.0f && d[1] == 0.0f) illuminationCount++;
        else {
          float subx = glsl_fract(sc.x * (float)p->shadow_map_width), suby = glsl_fract(sc.y * (float)p->shadow_map_height);
          illuminationCount += cs_revec(r, sc, subx, suby);
        }
      } else illuminationCount += si;
      count++;
    }
  return illuminationCount / (float)count;
}

#include "oracle_rbssm_impl.h"   /* RBSSM.frag: ss_rbssm */

/* ================================ drivers ===================================================== */
void orc_visibility(const orc_params* p, const orc_camera* cam, const float light_mvp_b[16], const float* pos4,
                    const float* nrm4, int W, int H, const float* shadow_map, float* vis) {
  Smap s = {shadow_map, p->shadow_map_width, p->shadow_map_height, (float)p->shadow_map_width, (float)p->shadow_map_height};
  int x0 = p->rect_x0, y0 = p->rect_y0, x1 = p->rect_x1, y1 = p->rect_y1;
  if (x1 <= x0 || y1 <= y0) { x0 = 0; y0 = 0; x1 = W; y1 = H; }
  int tech = p->technique;
#pragma omp parallel for schedule(dynamic, 4)
  for (int j = y0; j < y1; j++) {
    Rb r;
    r.p = p; r.s = &s;
    r.sx = (float)(1.0 / p->shadow_map_width); r.sy = (float)(1.0 / p->shadow_map_height);
    r.newDepth = 0.0f; r.filtered = (tech == ORC_TECH_RSMSS);
    for (int i = x0; i < x1; i++) {
      size_t o = (size_t)j * W + i;
      V4 vertex = {pos4[4 * o], pos4[4 * o + 1], pos4[4 * o + 2], pos4[4 * o + 3]};
      if (vertex.x == 0.0f) continue;                      /* discard: Shadow.frag:244 */
      V4 normal = {nrm4[4 * o], nrm4[4 * o + 1], nrm4[4 * o + 2], nrm4[4 * o + 3]};
      V4 sc = mat4_mul_v4(light_mvp_b, vertex);
      V4 c = {sc.x / sc.w, sc.y / sc.w, sc.z / sc.w, sc.w / sc.w};
      float shadow = pre_evaluation(cam, p->shadow_intensity, vertex, normal);
      if (tech == ORC_TECH_HARD || tech == ORC_TECH_PCF || tech == ORC_TECH_PCSS || tech == ORC_TECH_RBSSM || tech == ORC_TECH_PCF_TRICUBIC) {
        if (sc.w > 0.0f && shadow == 1.0f) {               /* Shadow.frag:251 / PlausibleSoftShadow.frag:616 / RBSSM.frag:1372 */
          if (tech == ORC_TECH_HARD) shadow = (c.z <= sm_fetch(&s, c.x, c.y)) ? 1.0f : p->shadow_intensity;
          else if (tech == ORC_TECH_PCF) shadow = pcf(p, &s, c);
          else if (tech == ORC_TECH_PCF_TRICUBIC) shadow = pcf_tricubic(p, &s, c);
          else if (tech == ORC_TECH_RBSSM) shadow = ss_rbssm(&r, c);
          else shadow = pcss(p, &s, c);
        }
      } else if (shadow == 1.0f) {                         /* NonConservativeSMSR.frag:379 */
        switch (tech) {
          case ORC_TECH_RBSM_NONCONS: shadow = nc_smsr(&r, c); break;
          case ORC_TECH_RBSM_CONS: shadow = cs_smsr(&r, c); break;
          case ORC_TECH_RPCF_NONCONS: shadow = nc_rpcf(&r, c); break;
          case ORC_TECH_RPCF_CONS: shadow = cs_rpcf(&r, c); break;
          case ORC_TECH_RSMSS: shadow = nc_rpcf(&r, c); break;   /* FilteredRBSM.frag:383 */
          default: break;
        }
      }
      vis[o] = shadow;
    }
  }
}

/* Measurement helper for bench.py (not a shader pass): how many shadow-map taps the PCSS program executes on this frame.
 * out[0] = pixels that reach the blocker search (each takes every tap of the blocker_search_size^2 grid), out[1] = pixels that
 * go on to the kernel_size^2 filter loop (PlausibleSoftShadow.frag:376-398: 0 < stepSize < 1), out[2] = foreground pixels. */
void orc_pcss_tap_count(const orc_params* p, const orc_camera* cam, const float light_mvp_b[16], const float* pos4,
                        const float* nrm4, int W, int H, const float* shadow_map, int64_t out[3]) {
  Smap s = {shadow_map, p->shadow_map_width, p->shadow_map_height, (float)p->shadow_map_width, (float)p->shadow_map_height};
  int64_t n_search = 0, n_filter = 0, n_fg = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : n_search, n_filter, n_fg)
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      size_t o = (size_t)j * W + i;
      V4 vertex = {pos4[4 * o], pos4[4 * o + 1], pos4[4 * o + 2], pos4[4 * o + 3]};
      if (vertex.x == 0.0f) continue;
      n_fg++;
      V4 normal = {nrm4[4 * o], nrm4[4 * o + 1], nrm4[4 * o + 2], nrm4[4 * o + 3]};
      V4 sc = mat4_mul_v4(light_mvp_b, vertex);
      V4 c = {sc.x / sc.w, sc.y / sc.w, sc.z / sc.w, sc.w / sc.w};
      if (!(sc.w > 0.0f && pre_evaluation(cam, p->shadow_intensity, vertex, normal) == 1.0f)) continue;
      n_search++;
      float stepSize = 2.0f * pcss_penumbra(p, &s, c) / (float)p->kernel_size;
      if (!(stepSize <= 0.0f || stepSize >= 1.0f)) n_filter++;
    }
  out[0] = n_search; out[1] = n_filter; out[2] = n_fg;
}

/* AccurateSoftShadow.frag:52-133 (monteCarlo branch: accFactor = 1, adaptiveSamplingLowerAccuracy = 0) */
void orc_visibility_multi(const orc_params* p, const float m[16], int N, const float* trans4, const float* pos4, int W,
                          int H, const float* shadow_maps, float* vis) {
  int x0 = p->rect_x0, y0 = p->rect_y0, x1 = p->rect_x1, y1 = p->rect_y1;
  if (x1 <= x0 || y1 <= y0) { x0 = 0; y0 = 0; x1 = W; y1 = H; }
  size_t layer = (size_t)p->shadow_map_width * p->shadow_map_height;
#pragma omp parallel for schedule(dynamic, 4)
  for (int j = y0; j < y1; j++)
    for (int i = x0; i < x1; i++) {
      size_t o = (size_t)j * W + i;
      float vx = pos4[4 * o], vy = pos4[4 * o + 1], vz = pos4[4 * o + 2];
      if (vx == 0.0f) continue;
      float accShadow = 0, count = 0, accFactor = 1.0f;
      float cx = m[0] * vx + m[4] * vy + m[8] * vz;
      float cy = m[1] * vx + m[5] * vy + m[9] * vz;
      float cz = m[2] * vx + m[6] * vy + m[10] * vz;
      float cw = m[3] * vx + m[7] * vy + m[11] * vz;
      for (int l = 0; l < N; l++) {
        float sx = cx + trans4[4 * l], sy = cy + trans4[4 * l + 1], sz = cz + trans4[4 * l + 2], sw = cw + trans4[4 * l + 3];
        sx = sx / sw; sy = sy / sw; sz = sz / sw;
        Smap s = {shadow_maps + layer * l, p->shadow_map_width, p->shadow_map_height, (float)p->shadow_map_width,
                  (float)p->shadow_map_height};
        float dfl = sm_fetch(&s, sx, sy);
        accShadow += ((sz <= dfl) ? 1.0f : p->shadow_intensity) * accFactor;
        count += accFactor;
      }
      vis[o] = p->multi_partial ? accShadow : accShadow / count;
    }
}

/* ================================ deferred Phong shading ====================================== */
/* ShadowMapping/Shaders/GBuffer/PhongShading.frag:11-47, literally: note `vec3 E = normalize(-vertex)` normalises the
 * vec4 (w = -1 included) before truncating to vec3, and specShadow = shadow - shadowIntensity. */
void orc_shade_phong(const orc_camera* cam, float si, const float* pos4, const float* nrm4, const float* albedo4,
                     const float* vis, int W, int H, const float clear4[4], float* out4) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      size_t o = (size_t)j * W + i;
      float* out = out4 + 4 * o;
      V4 vertex = {pos4[4 * o], pos4[4 * o + 1], pos4[4 * o + 2], pos4[4 * o + 3]};
      if (vertex.x == 0.0f) { out[0] = clear4[0]; out[1] = clear4[1]; out[2] = clear4[2]; out[3] = clear4[3]; continue; }
      float shadow = vis[o];
      float specShadow = shadow - si;
      V4 ev = mat4_mul_v4(cam->mv, vertex);
      const float* nm = cam->normal_matrix;
      float nx = nrm4[4 * o], ny = nrm4[4 * o + 1], nz = nrm4[4 * o + 2];
      float n[3], L[3], E[3], R[3];
      for (int k = 0; k < 3; k++) n[k] = (nm[0 + k] * nx + nm[3 + k] * ny) + nm[6 + k] * nz;
      { float inv = 1.0f / sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]); n[0] *= inv; n[1] *= inv; n[2] *= inv; }
      L[0] = cam->light_pos[0] - ev.x; L[1] = cam->light_pos[1] - ev.y; L[2] = cam->light_pos[2] - ev.z;
      { float inv = 1.0f / sqrtf((L[0] * L[0] + L[1] * L[1]) + L[2] * L[2]); L[0] *= inv; L[1] *= inv; L[2] *= inv; }
      { /* normalize(-vertex) on the vec4 */
        float mx = -ev.x, my = -ev.y, mz = -ev.z, mw = -ev.w;
        float inv = 1.0f / sqrtf(((mx * mx + my * my) + mz * mz) + mw * mw);
        E[0] = mx * inv; E[1] = my * inv; E[2] = mz * inv;
      }
      { /* R = normalize(-reflect(L, n)), reflect(I,N) = I - 2*dot(N,I)*N */
        float d = (n[0] * L[0] + n[1] * L[1]) + n[2] * L[2];
        float r0 = -(L[0] - 2.0f * d * n[0]), r1 = -(L[1] - 2.0f * d * n[1]), r2 = -(L[2] - 2.0f * d * n[2]);
        float inv = 1.0f / sqrtf((r0 * r0 + r1 * r1) + r2 * r2);
        R[0] = r0 * inv; R[1] = r1 * inv; R[2] = r2 * inv;
      }
      float ndl = glsl_max((n[0] * L[0] + n[1] * L[1]) + n[2] * L[2], 0.0f);
      float rde = glsl_max((R[0] * E[0] + R[1] * E[1]) + R[2] * E[2], 0.0f);
      float pw = powf(rde, 0.3f * 10.0f);
      const float amb[4] = {0.4f, 0.4f, 0.4f, 1.0f}, spec[4] = {0.25f, 0.25f, 0.25f, 1.0f}, diff[4] = {0.5f, 0.5f, 0.5f, 1.0f};
      for (int c = 0; c < 4; c++) {
        float col = albedo4 ? albedo4[4 * o + c] : 1.0f;
        float Idiff = diff[c] * ndl;
        float Ispec = (specShadow * spec[c]) * pw;
        out[c] = (shadow * col) * ((Idiff + Ispec) + amb[c]);
      }
    }
}

#include "oracle_edt_impl.h"    /* EDT shadow mapping: orc_edt_*, orc_mean_filter, orc_edtsm */
#include "oracle_moments_impl.h" /* VSM / ESM / EVSM / MSM: orc_moment_texel, orc_filter_moments, orc_visibility_moments */
