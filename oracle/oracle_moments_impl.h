/*
 * oracle_moments_impl.h — pre-filtered ("moment") shadow maps of the ShadowMapping program: VSM, ESM, EVSM, MSM
 * (SURVEY.md §8(f) row 4).  TEST INFRASTRUCTURE ONLY (see oracle.h); included at the end of oracle_shadow.c.
 *
 * Reference (paths relative to /root/reference/ShadowMapping):
 *   moment render     Shaders/ShadowMap/Moments.frag:19-49 (VSM, MSM), Exponential.frag:15-22 (ESM),
 *                     ExponentialMoments.frag:15-30 (EVSM); bound by displaySceneFromLightPOV, src/main.cpp:227-243
 *   quantisation      MyGLGeometryViewer::configureMoments, src/Viewers/MyGLGeometryViewer.cpp:188-213
 *                     (glm::transpose / glm::inverse, include/glm/core/func_matrix.inl:523-588)
 *   separable blur    filterShadowMap, src/main.cpp:374-398; Shaders/Filter/GaussianFilter.frag:10-34,
 *                     LogGaussianFilter.frag:10-49; weights Filter::buildGaussianKernel, src/Filter.cpp:17-46
 *   reconstruction    Shaders/Shadow.frag:118-220 (chebyshevUpperBound, varianceShadowMapping, exponentialShadowMapping,
 *                     exponentialVarianceShadowMapping, hamburger4MSM) and main :240-273
 *
 * Pinned bit-exactly against the unmodified shaders above compiled through oracle/ref_build (tests/test_oracle_golden.py)
 * and against the reference's GLM for the quantisation matrices.
 *
 * What GL leaves to the implementation and is DEFINED here (stated again in DESIGN.md §2):
 *   - every texture of the chain is GL_LINEAR_MIPMAP_LINEAR (MyGLTextureViewer.h:18) sampled with implicit derivatives; the
 *     level of detail is the driver's choice.  As for the EDT chain, lookups are bilinear filtering of level 0 in fp32
 *     (weights fract(u*size - 0.5), texels summed in the order 00, 10, 01, 11, border colour 0);
 *   - dFdx / dFdy (Moments.frag:34-35) are the fine 2x2-quad differences of the value on the fragment's own triangle:
 *     dFdx = v(x|1, y) - v(x&~1, y), dFdy = v(x, y|1) - v(x, y&~1), the partner pixel being evaluated on the same plane
 *     whether or not the triangle covers it (a helper invocation);
 *   - `position.z / position.w * 0.5 + 0.5` of the perspective-correct varying equals gl_FragCoord.z before polygon
 *     offset; it is taken from the rasteriser's depth plane without the offset and without the [0,1] clamp.
 */

/* ---- MyGLGeometryViewer.cpp:193-199: the optimised moment quantisation and its inverse -------------------------- */
void orc_msm_quantization(float m[16], float minv[16], float t[4]) {
  /* rows as typed at :193-196; glm::transpose(:198) makes row r, column c of this table element [c][r] of the uniform,
     i.e. (mQuantization * v)[r] = sum_c table[r][c] * v[c] */
  static const float table[4][4] = {
      {-2.07224649f, 32.2370378f, -68.5710746f, 39.3703274f},
      {13.7948857f, -59.4683976f, 82.035975f, -35.3649032f},
      {0.105877704f, -1.90774663f, 9.34965551f, -6.65434907f},
      {9.79240621f, -33.76521106f, 47.9456097f, -23.9728048f}};
  /* typed mQuantization[i][j] = table[i][j] (column i, row j), then transposed: column c, row r = table[r][c] */
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) m[c * 4 + r] = table[r][c];
  t[0] = 0.0359558848f; t[1] = 0.0f; t[2] = 0.0f; t[3] = 0.0f;       /* :208 */
#define M(c, r) m[(c) * 4 + (r)]
  /* glm::inverse(mat4), func_matrix.inl:530-587: 2x2 sub-determinants, cofactor columns, division by the determinant */
  float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3), c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
  float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3), c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
  float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2), c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
  float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3), c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
  float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2), c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
  float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1), c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
  const float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  const float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
  const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
  const float sa[4] = {1.0f, -1.0f, 1.0f, -1.0f}, sb[4] = {-1.0f, 1.0f, -1.0f, 1.0f};
  float inv[16];
  for (int k = 0; k < 4; k++) {
    inv[0 * 4 + k] = sa[k] * ((v1[k] * f0[k] - v2[k] * f1[k]) + v3[k] * f2[k]);
    inv[1 * 4 + k] = sb[k] * ((v0[k] * f0[k] - v2[k] * f3[k]) + v3[k] * f4[k]);
    inv[2 * 4 + k] = sa[k] * ((v0[k] * f1[k] - v1[k] * f3[k]) + v3[k] * f5[k]);
    inv[3 * 4 + k] = sb[k] * ((v0[k] * f2[k] - v1[k] * f4[k]) + v2[k] * f5[k]);
  }
  float det = ((M(0, 0) * inv[0] + M(0, 1) * inv[4]) + M(0, 2) * inv[8]) + M(0, 3) * inv[12];
#undef M
  for (int k = 0; k < 16; k++) minv[k] = inv[k] / det;
}

/* linearize(): Moments.frag:9-16 == Shadow.frag:32-39 */
static inline float mom_linearize(float depth, int z_near, int z_far) {
  float n = (float)z_near, f = (float)z_far;
  return (2.0f * n) / (f + n - depth * (f - n));
}

/* One texel of the moment target.  zwin = window depth of the fragment (position.z / position.w * 0.5 + 0.5),
 * zwin_px / zwin_py = the same plane at the quad partners (x^1, y) and (x, y^1); x_odd / y_odd say on which side they lie. */
void orc_moment_texel(int technique, float zwin, float zwin_px, float zwin_py, int x_odd, int y_odd, int z_near, int z_far,
                      float out4[4]) {
  float depth = mom_linearize(zwin, z_near, z_far);
  if (technique == ORC_TECH_ESM) { out4[0] = depth; out4[1] = 0.0f; out4[2] = 0.0f; out4[3] = 1.0f; return; }   /* Exponential.frag:21 */
  float m0 = depth, m1 = depth * depth;
  if (technique == ORC_TECH_VSM || technique == ORC_TECH_EVSM) {
    float dpx = mom_linearize(zwin_px, z_near, z_far), dpy = mom_linearize(zwin_py, z_near, z_far);
    float dx = x_odd ? depth - dpx : dpx - depth, dy = y_odd ? depth - dpy : dpy - depth;
    m1 = m1 + 0.25f * (dx * dx + dy * dy);                                                /* Moments.frag:36 */
    if (technique == ORC_TECH_VSM) { out4[0] = m0; out4[1] = m1; out4[2] = 0.0f; out4[3] = 0.0f; }
    else { out4[0] = m0; out4[1] = m1; out4[2] = depth; out4[3] = 1.0f; }                  /* ExponentialMoments.frag:28 */
    return;
  }
  /* MSM, Moments.frag:40-46 */
  float q[16], qi[16], t[4];
  orc_msm_quantization(q, qi, t);
  float m2 = depth * depth * depth, m3 = depth * depth * depth * depth;
  for (int r = 0; r < 4; r++) out4[r] = (((q[0 + r] * m0 + q[4 + r] * m1) + q[8 + r] * m2) + q[12 + r] * m3) + t[r];
}

/* Filter::buildGaussianKernel, Filter.cpp:17-46: row `order - 1` of Pascal's triangle over 2^(order-1) */
void orc_gaussian_kernel(int order, float* kernel) {
  float norm = powf(2.0f, (float)(order - 1));
  int coef = 1;
  for (int j = 0; j < order; j++) {
    if (j > 0) coef = coef * (order - 1 - j + 1) / j;
    kernel[j] = (float)coef / norm;
  }
}

/* GL_LINEAR of level 0, CLAMP_TO_BORDER (0,0,0,0), RGBA32F (definition in the header of this file) */
static inline void mom_fetch4(const float* img4, int w, int h, float u, float v, float out[4]) {
  float fw = (float)w, fh = (float)h;
  float x = u * fw - 0.5f, y = v * fh - 0.5f;
  float x0 = floorf(x), y0 = floorf(y), ax = x - x0, ay = y - y0;
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  for (int k = 0; k < 4; k++) {
    float wgt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
    float fi = x0 + (float)(k & 1), fj = y0 + (float)(k >> 1);
    float t[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (fi >= 0.0f && fi < fw && fj >= 0.0f && fj < fh) memcpy(t, img4 + 4 * ((size_t)(int)fj * w + (size_t)(int)fi), 16);
    for (int c = 0; c < 4; c++) out[c] = out[c] + t[c] * wgt;
  }
}

/* One pass of filterShadowMap (main.cpp:380-392): a W x H target, the full-screen quad of GaussianFilter.vert:4-9, the
 * source (sw x sh) read with `step = 1/width, 1/height` of the TARGET (drawTextureOnShader's imageWidth/imageHeight).
 * log_space = 0: GaussianFilter.frag:10-34 on all four channels; 1: LogGaussianFilter.frag:10-49 on .x, replicated. */
void orc_filter_moments(const float* src4, int sw, int sh, int W, int H, int order, const float* kernel, int horizontal,
                        int log_space, float* dst4) {
  const float step_s = 1.0f / (float)W, step_t = 1.0f / (float)H;
  const float dir_s = horizontal ? 1.0f : 0.0f, dir_t = horizontal ? 0.0f : 1.0f;
  const int kc = order / 2;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      float nx = ((float)i + 0.5f) / (float)W * 2.0f - 1.0f, ny = ((float)j + 0.5f) / (float)H * 2.0f - 1.0f;
      float cs = nx * 0.5f + 0.5f, ct = ny * 0.5f + 0.5f;
      float* o = dst4 + 4 * ((size_t)j * W + i);
      float t[4];
      if (!log_space) {
        float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int s = -kc; s <= kc; s++) {
          mom_fetch4(src4, sw, sh, cs + dir_s * (float)s * step_s, ct + dir_t * (float)s * step_t, t);
          for (int c = 0; c < 4; c++) sum[c] = sum[c] + t[c] * kernel[kc + s];
        }
        memcpy(o, sum, 16);
      } else {
        int ks = -kc;
        mom_fetch4(src4, sw, sh, cs + dir_s * (float)ks * step_s, ct + dir_t * (float)ks * step_t, t);
        float s0 = t[0];
        ks++;
        mom_fetch4(src4, sw, sh, cs + dir_s * (float)ks * step_s, ct + dir_t * (float)ks * step_t, t);
        float s1 = t[0];
        float sum = s0 + logf(kernel[0] + (kernel[1] * expf(s1 - s0)));                   /* log_conv, :10-13 */
        for (int k = 2; k < order; k++) {
          ks++;
          mom_fetch4(src4, sw, sh, cs + dir_s * (float)ks * step_s, ct + dir_t * (float)ks * step_t, t);
          sum = sum + logf(1.0f + (kernel[k] * expf(t[0] - sum)));
        }
        o[0] = o[1] = o[2] = o[3] = sum;
      }
    }
}

/* Shadow.frag:118-132 */
static float mom_chebyshev(float m0, float m1, float z, float si) {
  if (z <= m0) return 1.0f;
  float variance = m1 - (m0 * m0);
  float d = z - m0;
  float p_max = variance / (variance + d * d);
  float p = (z <= m0) ? 1.0f : 0.0f;
  p_max = glsl_max(p, p_max);
  return glsl_mix(p_max, 1.0f, si);
}
static inline float mom_clamp(float x, float lo, float hi) { return glsl_min(glsl_max(x, lo), hi); }

/* Shadow.frag:168-220 */
static float mom_hamburger(const float bq[4], float zrecv, float si) {
  float q[16], qi[16], t[4], b[4], v[4];
  orc_msm_quantization(q, qi, t);
  for (int r = 0; r < 4; r++) v[r] = bq[r] - t[r];
  for (int r = 0; r < 4; r++) b[r] = ((qi[0 + r] * v[0] + qi[4 + r] * v[1]) + qi[8 + r] * v[2]) + qi[12 + r] * v[3];
  const float bias = 0.00003f;
  float zx = zrecv;
  for (int r = 0; r < 4; r++) b[r] = (1.0f - bias) * b[r] + bias * 0.5f;
  float d0 = 1.0f, d1 = zx, d2 = zx * zx;
  float L10 = b[0], L20 = b[1];
  float D11 = b[1] - L10 * L10;
  float L21 = (b[2] - L20 * L10) / D11;
  float D22 = b[3] - L20 * L20 - L21 * L21 * D11;
  float y0 = d0, y1 = d1 - L10 * y0, y2 = d2 - L20 * y0 - L21 * y1;
  y1 /= D11; y2 /= D22;
  float cz = y2, cy = y1 - L21 * cz, cx = y0 - L10 * cy - L20 * cz;
  float p = cy / cz, qq = cx / cz;
  float D = ((p * p) / 4.0f) - qq;
  float r = sqrtf(D);
  float zy = -(p / 2.0f) - r, zz = -(p / 2.0f) + r;
  if (zx <= zy) return 1.0f;
  else if (zx <= zz)
    return mom_clamp((1.0f - mom_clamp((zx * zz - b[0] * (zx + zz) + b[1]) / ((zz - zy) * (zx - zy)), 0.0f, 1.0f)), si, 1.0f);
  else
    return mom_clamp((1.0f - mom_clamp(1.0f - (zy * zz - b[0] * (zy + zz) + b[1]) / ((zx - zy) * (zx - zz)), 0.0f, 1.0f)), si, 1.0f);
}

/* Shadow.frag main (:240-273) with VSM / ESM / EVSM / MSM == 1; fmap4 = FILTER_Y_MAP_COLOR (mw x mh, main.cpp:316) */
void orc_visibility_moments(const orc_params* p, const orc_camera* cam, const float light_mvp_b[16], const float* pos4,
                            const float* nrm4, int W, int H, const float* fmap4, int mw, int mh, float* vis) {
  int x0 = p->rect_x0, y0 = p->rect_y0, x1 = p->rect_x1, y1 = p->rect_y1;
  if (x1 <= x0 || y1 <= y0) { x0 = 0; y0 = 0; x1 = W; y1 = H; }
  const int tech = p->technique;
  const float si = p->shadow_intensity;
#pragma omp parallel for schedule(dynamic, 4)
  for (int j = y0; j < y1; j++)
    for (int i = x0; i < x1; i++) {
      size_t o = (size_t)j * W + i;
      V4 vertex = {pos4[4 * o], pos4[4 * o + 1], pos4[4 * o + 2], pos4[4 * o + 3]};
      if (vertex.x == 0.0f) continue;
      V4 normal = {nrm4[4 * o], nrm4[4 * o + 1], nrm4[4 * o + 2], nrm4[4 * o + 3]};
      V4 sc = mat4_mul_v4(light_mvp_b, vertex);
      V4 c = {sc.x / sc.w, sc.y / sc.w, sc.z / sc.w, sc.w / sc.w};
      float shadow = pre_evaluation(cam, si, vertex, normal);
      if (sc.w > 0.0f && shadow == 1.0f) {
        float b[4];
        mom_fetch4(fmap4, mw, mh, c.x, c.y, b);
        float z = mom_linearize(c.z, p->z_near, p->z_far);
        if (tech == ORC_TECH_VSM) shadow = mom_chebyshev(b[0], b[1], z, si);                 /* :134-141 */
        else if (tech == ORC_TECH_ESM) {                                                      /* :144-157 */
          float e2 = expf(80.0f * b[0]);
          float e1 = expf(-80.0f * z);
          shadow = mom_clamp(e1 * e2, si, 1.0f);
        } else if (tech == ORC_TECH_EVSM) {                                                   /* :160-175 */
          float variance = mom_chebyshev(b[0], b[1], z, si);
          float e1 = expf(-60.0f * z);
          float e2 = expf(60.0f * b[2]);
          shadow = glsl_min(variance, mom_clamp(e1 * e2, si, 1.0f));
        } else shadow = mom_hamburger(b, z, si);
      }
      vis[o] = shadow;
    }
}
