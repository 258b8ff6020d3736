/* oracle_edt_impl.h — TEST INFRASTRUCTURE (textually included at the end of oracle_shadow.c: uses its V4, mat4_mul_v4,
 * pre_evaluation, orc_visibility).
 *
 * CPU restatement of Euclidean-distance-transform shadow mapping (EDTSM), the ShadowMapping program's
 *   filterHardShadowsUsingEDT            ShadowMapping/src/main.cpp:416-447
 *   hard-shadow target with EDTSM == 1    Shaders/RBSM/NonConservativeSMSR.frag:362-395 (and ConservativeSMSR.frag main)
 *   initializeInput                       include/EDT/pba2DKernel.h:513-535   (site = boundary pixel of the hard shadow)
 *   pba2DVoronoiDiagram(16,16,16)         src/EDT/pba2DHost.cu:200-229        (exact nearest site per pixel)
 *   pbaNormalizeDistanceTransform         include/EDT/pba2DKernel.h:551-569   (world-space distance -> penumbra ramp)
 *   Shaders/Filter/MeanFilter.frag:35-77  two separable passes (:425-445)
 * What is pinned and how (tests/test_oracle_golden.py):
 *   - the hard-shadow target and MeanFilter.frag: bit-exact against the unmodified shaders compiled on the CPU;
 *   - initializeInput / pbaNormalizeDistanceTransform: restated from the CUDA source, double-precision sub-expressions
 *     kept where the C literals make them double (they cannot be compiled: legacy texture references, CUDA >= 12);
 *   - the Voronoi diagram: the exact Euclidean nearest site (what the Parallel Banding Algorithm computes); among
 *     equidistant sites PBA's pick depends on its band schedule, here the smallest (y, x) wins — checked against a
 *     brute-force search.
 * Where GL leaves the result open the oracle fixes it: the second filter pass samples a GL_LINEAR_MIPMAP_LINEAR texture
 * (main.cpp:888, MyGLTextureViewer.h:18) with derivatives taken inside a data-dependent loop, i.e. an undefined level of
 * detail; it is evaluated here as bilinear filtering of level 0 in fp32 (weights from fract(u - 0.5)). */

#define ORC_EDT_MARKER (-32768)                               /* pba2D.h:61 */
#define ORC_FOV 45.0f                                         /* MyGLGeometryViewer.cpp:6: passed to tan() as is */

/* pba2DKernel.h:503-510: `2.0 * n` is a double product, the quotient is double, narrowed on return */
static inline float edt_linearize_cuda(float depth) {
  const float n = 1.0f, f = 1000.0f;
  const float den = f + n - depth * (f - n);
  return (float)((2.0 * (double)n) / (double)den);
}
/* MeanFilter.frag:15-22: GLSL, all fp32 */
static inline float edt_linearize_glsl(float depth, int z_near, int z_far) {
  const float n = (float)z_near, f = (float)z_far;
  return (2.0f * n) / (f + n - depth * (f - n));
}

/* The RGBA32F hard-shadow target when EDTSM == 1: (shadow, camera window depth, pre-evaluated shadow, 1); cleared
 * (0,0,0,1) where the fragment is discarded.  NonConservativeSMSR.frag:362-395 */
void orc_edt_hard_image(const orc_params* p, const orc_camera* cam, const float cam_mvp[16], const float light_mvp_b[16],
                        const float* pos4, const float* nrm4, int W, int H, const float* shadow_map, float* img4) {
  orc_params q = *p;
  q.technique = (p->technique == ORC_TECH_EDTSM_CONS) ? ORC_TECH_RBSM_CONS : ORC_TECH_RBSM_NONCONS;
  q.rect_x0 = q.rect_y0 = q.rect_x1 = q.rect_y1 = 0;
  float* vis = (float*)calloc((size_t)W * H, sizeof(float));
  orc_visibility(&q, cam, light_mvp_b, pos4, nrm4, W, H, shadow_map, vis);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      const size_t o = (size_t)j * W + i;
      V4 vertex = {pos4[4 * o], pos4[4 * o + 1], pos4[4 * o + 2], pos4[4 * o + 3]};
      float* out = img4 + 4 * o;
      if (vertex.x == 0.0f) { out[0] = 0.0f; out[1] = 0.0f; out[2] = 0.0f; out[3] = 1.0f; continue; }
      V4 normal = {nrm4[4 * o], nrm4[4 * o + 1], nrm4[4 * o + 2], nrm4[4 * o + 3]};
      const V4 position = mat4_mul_v4(cam_mvp, vertex);
      float depth = position.z / position.w;
      depth = depth * 0.5f + 0.5f;
      out[0] = vis[o]; out[1] = depth; out[2] = pre_evaluation(cam, p->shadow_intensity, vertex, normal); out[3] = 1.0f;
    }
  free(vis);
}

/* initializeInput, pba2DKernel.h:513-535: site2[pixel] = (px, py) or (MARKER, MARKER) */
void orc_edt_sites(const float* img4, int W, int H, int16_t* site2) {
#pragma omp parallel for schedule(static)
  for (int py = 0; py < H; py++)
    for (int px = 0; px < W; px++) {
      const size_t o = (size_t)py * W + px;
      const float* c = img4 + 4 * o;
      int is_site = 0;
      for (int x = -1; x <= 1 && !is_site; x++)
        for (int y = -1; y <= 1 && !is_site; y++)
          if (px + x >= 0 && px + x < W && py + y >= 0 && py + y < H) {
            const float* q = img4 + 4 * ((size_t)(py + y) * W + (px + x));
            if (q[0] != c[0] && (double)fabsf(edt_linearize_cuda(c[1]) - edt_linearize_cuda(q[1])) <= 0.0025 && q[2] == 1.0f)
              is_site = 1;
          }
      site2[2 * o] = is_site ? (int16_t)px : (int16_t)ORC_EDT_MARKER;
      site2[2 * o + 1] = is_site ? (int16_t)py : (int16_t)ORC_EDT_MARKER;
    }
}

/* Exact nearest site per pixel (squared Euclidean distance in pixels; ties: smallest site y, then smallest site x).
 * near2 = (MARKER, MARKER) everywhere if the image has no site.  Two separable phases:
 *   1. per column, the nearest site row of that column for every row (tie: the smaller row);
 *   2. per pixel, the best of the columns' candidates, scanned outwards from the pixel's own column until the
 *      horizontal distance alone exceeds the best distance found. */
void orc_edt_nearest(const int16_t* site2, int W, int H, int16_t* near2) {
  int16_t* col = (int16_t*)malloc((size_t)W * H * sizeof(int16_t));     /* nearest site row in the same column */
#pragma omp parallel for schedule(static)
  for (int x = 0; x < W; x++) {
    int last = ORC_EDT_MARKER;
    for (int y = 0; y < H; y++) {                                       /* nearest at or below */
      if (site2[2 * ((size_t)y * W + x)] != ORC_EDT_MARKER) last = y;
      col[(size_t)y * W + x] = (int16_t)last;
    }
    last = ORC_EDT_MARKER;
    for (int y = H - 1; y >= 0; y--) {                                  /* nearest above: strictly closer wins */
      if (site2[2 * ((size_t)y * W + x)] != ORC_EDT_MARKER) last = y;
      const int below = col[(size_t)y * W + x];
      if (last != ORC_EDT_MARKER && (below == ORC_EDT_MARKER || last - y < y - below)) col[(size_t)y * W + x] = (int16_t)last;
    }
  }
#pragma omp parallel for schedule(dynamic, 4)
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      long long best = -1; int bx = ORC_EDT_MARKER, by = ORC_EDT_MARKER;
      for (int d = 0; d < W; d++) {
        if (best >= 0 && (long long)d * d > best) break;
        for (int s = 0; s < (d ? 2 : 1); s++) {
          const int c = s ? x + d : x - d;
          if (c < 0 || c >= W) continue;
          const int sy = col[(size_t)y * W + c];
          if (sy == ORC_EDT_MARKER) continue;
          const long long dd = (long long)d * d + (long long)(y - sy) * (y - sy);
          if (best < 0 || dd < best || (dd == best && (sy < by || (sy == by && c < bx)))) { best = dd; bx = c; by = sy; }
        }
      }
      near2[2 * ((size_t)y * W + x)] = (int16_t)bx; near2[2 * ((size_t)y * W + x) + 1] = (int16_t)by;
    }
  free(col);
}

/* pbaNormalizeDistanceTransform, pba2DKernel.h:551-569.  out2 = (normalised visibility, depth).  tex2D with the site
 * coordinates clamps (cudaAddressModeClamp): a MARKER site reads texel (0,0). */
void orc_edt_normalize(const float* img4, const float* pos4, const int16_t* near2, int W, int H, float penumbraSize,
                       float shadowIntensity, float* out2) {
#pragma omp parallel for schedule(static)
  for (int py = 0; py < H; py++)
    for (int px = 0; px < W; px++) {
      const size_t o = (size_t)py * W + px;
      int sx = near2[2 * o], sy = near2[2 * o + 1];
      sx = sx < 0 ? 0 : (sx > W - 1 ? W - 1 : sx); sy = sy < 0 ? 0 : (sy > H - 1 ? H - 1 : sy);
      const size_t so = (size_t)sy * W + sx;
      const float* ip = img4 + 4 * o; const float* sp = img4 + 4 * so;
      const float* p1 = pos4 + 4 * o; const float* p2 = pos4 + 4 * so;
      const float dx = p1[0] - p2[0], dy = p1[1] - p2[1], dz = p1[2] - p2[2];
      const float distance = sqrtf((dx * dx + dy * dy) + dz * dz);
      float r;
      if (sp[2] != 1.0f || (double)fabsf(edt_linearize_cuda(sp[1]) - edt_linearize_cuda(ip[1])) > 0.0005 ||
          distance > penumbraSize / 2) r = ip[0];
      else {
        const float q = distance / penumbraSize;
        const float v0 = (ip[0] == shadowIntensity) ? (float)(0.5 - (double)q) : (float)(0.5 + (double)q);
        r = (1 - shadowIntensity) * v0 + shadowIntensity * 1.0f;       /* plerp<float>(v0, 1.0, shadowIntensity) */
      }
      out2[2 * o] = r; out2[2 * o + 1] = ip[1];
    }
}

/* one texel of an (r,g) image with CLAMP_TO_BORDER (0): NEAREST, or bilinear of level 0 */
static inline void edt_fetch2(const float* img2, int W, int H, float u, float v, int linear, float* r, float* g) {
  if (!linear) {
    const float fi = floorf(u * (float)W), fj = floorf(v * (float)H);
    if (!(fi >= 0.0f && fi < (float)W && fj >= 0.0f && fj < (float)H)) { *r = 0.0f; *g = 0.0f; return; }
    const size_t o = (size_t)(int)fj * W + (int)fi;
    *r = img2[2 * o]; *g = img2[2 * o + 1];
    return;
  }
  const float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
  const float x0 = floorf(x), y0 = floorf(y);
  const float ax = x - x0, ay = y - y0;
  float acc[2] = {0.0f, 0.0f};
  for (int k = 0; k < 4; k++) {
    const float fx = x0 + (float)(k & 1), fy = y0 + (float)(k >> 1);
    const float wgt = ((k & 1) ? ax : 1.0f - ax) * ((k >> 1) ? ay : 1.0f - ay);
    float tr = 0.0f, tg = 0.0f;
    if (fx >= 0.0f && fx < (float)W && fy >= 0.0f && fy < (float)H) {
      const size_t o = (size_t)(int)fy * W + (int)fx;
      tr = img2[2 * o]; tg = img2[2 * o + 1];
    }
    acc[0] += wgt * tr; acc[1] += wgt * tg;
  }
  *r = acc[0]; *g = acc[1];
}

/* MeanFilter.frag:35-77, one pass.  in2/out2 = (r,g) per pixel; discarded pixels keep the cleared (0,0). */
void orc_mean_filter(const float* in2, const float* pos4, const float cam_mv[16], int W, int H, int order, int horizontal,
                     int z_near, int z_far, int linear, float* out2) {
  const float dscreen = 1.0f / (2.0f * tanf(ORC_FOV / 2.0f));
  const float scaleFactor = 50.0f;
  const float steps = 1.0f / (float)W, stept = 1.0f / (float)H;
  const float dirs = horizontal ? 1.0f : 0.0f, dirt = horizontal ? 0.0f : 1.0f;
#pragma omp parallel for schedule(dynamic, 4)
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      const size_t o = (size_t)j * W + i;
      out2[2 * o] = 0.0f; out2[2 * o + 1] = 0.0f;
      V4 vertex = {pos4[4 * o], pos4[4 * o + 1], pos4[4 * o + 2], pos4[4 * o + 3]};
      if (vertex.x == 0.0f) continue;
      /* f_texcoord of the pixel centre (MeanFilter.vert: texcoord*0.5+0.5 of the NDC position) */
      const float cs = (((float)i + 0.5f) / (float)W * 2.0f - 1.0f) * 0.5f + 0.5f;
      const float ct = (((float)j + 0.5f) / (float)H * 2.0f - 1.0f) * 0.5f + 0.5f;
      float cr, cg;
      edt_fetch2(in2, W, H, cs, ct, linear, &cr, &cg);
      int count = 0;
      float sum = 0.0f;
      const float deye = -(mat4_mul_v4(cam_mv, vertex)).z;
      float kernelCenter = (dscreen * (float)order * scaleFactor) / (deye * 2.0f);
      if (kernelCenter > 4096.0f) kernelCenter = 4096.0f;   /* guard (eye distance ~ 0): the shader would loop without end */
      for (float sample = -kernelCenter; sample <= kernelCenter; sample++) {
        float r, g;
        edt_fetch2(in2, W, H, cs + dirs * sample * steps, ct + dirt * sample * stept, linear, &r, &g);
        /* adjustColor :24-33 */
        if (r == 0.0f) r = cr;
        else if (fabsf(edt_linearize_glsl(g, z_near, z_far) - edt_linearize_glsl(cg, z_near, z_far)) >= 0.0005f) r = cr;
        sum += r;
        count++;
      }
      sum /= (float)count;
      out2[2 * o] = sum; out2[2 * o + 1] = cg;
    }
}

/* the whole pass sequence of display() with shadowParams.EDTSM: hard shadows, EDT filter, two mean-filter passes */
void orc_edtsm(const orc_params* p, const orc_camera* cam, const float cam_mvp[16], const float light_mvp_b[16],
               const float* pos4, const float* nrm4, int W, int H, const float* shadow_map, float* vis, int16_t* near2_out) {
  const size_t px = (size_t)W * H;
  float* img4 = (float*)malloc(px * 16);
  int16_t* site2 = (int16_t*)malloc(px * 4);
  int16_t* near2 = (int16_t*)malloc(px * 4);
  float* a2 = (float*)malloc(px * 8); float* b2 = (float*)malloc(px * 8);
  orc_edt_hard_image(p, cam, cam_mvp, light_mvp_b, pos4, nrm4, W, H, shadow_map, img4);
  orc_edt_sites(img4, W, H, site2);
  orc_edt_nearest(site2, W, H, near2);
  orc_edt_normalize(img4, pos4, near2, W, H, (float)((double)p->penumbra_size / 5.0), p->shadow_intensity, a2);   /* main.cpp:421: int / 5.0, narrowed */
  orc_mean_filter(a2, pos4, cam->mv, W, H, p->kernel_order, 1, p->z_near, p->z_far, 0, b2);                /* :425-433 */
  orc_mean_filter(b2, pos4, cam->mv, W, H, p->kernel_order, 0, p->z_near, p->z_far, 1, a2);                /* :435-443 */
  for (size_t o = 0; o < px; o++) vis[o] = a2[2 * o];
  if (near2_out) memcpy(near2_out, near2, px * 4);
  free(img4); free(site2); free(near2); free(a2); free(b2);
}
