/*
 * oracle.h — CPU restatement of the reference's shadow hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and only as
 * the checker or the reported CPU baseline.  The product (globalillumination_b200/) never links,
 * imports or calls this.
 *
 * What it restates (reference = MarcioCerqueira/GlobalIllumination, paths relative to /root/reference):
 *   matrices            ShadowMapping/include/glm/gtc/matrix_transform.inl:44-78,223-244,383-411,
 *                       ShadowMapping/src/Viewers/MyGLGeometryViewer.cpp:14-19,108-134,136-186
 *   light depth pass    ShadowMapping/Shaders/Scene.vert:11-20 + fixed-function GL raster
 *                       (ShadowMapping/src/main.cpp:246-247,350-361)
 *   G-buffer pass       ShadowMapping/Shaders/GBuffer/GBuffer.vert:12-23, GBuffer.frag:32-38
 *   hard / PCF          ShadowMapping/Shaders/Shadow.frag:86-116,222-273
 *   PCSS                SoftShadowMapping/Shaders/SoftShadow/PlausibleSoftShadow.frag:33-49,166-194,365-398,556-563,605-633
 *   RBSM                ShadowMapping/Shaders/RBSM/{NonConservativeSMSR,ConservativeSMSR,FilteredRBSM}.frag
 *   many-light          SoftShadowMapping/Shaders/SoftShadow/AccurateSoftShadow.frag:52-133,
 *                       SoftShadowMapping/src/Scene/LightSource/UniformSampledLightSource.cpp:27-38
 *   shadow volumes      ShadowVolumes/src/ShadowVolume.cpp:15-195, ShadowVolumes/src/main.cpp:120-206
 *
 * Pinning: the reference ships no tests or golden vectors.  The per-pixel passes are
 * pinned against the reference's own, unmodified GLSL sources compiled as C++ (oracle/ref_build ->
 * oracle/_ref/libref_shaders.so) and the host arithmetic against the reference's own C++ sources
 * (Mesh/SceneLoader/OBJLoader/ShadowVolume/UniformSampledLightSource + vendored GLM) compiled into
 * oracle/_ref/libref_host.so; goldens from both live in tests/golden/.  The fixed-function rasteriser
 * (triangle coverage, depth interpolation, polygon offset, clipping) has no source in the reference — it
 * is the OpenGL driver — so that part is DEFINED here (rules in DESIGN.md §3) and is "parity unpinned"
 * against a real GL implementation.
 *
 * All arithmetic is IEEE fp32 evaluated in source order, no FMA contraction (-ffp-contract=off).
 * Matrices are column-major float[16] (m[c*4+r]) as in GLM / glUniformMatrix4fv(GL_FALSE).
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  ORC_TECH_HARD = 0,          /* Shadow.frag naive==1                                  */
  ORC_TECH_PCF = 1,           /* Shadow.frag PCF (bilinearPCF==1)                      */
  ORC_TECH_PCSS = 2,          /* PlausibleSoftShadow.frag PCSS==1                      */
  ORC_TECH_RBSM_NONCONS = 3,  /* NonConservativeSMSR.frag SMSR==1                      */
  ORC_TECH_RBSM_CONS = 4,     /* ConservativeSMSR.frag SMSR==1                         */
  ORC_TECH_RPCF_NONCONS = 5,  /* NonConservativeSMSR.frag RPCFPlusSMSR==1              */
  ORC_TECH_RPCF_CONS = 6,     /* ConservativeSMSR.frag RPCFPlusSMSR==1                 */
  ORC_TECH_RSMSS = 7,         /* FilteredRBSM.frag (always the accurate-RPCF branch)   */
  ORC_TECH_MULTI_HARD = 8,    /* AccurateSoftShadow.frag monteCarlo (N lights)         */
  ORC_TECH_RBSSM = 9,         /* RBSSM.frag (revectorization-based soft shadows)       */
  ORC_TECH_EDTSM_NONCONS = 10,/* EDT shadow mapping over NonConservativeSMSR.frag      */
  ORC_TECH_EDTSM_CONS = 11,   /* EDT shadow mapping over ConservativeSMSR.frag         */
  ORC_TECH_VSM = 12,          /* Shadow.frag VSM==1  (variance shadow mapping)         */
  ORC_TECH_ESM = 13,          /* Shadow.frag ESM==1  (exponential)                     */
  ORC_TECH_EVSM = 14,         /* Shadow.frag EVSM==1 (exponential variance)            */
  ORC_TECH_MSM = 15,          /* Shadow.frag MSM==1  (Hamburger 4-moment)              */
  ORC_TECH_PCF_TRICUBIC = 16  /* Shadow.frag PCF with tricubicPCF==1 (textureBicubic)  */
};

enum { ORC_DEPTH_LESS = 0, ORC_DEPTH_LEQUAL = 1 };

/* Same field order and types as sgi_params in include/shadowgi.h (kept separate on purpose). */
typedef struct orc_params {
  int32_t technique;
  int32_t shadow_map_width, shadow_map_height;
  float   shadow_intensity;
  int32_t kernel_order;          /* Shadow.frag / RPCF float-loop PCF                  */
  int32_t penumbra_size;
  int32_t blocker_search_size;   /* PCSS                                               */
  int32_t kernel_size;
  int32_t light_source_radius;
  int32_t max_search;            /* RBSM                                               */
  float   depth_threshold;
  int32_t z_near, z_far;         /* `uniform int zNear/zFar`                           */
  float   polygon_offset_factor, polygon_offset_units;
  int32_t sv_depth_func;
  int32_t sv_infinity;
  int32_t rect_x0, rect_y0, rect_x1, rect_y1;   /* screen rectangle to evaluate        */
  int32_t multi_partial;         /* many-light: write the un-normalised sum over the given lights */
  int32_t multi_fused;           /* (implementation switch of the CUDA path; the oracle has one form) */
  int32_t sv_silhouette;         /* shadow volumes: drop the side quads of interior edges in cancelling pairs */
  int32_t sv_zfail;              /* shadow volumes: depth-fail counting over capped volumes, depth clamp */
} orc_params;

/* per-frame camera-side uniforms of the full-screen shadow passes */
typedef struct orc_camera {
  float mv[16];            /* camera view*model                                   */
  float normal_matrix[9];  /* frozen inverseTranspose(mat3(mv)), column-major     */
  float light_pos[3];      /* `lightPosition` uniform: light eye rotated 180° about Y */
} orc_camera;

/* ---- matrices (GLM 0.9.3.1, degrees API) ------------------------------------------------------ */
void orc_mat4_identity(float m[16]);
void orc_mat4_mul(const float a[16], const float b[16], float out[16]);     /* out = a*b            */
void orc_perspective(float fovy_deg, float aspect, float z_near, float z_far, float out[16]);
void orc_look_at(const float eye[3], const float at[3], const float up[3], float out[16]);
void orc_rotate(const float m[16], float angle_deg, const float axis[3], float out[16]);
void orc_translate(const float m[16], const float v[3], float out[16]);
void orc_bias_mul(const float light_mvp[16], float out[16]);                /* bias * lightMVP      */
void orc_normal_matrix(const float mv[16], float out9[9]);                  /* inverseTranspose(mat3) */
void orc_rotate_light_180(const float eye[3], float out[3]);               /* main.cpp:283         */
void orc_uniform_light_sample(const float p[3], int size, int n_lights, int index, float out[3]);

/* ---- float-loop PCF tap offsets (F3) ---------------------------------------------------------- */
int  orc_pcf_offsets(int kernel_order, int penumbra_size, int inclusive, float* out, int cap);

/* ---- rasteriser (rules defined in DESIGN.md §3) ----------------------------------------------- */
/* depth[H][W] fp32, row 0 = bottom, cleared to 1.0, keep-min, polygon offset (factor, units) */
int  orc_raster_depth(const float* xyz, int V, const int32_t* idx, int T, const float mvp[16],
                      int W, int H, float factor, float units, float* depth);
/* pos4/nrm4: float4[H][W]; background pos=(0,0,0,1), nrm=(0,0,0,1); depth as above w/o offset */
int  orc_raster_gbuffer(const float* xyz, const float* nrm, int V, const int32_t* idx, int T,
                        const float mvp[16], int W, int H, float* pos4, float* nrm4, float* depth);

/* same, plus the third MRT of GBuffer.frag:32-38 with useMeshColor==1: albedo4 = (interpolated vertex colour, 1),
 * background (0,0,0,1).  rgb/albedo4 may be NULL. */
int  orc_raster_gbuffer_ex(const float* xyz, const float* nrm, const float* rgb, int V, const int32_t* idx, int T,
                           const float mvp[16], int W, int H, float* pos4, float* nrm4, float* albedo4, float* depth);

/* scene textures (Mesh::loadTexture / loadRGBTexture, MyGLTextureViewer.cpp:45-56): RGB8, row 0 = t 0, GL_LINEAR, GL_REPEAT */
typedef struct orc_texture { const uint8_t* rgb; int32_t w, h; } orc_texture;
/* GBuffer.frag:11-30 computeFragmentColor with useTextureForColoring == 1 on one fragment (tex[3] = texture0..2) */
void orc_fragment_color(const float uvw[3], const float rgb[3], const orc_texture tex[3], float out[4]);
/* the G-buffer with the texture select on the (u, v, texture id) varying; uv / tex NULL = orc_raster_gbuffer_ex */
int  orc_raster_gbuffer_tex(const float* xyz, const float* nrm, const float* rgb, const float* uv, int V, const int32_t* idx, int T,
                            const float mvp[16], int W, int H, const orc_texture* tex, float* pos4, float* nrm4, float* albedo4,
                            float* depth);

/* ---- deferred shading: ShadowMapping/Shaders/GBuffer/PhongShading.frag:11-47 (shadeScene, main.cpp:449-457) ----
 * out4[H][W] float4; discarded (background) pixels get clear4 (0.63, 0.82, 0.96, 1). */
void orc_shade_phong(const orc_camera* cam, float shadow_intensity, const float* pos4, const float* nrm4,
                     const float* albedo4, const float* vis, int W, int H, const float clear4[4], float* out4);

/* ---- per-pixel shadow passes ------------------------------------------------------------------ */
/* light_mvp_b = bias*lightMVP.  vis[H][W], background pixels keep 0.                              */
void orc_visibility(const orc_params* p, const orc_camera* cam, const float light_mvp_b[16],
                    const float* pos4, const float* nrm4, int W, int H,
                    const float* shadow_map, float* vis);
/* bench.py measurement helper: out = {pixels running the blocker search, pixels running the filter loop, foreground pixels} */
void orc_pcss_tap_count(const orc_params* p, const orc_camera* cam, const float light_mvp_b[16], const float* pos4,
                        const float* nrm4, int W, int H, const float* shadow_map, int64_t out[3]);
/* many-light: maps[N][S][S]; common3x4 taken from light_mvp_b_common; trans[N][4]; weights NULL=1 */
void orc_visibility_multi(const orc_params* p, const float light_mvp_b_common[16], int N,
                          const float* trans4, const float* pos4, int W, int H,
                          const float* shadow_maps, float* vis);

/* ---- EDT shadow mapping (oracle_edt_impl.h) --------------------------------------------------- */
void orc_edt_hard_image(const orc_params* p, const orc_camera* cam, const float cam_mvp[16], const float light_mvp_b[16],
                        const float* pos4, const float* nrm4, int W, int H, const float* shadow_map, float* img4);
void orc_edt_sites(const float* img4, int W, int H, int16_t* site2);
void orc_edt_nearest(const int16_t* site2, int W, int H, int16_t* near2);
void orc_edt_normalize(const float* img4, const float* pos4, const int16_t* near2, int W, int H, float penumbraSize,
                       float shadowIntensity, float* out2);
void orc_mean_filter(const float* in2, const float* pos4, const float cam_mv[16], int W, int H, int order, int horizontal,
                     int z_near, int z_far, int linear, float* out2);
void orc_edtsm(const orc_params* p, const orc_camera* cam, const float cam_mvp[16], const float light_mvp_b[16],
               const float* pos4, const float* nrm4, int W, int H, const float* shadow_map, float* vis, int16_t* near2_out);

/* ---- moment shadow maps: VSM / ESM / EVSM / MSM (oracle_moments_impl.h, oracle_raster.c) ------- */
void orc_msm_quantization(float m[16], float minv[16], float t[4]);
void orc_moment_texel(int technique, float zwin, float zwin_px, float zwin_py, int x_odd, int y_odd, int z_near, int z_far,
                      float out4[4]);
/* mom4: float4[H][W] moment target of the light-view pass, cleared to (0,0,0,1) */
int  orc_raster_moments(const float* xyz, int V, const int32_t* idx, int T, const float mvp[16], int W, int H, float factor,
                        float units, int technique, int z_near, int z_far, float* mom4);
void orc_gaussian_kernel(int order, float* kernel);
/* one separable pass of filterShadowMap into a W x H target; src4 is sw x sh */
void orc_filter_moments(const float* src4, int sw, int sh, int W, int H, int order, const float* kernel, int horizontal,
                        int log_space, float* dst4);
/* fmap4: the twice-filtered map (mw x mh = the window size in the reference) */
void orc_visibility_moments(const orc_params* p, const orc_camera* cam, const float light_mvp_b[16], const float* pos4,
                            const float* nrm4, int W, int H, const float* fmap4, int mw, int mh, float* vis);

/* ---- shadow volumes --------------------------------------------------------------------------- */
/* prism_xyz: 6T vertices*3, prism_idx: 6T triangles*3 (ShadowVolume::build/update)                */
void orc_sv_build_prisms(const float* xyz, const float* nrm, int V, const int32_t* idx, int T,
                         const float light[3], int infinity, float* prism_xyz, int32_t* prism_idx);
/* z-pass signed count per pixel of prism fragments passing the depth test against scene_depth     */
int  orc_sv_count(const float* prism_xyz, int PV, const int32_t* prism_idx, int PT, const float mvp[16],
                  int W, int H, const float* scene_depth, int depth_func,
                  int32_t* count, uint8_t* stencil);

/* silhouette form: keep[3T] = which side quads survive the pairwise cancellation of interior edges */
void orc_sv_silhouette_keep(const float* nrm, const int32_t* idx, int T, const float light[3], uint8_t* keep);
/* 6 (caps == 0) or 8 (sides, near cap, far cap) triangles per source triangle; dropped quads are degenerate (0,0,0) */
void orc_sv_build_volumes(const float* xyz, const float* nrm, int V, const int32_t* idx, int T, const float light[3], int infinity,
                          const uint8_t* keep, int caps, float* prism_xyz, int32_t* vol_idx);
int  orc_sv_count_ex(const float* prism_xyz, int PV, const int32_t* prism_idx, int PT, const float mvp[16], int W, int H,
                     const float* scene_depth, int depth_func, int zfail, int per, int32_t* count, uint8_t* stencil);

int  orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
