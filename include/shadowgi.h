/*
 * shadowgi.h — C ABI of the B200-native shadow hot path (libshadowgi.so).
 *
 * Drop-in boundary: the reference (MarcioCerqueira/GlobalIllumination) has no plugin API; the seam is
 * the set of free functions its display() calls once per frame plus the ShadowParams block they share.
 * Each entry point below names the reference interface it replaces (paths relative to the reference
 * root).  The only C ABI precedent in the reference is the EDT module
 * (ShadowMapping/include/EDT/pba2D.h:51-56: init / compute / deinit on caller-owned device memory);
 * this keeps that shape with an opaque per-GPU context and error codes instead of global state.
 *
 * Conventions
 *   - every function returns SGI_OK (0) or a negative sgi_status; it never exits or throws;
 *     sgi_last_error(ctx) gives the message of the last failure on that context;
 *   - host input pointers are borrowed for the duration of the call and copied to the device;
 *   - outputs live in device memory owned by the context (HBM-resident between passes);
 *     sgi_read copies one to a host buffer, sgi_device_ptr lends the device pointer;
 *   - one CUDA stream per context; all passes are asynchronous on it, sgi_read / sgi_synchronize block;
 *   - matrices are column-major float[16] exactly as the reference hands them to
 *     glUniformMatrix4fv(..., GL_FALSE, &m[0][0]); images are row-major, row 0 = bottom (GL window
 *     coordinates), so buffer texel (i, j) is what the reference's full-screen passes read at pixel (i, j);
 *   - one context per GPU / process rank; a context is not thread-safe.
 *
 * No OpenGL, no torch types, no C++ in this header.
 */
#ifndef SHADOWGI_H
#define SHADOWGI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgi_ctx sgi_ctx;

typedef enum sgi_status {
  SGI_OK = 0,
  SGI_ERR_INVALID = -1,     /* bad argument / pass called before its inputs were set            */
  SGI_ERR_CUDA = -2,        /* CUDA runtime failure (message in sgi_last_error)                 */
  SGI_ERR_NOMEM = -3,
  SGI_ERR_OVERFLOW = -4,    /* internal bin list overflow: the frame was re-sized, call again   */
  SGI_ERR_NO_DEVICE = -5    /* no CUDA device: this library has no CPU fallback                 */
} sgi_status;

/* Technique selector = which fragment program computeHardShadows()/renderSoftShadows() binds
 * (ShadowMapping/src/main.cpp:400-414, SoftShadowMapping/src/main.cpp:925-1022,756-811) and which
 * ShadowParams bools select the branch inside it. */
typedef enum sgi_technique {
  SGI_TECH_HARD = 0,          /* Shadow.frag, naive                         (Shadow.frag:253-256)            */
  SGI_TECH_PCF = 1,           /* Shadow.frag, PCF float-loop taps           (Shadow.frag:86-116)             */
  SGI_TECH_PCSS = 2,          /* PlausibleSoftShadow.frag, PCSS             (:166-194,365-398,556-563)       */
  SGI_TECH_RBSM_NONCONS = 3,  /* NonConservativeSMSR.frag, SMSR             (:288-304)                       */
  SGI_TECH_RBSM_CONS = 4,     /* ConservativeSMSR.frag, SMSR                (:128-144)                       */
  SGI_TECH_RPCF_NONCONS = 5,  /* NonConservativeSMSR.frag, RPCFPlusSMSR     (:306-349)                       */
  SGI_TECH_RPCF_CONS = 6,     /* ConservativeSMSR.frag, RPCFPlusSMSR        (:146-198)                       */
  SGI_TECH_RSMSS = 7,         /* FilteredRBSM.frag                          (:241,266-272,381-383)           */
  SGI_TECH_MULTI_HARD = 8,    /* AccurateSoftShadow.frag, monteCarlo        (:52-133), N lights              */
  SGI_TECH_RBSSM = 9,         /* SoftShadow/RBSSM.frag: revectorization-based soft shadows (:1202-1376)     */
  SGI_TECH_EDTSM_NONCONS = 10,/* EDT shadow mapping (main.cpp:416-447, EDT/pba2D*.cu|h, MeanFilter.frag) over   */
  SGI_TECH_EDTSM_CONS = 11,   /*   the non-conservative / conservative SMSR hard shadows; whole screen only      */
  /* pre-filtered ("moment") shadow maps of the ShadowMapping program: sgi_render_shadow_map renders the moment target
   * (Shaders/ShadowMap/{Moments,Exponential,ExponentialMoments}.frag, main.cpp:227-243) instead of a depth map,
   * sgi_filter_shadow_map blurs it (main.cpp:374-398), sgi_compute_visibility reconstructs (Shadow.frag:118-220,257-264) */
  SGI_TECH_VSM = 12,          /* variance shadow mapping, chebyshevUpperBound       (Shadow.frag:118-141)            */
  SGI_TECH_ESM = 13,          /* exponential shadow mapping, c = 80, log-space blur (Shadow.frag:144-157)            */
  SGI_TECH_EVSM = 14,         /* exponential variance shadow mapping, c = 60        (Shadow.frag:160-175)            */
  SGI_TECH_MSM = 15,          /* Hamburger 4-moment shadow mapping, quantised       (Shadow.frag:168-220)            */
  SGI_TECH_PCF_TRICUBIC = 16  /* Shadow.frag PCF whose taps are textureBicubic() (tricubicPCF == 1, :41-84,101)          */
} sgi_technique;

typedef enum sgi_depth_func { SGI_DEPTH_LESS = 0, SGI_DEPTH_LEQUAL = 1 } sgi_depth_func;

/* Replaces `struct ShadowParams` (ShadowMapping/include/Viewers/ShadowParams.h:6-37 and the superset
 * SoftShadowMapping/include/Viewers/ShadowParams.h:6-69): the per-frame parameter block every pass reads.
 * The matrices and GL texture names of the original live in sgi_set_camera / sgi_set_lights / the context.
 * Defaults in comments are the reference's (ShadowMapping/src/main.cpp:859-877,
 * SoftShadowMapping/src/main.cpp:1598-1614, ShadowVolumes/src/main.cpp:469). */
typedef struct sgi_params {
  int32_t technique;              /* sgi_technique                                          */
  int32_t shadow_map_width;       /* informational; the authoritative size is sgi_set_lights */
  int32_t shadow_map_height;
  float   shadow_intensity;       /* 0.25                                                   */
  int32_t kernel_order;           /* 7   PCF / RPCF taps per axis (float loop); for VSM / ESM / EVSM / MSM the order
                                     of the separable blur (shadowParams.kernelOrder = gaussianFilter->getOrder(), main.cpp:321) */
  int32_t penumbra_size;          /* 1                                                      */
  int32_t blocker_search_size;    /* 7   PCSS                                               */
  int32_t kernel_size;            /* 15  PCSS                                               */
  int32_t light_source_radius;    /* 8   PCSS                                               */
  int32_t max_search;             /* 16  RBSM                                               */
  float   depth_threshold;        /* Configs `d` line                                       */
  int32_t z_near, z_far;          /* 1, 1000 (`uniform int zNear/zFar`)                     */
  float   polygon_offset_factor;  /* 4   glPolygonOffset(4, 20), main.cpp:246               */
  float   polygon_offset_units;   /* 20                                                     */
  int32_t sv_depth_func;          /* SGI_DEPTH_LEQUAL (steady state of ShadowVolumes/src/main.cpp:174) */
  int32_t sv_infinity;            /* 100                                                    */
  int32_t rect_x0, rect_y0, rect_x1, rect_y1;  /* screen rectangle this context evaluates in
                                     sgi_compute_visibility / shadow-volume counting (multi-GPU
                                     screen tiles); an empty rectangle means the whole screen */
  int32_t multi_partial;          /* SGI_TECH_MULTI_HARD on a light shard: 1 = write the un-normalised sum
                                     over this context's lights (ranks are summed, then divided by the total);
                                     2 (with multi_fused, at most 32 lights in the whole set) = write which of this
                                     context's lights reach the pixel, one bit per light at its index in the whole set
                                     (sgi_set_light_ids) into SGI_BUF_LIGHT_MASK - a quarter to a half of the bytes to
                                     exchange, and the strip's visibility is then accumulated in the reference's light
                                     order: the bits of the un-sharded frame for every shadow intensity */
  int32_t multi_fused;            /* SGI_TECH_MULTI_HARD: 1 = the accumulation kernel resolves each pixel's world position itself from
                                     SGI_BUF_PRIM_ID (sgi_render_prim_ids) instead of reading a materialised G-buffer: the same
                                     positions to the bit, without the 16 B/pixel write and read of the position target */
  int32_t sv_silhouette;          /* shadow volumes: 0 = one open prism per triangle, as ShadowVolume::update builds them (parity mode);
                                     1 = side quads of edges shared by two triangles of the same orientation class are dropped in
                                     pairs (they cancel +1/-1): silhouette, boundary and non-manifold edges only */
  int32_t sv_zfail;               /* shadow volumes: 0 = depth-pass counting (the reference's stencil ops, main.cpp:166-168),
                                     1 = depth-fail counting over capped volumes with depth clamp (robust when the eye is inside a volume) */
} sgi_params;

typedef enum sgi_buffer {
  SGI_BUF_SHADOW_MAP = 0,   /* float   [N][Sh][Sw]  light-view window depth, cleared to 1.0          */
  SGI_BUF_GBUF_POS = 1,     /* float4  [H][W]       (world x,y,z,1); background (0,0,0,1)             */
  SGI_BUF_GBUF_NRM = 2,     /* float4  [H][W]       (object normal xyz, gl_FrontFacing); bg (0,0,0,1) */
  SGI_BUF_CAM_DEPTH = 3,    /* float   [H][W]       camera-view window depth, cleared to 1.0          */
  SGI_BUF_VISIBILITY = 4,   /* float   [H][W]       1 = lit, shadow_intensity = shadowed, bg 0        */
  SGI_BUF_SV_COUNT = 5,     /* int32   [H][W]       signed z-pass count                               */
  SGI_BUF_SV_STENCIL = 6,   /* uint8   [H][W]       count mod 256 (what the 8-bit stencil holds)      */
  SGI_BUF_SV_PRISM_XYZ = 7, /* float   [6T][3]      ShadowVolume::update vertices                     */
  SGI_BUF_SV_PRISM_IDX = 8, /* int32   [6T][3]      ShadowVolume::update indices ([8T][3] with sv_zfail: + near cap, far cap per triangle;
                                                    sv_silhouette: the two triangles of a dropped side quad read (0,0,0))          */
  SGI_BUF_GBUF_ALBEDO = 9,  /* float4  [H][W]       (vertex colour rgb, 1) when colours are set; bg (0,0,0,1)  */
  SGI_BUF_SHADED = 10,      /* float4  [H][W]       deferred Phong image; background = the clear colour        */
  SGI_BUF_EDT_NEAREST = 11, /* int16x2 [H][W]       EDT shadow mapping: nearest shadow-boundary pixel (x, y), -32768 = none */
  SGI_BUF_MOMENTS = 12,     /* float4  [Sh][Sw]     VSM/ESM/EVSM/MSM: moment target of the light-view pass, cleared to (0,0,0,1) */
  SGI_BUF_MOMENTS_X = 13,   /* float4  [H][W]       filterShadowMap: after the horizontal pass (FILTER_X_MAP_COLOR, window-sized) */
  SGI_BUF_MOMENTS_FILTERED = 14, /* float4 [H][W]   filterShadowMap: after the vertical pass (FILTER_Y_MAP_COLOR), what Shadow.frag samples */
  SGI_BUF_PRIM_ID = 15,     /* uint32  [H][W]       sgi_render_prim_ids: winning primitive per pixel = source triangle * 8 + fan index of its clipped
                                                    polygon; 0xFFFFFFFF = background */
  SGI_BUF_LIGHT_MASK = 16,  /* uint8   [ranks][ceil(lights/8)][strip rows][W]  params.multi_partial == 2: lit lights per pixel, 8 per byte plane */
  SGI_BUF_COUNT_ = 17
} sgi_buffer;

/* passes that can be timed with sgi_pass_time_ms */
typedef enum sgi_pass {
  SGI_PASS_SHADOW_MAP = 0, SGI_PASS_GBUFFER = 1, SGI_PASS_VISIBILITY = 2, SGI_PASS_SHADOW_VOLUME = 3,
  SGI_PASS_VIS_KERNEL = 4,  /* only the per-pixel shadow kernel inside SGI_PASS_VISIBILITY             */
  SGI_PASS_TILE_DEPTH = 5,  /* only the per-tile raster kernel of the light-view depth pass (per light)  */
  SGI_PASS_TILE_GBUFFER = 6,/* only the per-tile raster+resolve kernel of the G-buffer pass              */
  SGI_PASS_TILE_SV = 7,     /* only the per-tile counting kernel of the shadow-volume pass               */
  SGI_PASS_MOMENT_FILTER = 8, /* both blur passes of sgi_filter_shadow_map                                  */
  SGI_PASS_COUNT_ = 9
} sgi_pass;

/* lifecycle — replaces initGL()'s FBO/texture/VBO creation (ShadowMapping/src/main.cpp:839-953) */
int sgi_create(sgi_ctx** out, int device);
int sgi_destroy(sgi_ctx* ctx);
/* run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL = the context's own */
int sgi_set_stream(sgi_ctx* ctx, void* cuda_stream);

/* geometry — replaces MyGLGeometryViewer::loadVBOs (ShadowMapping/src/Viewers/MyGLGeometryViewer.cpp:383-405);
 * arguments are the Mesh getters getPointCloud/getNormalVector/getIndices (ShadowMapping/include/Mesh.h:35-48).
 * Uploaded once and kept resident (the reference re-uploads on every draw). */
int sgi_set_mesh(sgi_ctx* ctx, const float* xyz, const float* nrm, int32_t num_vertices,
                 const int32_t* idx, int32_t num_triangles);

/* per-vertex colours — Mesh::getColors() (`c r g b` / `cf` directives), the `color` attribute of GBuffer.vert:14,21 with
 * useMeshColor == 1 (MyGLGeometryViewer.cpp:295-296).  rgb = 3 floats per vertex of the current mesh, or NULL to stop
 * writing SGI_BUF_GBUF_ALBEDO (shading then uses white).  Call after sgi_set_mesh. */
int sgi_set_mesh_colors(sgi_ctx* ctx, const float* rgb);

/* texture coordinates and scene textures — the `uv` attribute of GBuffer.vert (Mesh::getTextureCoords: u, v and in the third component
 * the 1-based id of the object's texture, set by the `m` directive through Mesh::loadTexture, Mesh.cpp:315-325) and the textures
 * MyGLTextureViewer::loadRGBTexture creates (MyGLTextureViewer.cpp:45-56: RGB8, GL_LINEAR, GL_REPEAT, no mipmaps; index 0..2 =
 * texture0..2 of GBuffer.frag; rgb rows as cv::Mat stores them, row 0 = t 0).  With both set the G-buffer pass runs the texture
 * select of GBuffer.frag:11-30 (useTextureForColoring): SGI_BUF_GBUF_ALBEDO = bilinear texel where uv.z picks a texture, the
 * vertex colour (or 0 without colours) elsewhere.  uv == NULL / rgb == NULL switch it off again.  Call after sgi_set_mesh. */
int sgi_set_mesh_uv(sgi_ctx* ctx, const float* uv);
int sgi_set_texture(sgi_ctx* ctx, int32_t index, const uint8_t* rgb, int32_t width, int32_t height);

/* camera uniforms — replaces configureAmbient + configurePhong for the camera view
 * (MyGLGeometryViewer.cpp:14-19,108-134): MVP, MV, frozen normalMatrix, window size. */
int sgi_set_camera(sgi_ctx* ctx, const float mvp[16], const float mv[16], const float normal_matrix[9],
                   int32_t width, int32_t height);

/* light uniforms — replaces displaySceneFromLightPOV's lightMVP (ShadowMapping/src/main.cpp:249-266) and
 * configureShadow's bias multiply (MyGLGeometryViewer.cpp:136-186) done by the caller:
 *   light_mvp        N x 16, un-biased  (rasterised by sgi_render_shadow_map)
 *   light_mvp_biased N x 16, bias*lightMVP (sampled by sgi_compute_visibility; for SGI_TECH_MULTI_HARD
 *                    the 3x4 part of the LAST one is the shader's common term and column 3 of each is
 *                    lightMVPTrans[i], SoftShadowMapping/src/Viewers/MyGLGeometryViewer.cpp:184-196)
 *   light_pos_shading  the `lightPosition` uniform: light eye rotated 180 deg about Y (main.cpp:283) */
int sgi_set_lights(sgi_ctx* ctx, int32_t num_lights, const float* light_mvp, const float* light_mvp_biased,
                   const float light_pos_shading[3], int32_t map_width, int32_t map_height);

/* many-light shards: the shader's common 3x4 term comes from the LAST light of the whole set
 * (SoftShadowMapping/src/main.cpp:790,806); a rank that owns a subset passes that matrix (bias*lightMVP) here.
 * NULL restores the default (last light given to sgi_set_lights). */
int sgi_set_multi_light_common(sgi_ctx* ctx, const float light_mvp_biased[16]);

int sgi_set_params(sgi_ctx* ctx, const sgi_params* params);
void sgi_default_params(sgi_params* params);

/* passes */
int sgi_render_shadow_map(sgi_ctx* ctx);      /* renderShadowMap(), ShadowMapping/src/main.cpp:350-361 (all N lights) */
int sgi_render_gbuffer(sgi_ctx* ctx);         /* renderGBuffer(),   ShadowMapping/src/main.cpp:363-372               */
/* The camera pass reduced to what AccurateSoftShadow.frag consumes (it reads only the vertex map; the normal test is commented
 * out, :59-60): the same rasterisation as sgi_render_gbuffer, but only the winning primitive of every pixel of the rectangle is
 * stored (4 B/pixel, SGI_BUF_PRIM_ID).  With params.multi_fused the many-light pass interpolates the position from it on the fly.
 * On a light shard each rank rasterises its screen strip and the strips are exchanged with sgi_gather. */
int sgi_render_prim_ids(sgi_ctx* ctx);
/* filterShadowMap(), ShadowMapping/src/main.cpp:374-398 (display() calls it between renderShadowMap and renderGBuffer when
 * VSM / ESM / EVSM / MSM is on, :471): the separable binomial blur of order params.kernel_order (Filter::buildGaussianKernel,
 * src/Filter.cpp:17-46) - GaussianFilter.frag, or LogGaussianFilter.frag for ESM - from the moment target into the two
 * window-sized targets.  Needs sgi_render_shadow_map with the same technique and sgi_set_camera (window size) first. */
int sgi_filter_shadow_map(sgi_ctx* ctx);
/* MyGLGeometryViewer::configureMoments (ShadowMapping/src/Viewers/MyGLGeometryViewer.cpp:188-213): the MSM quantisation
 * uniforms mQuantization, mQuantizationInverse (column-major) and tQuantization, computed in fp32 in GLM's operation order.
 * Pure host arithmetic (no device needed); the passes use the same values internally. */
void sgi_moment_quantization(float m[16], float m_inverse[16], float t[4]);
int sgi_compute_visibility(sgi_ctx* ctx);     /* computeHardShadows() :400-414 / renderSoftShadows()
                                                 SoftShadowMapping/src/main.cpp:925-1022 / renderMonteCarlo() :756-811 */
/* ShadowVolume::update (ShadowVolumes/src/ShadowVolume.cpp:116-195) + the stencil pass of display()
 * (ShadowVolumes/src/main.cpp:154-172).  light_pos = un-rotated light eye.  Needs sgi_render_gbuffer first
 * (its SGI_BUF_CAM_DEPTH is the depth pre-pass). */
int sgi_compute_shadow_volume(sgi_ctx* ctx, const float light_pos[3]);
/* with the option "sv_count_fragments" = 1: how many prism fragments (pixel centres covered by a volume triangle inside the
 * rectangle, before the depth test) the last sgi_compute_shadow_volume visited - the unit of work of the stencil pass */
int sgi_sv_fragments(sgi_ctx* ctx, int64_t* fragments);
/* Diagnostic (no reference counterpart): the shadow pass shares one reciprocal between the three divisions of a projective
 * divide (`shadowCoord / shadowCoord.w`, Shadow.frag:241); this runs that sequence and the plain IEEE division on n pseudo-random
 * operand quadruples (every exponent, denormals, infinities, NaNs) and returns how many quotients differ in any bit - 0. */
int sgi_divide_selftest(sgi_ctx* ctx, uint64_t n, uint32_t seed, uint64_t* mismatches);

/* shadeScene(), ShadowMapping/src/main.cpp:449-457: deferred Phong shading of the G-buffer with the visibility buffer as
 * hardShadowMap (PhongShading.frag:11-47) into SGI_BUF_SHADED; clear_rgba = glClearColor (0.63, 0.82, 0.96, 1). */
int sgi_shade_phong(sgi_ctx* ctx, const float clear_rgba[4]);

/* results */
int sgi_read(sgi_ctx* ctx, int32_t which, void* host_dst, size_t bytes);          /* blocking D2H */
int sgi_device_ptr(sgi_ctx* ctx, int32_t which, void** device_ptr, size_t* bytes);/* borrowed     */
/* Non-blocking readback for frame pipelining: the copy runs on the context's copy stream as soon as the buffer's
 * producers have finished, while later passes (the next frame) proceed; a pass that would overwrite the buffer waits
 * for the copy on the device.  host_dst should be page-locked (sgi_alloc_host).  sgi_read_wait blocks the host until
 * the copy with that ticket has landed (up to 4 may be outstanding). */
int sgi_read_async(sgi_ctx* ctx, int32_t which, void* host_dst, size_t bytes, int32_t* ticket);
int sgi_read_wait(sgi_ctx* ctx, int32_t ticket);   /* SGI_ERR_OVERFLOW: a tile list overflowed in this frame or in one issued after it (the
                                                      frames are pipelined): every ticket issued before the overflow was seen reports it; issue
                                                      those frames again - the lists have been re-sized */
int sgi_synchronize(sgi_ctx* ctx);
/* Orders the context's stream after every pass queued so far (the G-buffer and shadow passes run on internal streams and
 * overlap the next frame's passes); does not block the host.  sgi_read*, sgi_device_ptr, sgi_shade_phong and
 * sgi_synchronize do this themselves. */
int sgi_join(sgi_ctx* ctx);

/* ---- multi-GPU (one context per GPU / process rank; SURVEY 8e: screen tiles, and lights for many-light scenes) ------------
 * NCCL over NVLink / NVSwitch, loaded at run time (libnccl.so.2); nothing here is needed on one GPU.
 *   sgi_comm_unique_id  rank 0 makes the 128-byte NCCL id; the caller ships it to the other ranks (MPI, torch.distributed, a file)
 *   sgi_comm_init       every rank joins; screen strips = ceil(H / nranks) rows per rank (sgi_comm_strip), the exchanged buffers
 *                       are padded to nranks equal strips so that every exchange is ONE collective, in place, with no packing
 *   sgi_gather          tile sharding: every rank has produced its strip of `which` (rect_* = its strip) -> ncclAllGather in place;
 *                       afterwards every rank holds the whole buffer (visibility, shadow-volume counts, primitive ids, ...)
 *   sgi_reduce_lights   light sharding (renderMonteCarlo, AccurateSoftShadow.frag:127): the ranks' un-normalised partial sums in
 *                       SGI_BUF_VISIBILITY (params.multi_partial) -> ncclReduceScatter in place, then divided by the number of
 *                       lights of the whole set: rank r ends with the final visibility of ITS strip (what tile-local shading
 *                       consumes); rows outside the strip are left as partial sums.  With params.multi_partial == 2 the ranks'
 *                       lit masks (SGI_BUF_LIGHT_MASK, disjoint bits) are reduce-scattered instead - 1 B/pixel per 8 lights -
 *                       and the strip's visibility is accumulated from the union in the light order of the whole set
 *   sgi_set_light_ids   light sharding with masks: the index in the whole set of each light passed to sgi_set_lights, and the
 *                       size of the whole set (<= 32)
 * The collectives run on the context's communication stream behind the passes that produce their input and ahead of the passes
 * that consume their output (device-side ordering, the host does not block). */
int sgi_comm_unique_id(void* id128, size_t bytes);
int sgi_comm_init(sgi_ctx* ctx, const void* id128, size_t bytes, int32_t rank, int32_t nranks);
int sgi_comm_destroy(sgi_ctx* ctx);
int sgi_comm_strip(sgi_ctx* ctx, int32_t rank, int32_t* row0, int32_t* row1);
int sgi_gather(sgi_ctx* ctx, int32_t which);
int sgi_reduce_lights(sgi_ctx* ctx, int32_t total_lights);
int sgi_set_light_ids(sgi_ctx* ctx, int32_t n, const int32_t* ids, int32_t total_lights);

/* page-locked host memory for callers that want sgi_set_mesh / sgi_read to be true async DMA
 * (the reference keeps its Mesh arrays in malloc'd memory and lets the GL driver stage them) */
int sgi_alloc_host(void** host_ptr, size_t bytes);
int sgi_free_host(void* host_ptr);
/* page-lock / release caller-owned arrays in place (e.g. the Mesh's own vectors).  With the option "borrow_pinned" = 1,
 * sgi_set_mesh / sgi_set_mesh_colors read page-locked inputs by DMA after the call has returned instead of copying them
 * inside it: the caller then keeps those arrays unchanged until the frame that uses them has completed
 * (sgi_synchronize, sgi_read, sgi_read_wait). */
int sgi_register_host(void* host_ptr, size_t bytes);
int sgi_unregister_host(void* host_ptr);

/* implementation switches (experiments / A-B measurements; results are identical either way):
 *   "vis_staged"      0 (default) taps through L1/L2, 1 = PCF/PCSS stage the CTA's shadow-map window in shared memory
 *   "overlap_passes"  1 (default) G-buffer pass on the auxiliary stream, 0 = everything on the main stream
 *   "tile_threads"    0 (default) automatic, or 256 / 512 / 1024 threads per tile CTA
 *   "tile_order"      1 (default) tile CTAs are launched busiest tile first, 0 = in raster order
 *   "tile_split"      subdivision threshold of hot tiles in list records (default 256, 0 = off)
 *   "rbssm_compact"   1 (default) RBSSM as work list + one warp per penumbra pixel, 0 = one thread per pixel
 *   "pcss_early_out"  0 (default) every PCSS pixel runs its blocker search as the shader does, 1 = pixels whose light-space depth
 *                     is in (0, 0.989) return 1.0 without a tap: provably what the program computes there (the 0.99 cut-off of
 *                     PlausibleSoftShadow.frag:368), so results stay bit-identical; off by default so that timings count the taps
 *   "vis_minmax_cull" 0 (default); 1 = PCF / PCSS with one light: the depth pass also records the extrema of every 32x32-texel block
 *                     of the map; a pixel nearer than every depth its tap window can reach (or beyond all of them) is decided
 *                     without taps - the same value the tap loop produces, bit for bit.  Off by default: measured on the bench
 *                     scenes it decides few windows (a surface sloped against the light blocks itself within the window's reach)
 *                     and costs more than it saves (profiles/r2_experiments.txt)
 *   "sv_count_fragments" 0 (default); 1 = shadow-volume passes tally their fragments (sgi_sv_fragments; costs a little)
 *   "pdl"             1 (default) k_order and the tile kernel are launched as programmatic dependents (their launch latency overlaps the
 *                     predecessor's tail; griddepcontrol.wait in the kernels), 0 = plain stream order
 *   "tile_bulk_flush" 1 (default) depth tiles leave shared memory by cp.async.bulk row copies, 0 = by 16-byte stores
 *   "tile_direct"     depth tiles whose triangle list has at most this many entries (default 32, at most 128, 0 = never) are
 *                       rasterised in registers, one 4 x 4 texel patch per thread, and stored straight to the map
 *   "tile_bin_big"    passes of at least this many tiles (default 4096, 0 = never) bin the records that span more than 16 tiles
 *                       in a kernel of their own (k_bin_big: a warp per record, all CTAs together on the largest) when the last
 *                       pass of the kind held enough records beyond 256 tiles - records x tiles >= "tile_bin_big_work" (default
 *                       2^20); otherwise every tile tests those records itself
 *   "comm_split"      0 (default); 1 (before sgi_comm_init) = sgi_reduce_lights runs on a second communicator and stream (ncclCommSplit)
 *   "tile_few_walk"   1 (default) passes of at most 128 tiles bin their larger records tile-major, one list atomic per warp and tile
 *                       (the cursors of a few dozen tiles are contended by the whole GPU), 0 = pair by pair
 *   "sv_split_lists"  1 (default) hot tiles of the stencil pass are shared by list segment (every CTA counts its part of the
 *                       list over the whole tile, counts are added atomically), 0 = by sub-region as in the depth passes
 *   "tile_static_items" work items of a tile's triangle list are dealt to the warps round robin (1), drawn from a shared cursor (0),
 *                       or either by the list's length (2, default: static for lists of up to 48 triangles and for stencil counting)
 *   "tile_refresh_full" the per-block depth bound is refreshed after fully covered blocks only (1), after every block (0), or by the
 *                       list's length (2, default).  Scheduling only: results are identical for every value of both options.
 *   "sv_tile_cull"    1 (default) shadow volumes: (prism, tile) pairs behind the tile's farthest scene depth are not listed
 *   "borrow_pinned"   0 (default) inputs are copied inside the call, 1 = page-locked inputs are read later by DMA */
int sgi_set_option(sgi_ctx* ctx, const char* name, int32_t value);

/* instrumentation */
int sgi_enable_timing(sgi_ctx* ctx, int32_t on);                 /* CUDA events around every pass            */
int sgi_pass_time_ms(sgi_ctx* ctx, int32_t pass, double* total_ms, int64_t* calls);  /* since last reset      */
int sgi_reset_timing(sgi_ctx* ctx);
int sgi_kernel_launches(sgi_ctx* ctx, int64_t* launches);        /* kernels this context has launched so far */
const char* sgi_last_error(sgi_ctx* ctx);
const char* sgi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SHADOWGI_H */
