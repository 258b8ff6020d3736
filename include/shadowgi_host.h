/*
 * shadowgi_host.h — C entry points of the host side (libshadowgi_host.so): the reference's scene loading
 * (SceneLoader/Mesh/OBJ, ShadowMapping/src/IO/SceneLoader.cpp, src/Mesh.cpp, src/IO/OBJLoader.cpp), its matrix
 * set-up (MyGLGeometryViewer.cpp:14-19,108-186) and its per-technique render-pass interface
 * (display()/renderShadowMap()/renderGBuffer()/computeHardShadows()/renderSoftShadows()/renderMonteCarlo(),
 * ShadowVolumes display()) re-typed as portable C++17 over the C ABI of shadowgi.h.  These wrappers exist so
 * tests and the bench driver can call the C++ classes (globalillumination_b200/host/) through ctypes.
 */
#ifndef SHADOWGI_HOST_H
#define SHADOWGI_HOST_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgh_scene sgh_scene;
typedef struct sgh_app sgh_app;

enum { SGH_PROGRAM_SHADOW_MAPPING = 0, SGH_PROGRAM_SOFT_SHADOW_MAPPING = 1, SGH_PROGRAM_SHADOW_VOLUMES = 2 };

const char* sgh_last_error(void);

/* SceneLoader::load on a Configs .txt file; paths inside are resolved against base_dir. NULL on error. */
sgh_scene* sgh_scene_load(const char* config, const char* base_dir);
void sgh_scene_free(sgh_scene* s);
int sgh_scene_counts(sgh_scene* s, int32_t* num_vertices, int32_t* num_triangles);
int sgh_scene_copy(sgh_scene* s, float* xyz, float* nrm, int32_t* idx);
int sgh_scene_views(sgh_scene* s, float* cam_eye, float* cam_at, float* light_eye, float* light_at, float* depth_threshold);
const char* sgh_scene_substitutions(sgh_scene* s);   /* ';'-separated list of missing assets replaced by procedural stand-ins */

/* one frame's uniforms as display() derives them (no user transform, no animation) */
int sgh_frame_matrices(const float* cam_eye, const float* cam_at, const float* light_eye, const float* light_at, int32_t W, int32_t H,
                       int32_t SW, int32_t SH, float* cam_mvp, float* cam_mv, float* normal_matrix9, float* light_mvp,
                       float* light_mvp_biased, float* light_pos_shading);

/* the application object: one per GPU */
sgh_app* sgh_app_create(int32_t device);
void sgh_app_destroy(sgh_app* a);
const char* sgh_app_error(sgh_app* a);
void* sgh_app_context(sgh_app* a);                    /* the underlying sgi_ctx* */
int sgh_app_load_scene(sgh_app* a, const char* config, const char* base_dir);
int sgh_app_set_scene(sgh_app* a, const float* xyz, const float* nrm, int32_t nv, const int32_t* idx, int32_t nt, const float* cam_eye,
                      const float* cam_at, const float* light_eye, const float* light_at, float depth_threshold);
int sgh_app_scene_counts(sgh_app* a, int32_t* nv, int32_t* nt);
int sgh_app_scene_copy(sgh_app* a, float* xyz, float* nrm, int32_t* idx);
int sgh_app_configure(sgh_app* a, int32_t W, int32_t H, int32_t SW, int32_t SH);
int sgh_app_set_rect(sgh_app* a, int32_t x0, int32_t y0, int32_t x1, int32_t y1);
int sgh_app_set_light_shard(sgh_app* a, int32_t rank, int32_t world);   /* many-light: own lights l = rank (mod world) */
/* multi-GPU inside the library: sgi_comm_init on the app's context (id from sgi_comm_unique_id on rank 0); renderMonteCarlo then
 * shards the lights over the ranks and exchanges primitive-id strips / partial sums over NCCL itself (sgi_gather, sgi_reduce_lights) */
/* scene texture of an `m` directive (Mesh::loadTexture): the reference decodes the image with OpenCV, which the host side here does not
 * link; the caller supplies the decoded RGB8 pixels for texture<index> of GBuffer.frag (index = the directive's running number - 1) */
int sgh_app_set_texture(sgh_app* a, int32_t index, const uint8_t* rgb, int32_t width, int32_t height);
int sgh_app_comm_init(sgh_app* a, const void* id128, size_t bytes, int32_t rank, int32_t world);
/* light shards balanced by cost: the depth pass of a light costs what the light sees; ms[n] = each light's pass time (one at a time,
 * CUDA events), owner[n] = the rank that renders light s (the same table on every rank; default s mod world) */
int sgh_app_light_costs(sgh_app* a, float* ms, int32_t n);
int sgh_app_set_light_owners(sgh_app* a, const int32_t* owner, int32_t n);
int sgh_app_set_technique(sgh_app* a, const char* name);
int sgh_app_set_int(sgh_app* a, const char* name, int32_t v);
int sgh_app_set_float(sgh_app* a, const char* name, float v);
int sgh_app_upload_scene(sgh_app* a);
int sgh_app_render_shadow_map(sgh_app* a);
int sgh_app_render_gbuffer(sgh_app* a);
int sgh_app_filter_shadow_map(sgh_app* a);       /* filterShadowMap(), ShadowMapping/src/main.cpp:374-398 (VSM/ESM/EVSM/MSM) */
int sgh_app_compute_hard_shadows(sgh_app* a);
int sgh_app_render_soft_shadows(sgh_app* a);
int sgh_app_render_monte_carlo(sgh_app* a);
int sgh_app_render_shadow_volumes(sgh_app* a);
int sgh_app_shade_scene(sgh_app* a);              /* shadeScene(): deferred Phong of the last frame */
int sgh_app_save_image(sgh_app* a, const char* path);   /* shadeScene() + the frame as an 8-bit RGBA PNG (top-down)    */
int sgh_write_png(const char* path, const uint8_t* rgba, int32_t W, int32_t H);   /* the encoder alone (no GPU needed) */
int sgh_app_display(sgh_app* a, int32_t program);
int sgh_app_display_e2e(sgh_app* a, int32_t program, int32_t result_buffer, void* host_dst, size_t bytes);
int sgh_app_display_e2e_async(sgh_app* a, int32_t program, int32_t result_buffer, void* host_dst, size_t bytes, int32_t* ticket);
int sgh_app_e2e_wait(sgh_app* a, int32_t ticket);
int sgh_app_step_animation(sgh_app* a, float delta);

/* procedural stand-ins (malloc'd arrays, release with sgh_free) */
int sgh_procedural(const char* spec, float** xyz, int32_t* nv, int32_t** idx, int32_t* nt);
void sgh_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
