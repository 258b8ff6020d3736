// host_capi.cpp — C entry points of libshadowgi_host.so (declared in include/shadowgi_host.h) so that the
// host side (scene loader, matrices, render-pass interface) can be driven from tests/ and bench.py via ctypes.
#include <cstring>
#include <string>

#include "../../include/shadowgi_host.h"
#include "procedural.h"
#include "shadow_app.h"
#include "png_writer.h"

using namespace sgh;

struct sgh_scene { Mesh mesh; float cam_eye[3], cam_at[3], light_eye[3], light_at[3]; float depth_threshold; std::string err; std::string subs; };
struct sgh_app { ShadowApp app; explicit sgh_app(int d) : app(d) {} };

static thread_local std::string g_err;

extern "C" {

const char* sgh_last_error(void) { return g_err.c_str(); }

sgh_scene* sgh_scene_load(const char* config, const char* base_dir) {
  sgh_scene* s = new sgh_scene();
  SceneLoader loader(config, &s->mesh);
  int rc = loader.load(base_dir ? base_dir : "");
  if (rc) { g_err = loader.error(); delete s; return nullptr; }
  for (int a = 0; a < 3; a++) {
    s->cam_eye[a] = loader.getCameraPosition()[a]; s->cam_at[a] = loader.getCameraAt()[a];
    s->light_eye[a] = loader.getLightPosition()[a]; s->light_at[a] = loader.getLightAt()[a];
  }
  s->depth_threshold = loader.getDepthThreshold();
  for (const std::string& x : loader.substitutions()) s->subs += x + ";";
  return s;
}
void sgh_scene_free(sgh_scene* s) { delete s; }
int sgh_scene_counts(sgh_scene* s, int32_t* nv, int32_t* nt) {
  if (!s) return -1;
  *nv = s->mesh.getPointCloudSize() / 3; *nt = s->mesh.getNumberOfTriangles();
  return 0;
}
int sgh_scene_copy(sgh_scene* s, float* xyz, float* nrm, int32_t* idx) {
  if (!s) return -1;
  std::memcpy(xyz, s->mesh.getPointCloud(), sizeof(float) * s->mesh.getPointCloudSize());
  std::memcpy(nrm, s->mesh.getNormalVector(), sizeof(float) * s->mesh.getPointCloudSize());
  std::memcpy(idx, s->mesh.getIndices(), sizeof(int) * s->mesh.getIndicesSize());
  return 0;
}
int sgh_scene_views(sgh_scene* s, float* cam_eye, float* cam_at, float* light_eye, float* light_at, float* depth_threshold) {
  if (!s) return -1;
  std::memcpy(cam_eye, s->cam_eye, 12); std::memcpy(cam_at, s->cam_at, 12);
  std::memcpy(light_eye, s->light_eye, 12); std::memcpy(light_at, s->light_at, 12);
  *depth_threshold = s->depth_threshold;
  return 0;
}
const char* sgh_scene_substitutions(sgh_scene* s) { return s ? s->subs.c_str() : ""; }

int sgh_frame_matrices(const float* cam_eye, const float* cam_at, const float* light_eye, const float* light_at, int32_t W, int32_t H,
                       int32_t SW, int32_t SH, float* cam_mvp, float* cam_mv, float* normal_matrix9, float* light_mvp,
                       float* light_mvp_biased, float* light_pos_shading) {
  // a context-free ShadowApp cannot be built without a GPU, so compose through a scratch object's maths only
  struct Scratch : ShadowApp { Scratch() : ShadowApp(-1) {} };
  Scratch a;
  a.cameraEye = Vec3{cam_eye[0], cam_eye[1], cam_eye[2]}; a.cameraAt = Vec3{cam_at[0], cam_at[1], cam_at[2]};
  a.lightPositionConfig = Vec3{light_eye[0], light_eye[1], light_eye[2]}; a.lightAt = Vec3{light_at[0], light_at[1], light_at[2]};
  a.setWindowSize(W, H); a.setShadowMapSize(SW, SH);
  FrameMatrices f = a.frameMatrices();
  std::memcpy(cam_mvp, f.cameraMVP.m, 64); std::memcpy(cam_mv, f.cameraMV.m, 64); std::memcpy(normal_matrix9, f.normalMatrix.m, 36);
  std::memcpy(light_mvp, f.lightMVP.m, 64); std::memcpy(light_mvp_biased, f.lightMVPBiased.m, 64);
  std::memcpy(light_pos_shading, &f.lightPositionShading.x, 12);
  return 0;
}

sgh_app* sgh_app_create(int32_t device) {
  sgh_app* a = new sgh_app(device);
  if (!a->app.ok()) { g_err = a->app.error(); delete a; return nullptr; }
  return a;
}
void sgh_app_destroy(sgh_app* a) { delete a; }
const char* sgh_app_error(sgh_app* a) { return a ? a->app.error().c_str() : "null app"; }
void* sgh_app_context(sgh_app* a) { return a ? (void*)a->app.context() : nullptr; }

int sgh_app_load_scene(sgh_app* a, const char* config, const char* base_dir) { return a ? a->app.loadScene(config, base_dir) : -1; }
int sgh_app_set_scene(sgh_app* a, const float* xyz, const float* nrm, int32_t nv, const int32_t* idx, int32_t nt, const float* cam_eye,
                      const float* cam_at, const float* light_eye, const float* light_at, float depth_threshold) {
  return a ? a->app.setScene(xyz, nrm, nv, idx, nt, cam_eye, cam_at, light_eye, light_at, depth_threshold) : -1;
}
int sgh_app_scene_counts(sgh_app* a, int32_t* nv, int32_t* nt) {
  if (!a) return -1;
  *nv = a->app.getScene()->getPointCloudSize() / 3; *nt = a->app.getScene()->getNumberOfTriangles();
  return 0;
}
int sgh_app_scene_copy(sgh_app* a, float* xyz, float* nrm, int32_t* idx) {
  if (!a) return -1;
  Mesh* m = a->app.getScene();
  std::memcpy(xyz, m->getPointCloud(), sizeof(float) * m->getPointCloudSize());
  std::memcpy(nrm, m->getNormalVector(), sizeof(float) * m->getPointCloudSize());
  std::memcpy(idx, m->getIndices(), sizeof(int) * m->getIndicesSize());
  return 0;
}
int sgh_app_configure(sgh_app* a, int32_t W, int32_t H, int32_t SW, int32_t SH) {
  if (!a || W <= 0 || H <= 0 || SW <= 0 || SH <= 0) return -1;
  a->app.setWindowSize(W, H); a->app.setShadowMapSize(SW, SH);
  return 0;
}
int sgh_app_set_rect(sgh_app* a, int32_t x0, int32_t y0, int32_t x1, int32_t y1) {
  if (!a) return -1;
  a->app.rect[0] = x0; a->app.rect[1] = y0; a->app.rect[2] = x1; a->app.rect[3] = y1;
  return 0;
}

int sgh_app_set_light_shard(sgh_app* a, int32_t rank, int32_t world) {
  if (!a || world <= 0 || rank < 0 || rank >= world) return -1;
  a->app.lightShardRank = rank; a->app.lightShardWorld = world;
  return 0;
}

// multi-GPU: join the NCCL communicator (id from sgi_comm_unique_id on rank 0, shipped by the caller); the many-light frame is
// then sharded by lights with the exchanges done by the library (ShadowApp::renderMonteCarlo)
int sgh_app_comm_init(sgh_app* a, const void* id128, size_t bytes, int32_t rank, int32_t world) {
  if (!a) return -1;
  int rc = a->app.commInit(id128, bytes, rank, world);
  if (rc) g_err = a->app.error();
  return rc;
}

// many-light shards balanced by cost: time each light's depth pass (sgh_app_light_costs), then say which rank owns which light
// (sgh_app_set_light_owners; every rank must be given the same table)
int sgh_app_light_costs(sgh_app* a, float* ms, int32_t n) {
  if (!a || !ms) return -1;
  int rc = a->app.measureLightCosts(ms, n);
  if (rc) g_err = a->app.error();
  return rc;
}
int sgh_app_set_light_owners(sgh_app* a, const int32_t* owner, int32_t n) {
  if (!a || n < 0 || (n > 0 && !owner)) return -1;
  a->app.lightOwner.assign(owner, owner + n);
  return 0;
}

// technique names = the reference's menu entries / ShadowParams flags
int sgh_app_set_technique(sgh_app* a, const char* name) {
  if (!a || !name) return -1;
  ShadowParams& p = a->app.shadowParams;
  p.naive = p.SMSR = p.RPCFPlusSMSR = p.RSMSS = p.RPCFPlusRSMSS = p.EDTSM = p.conservative = false;
  p.VSM = p.ESM = p.EVSM = p.MSM = p.tricubicPCF = false;
  p.bilinearPCF = true; p.PCSS = true; p.monteCarlo = false; p.RBSSM = false;
  std::string n(name);
  if (n == "naive" || n == "hard") p.naive = true;
  else if (n == "pcf" || n == "bilinearPCF") {}
  else if (n == "pcss" || n == "PCSS") p.PCSS = true;
  else if (n == "smsr" || n == "rbsm_noncons") p.SMSR = true;
  else if (n == "smsr_conservative" || n == "rbsm_cons") { p.SMSR = true; p.conservative = true; }
  else if (n == "rpcf" || n == "rpcf_noncons") p.RPCFPlusSMSR = true;
  else if (n == "rpcf_conservative" || n == "rpcf_cons") { p.RPCFPlusSMSR = true; p.conservative = true; }
  else if (n == "rsmss") p.RSMSS = true;
  else if (n == "rbssm" || n == "RBSSM") p.RBSSM = true;
  else if (n == "edtsm" || n == "edtsm_noncons") p.EDTSM = true;
  else if (n == "edtsm_conservative" || n == "edtsm_cons") { p.EDTSM = true; p.conservative = true; }
  else if (n == "tricubic" || n == "tricubicPCF" || n == "pcf_tricubic") { p.tricubicPCF = true; p.bilinearPCF = false; }   // shadowFilteringMenu case 1, main.cpp:672-675
  else if (n == "vsm" || n == "VSM") p.VSM = true;
  else if (n == "esm" || n == "ESM") p.ESM = true;
  else if (n == "evsm" || n == "EVSM") p.EVSM = true;
  else if (n == "msm" || n == "MSM") p.MSM = true;
  else if (n == "montecarlo" || n == "multi_hard") p.monteCarlo = true;
  else { g_err = "unknown technique " + n; return -2; }
  return 0;
}
int sgh_app_set_int(sgh_app* a, const char* name, int32_t v) {
  if (!a || !name) return -1;
  ShadowParams& p = a->app.shadowParams;
  std::string n(name);
  if (n == "kernelOrder") p.kernelOrder = v; else if (n == "penumbraSize") p.penumbraSize = v;
  else if (n == "maxSearch") p.maxSearch = v; else if (n == "blockerSearchSize") p.blockerSearchSize = v;
  else if (n == "kernelSize") p.kernelSize = v; else if (n == "lightSourceRadius") p.lightSourceRadius = v;
  else if (n == "numberOfSamples") p.numberOfSamples = v; else if (n == "lightSourceSize") p.lightSourceSize = v;
  else if (n == "svInfinity") a->app.svInfinity = v; else if (n == "svDepthFunc") a->app.svDepthFunc = v;
  else if (n == "svSilhouette") a->app.svSilhouette = v != 0; else if (n == "svZfail") a->app.svZfail = v != 0;
  else if (n == "animationOn") a->app.animationOn = v != 0;
  else if (n == "fusedMonteCarlo") a->app.fusedMonteCarlo = v != 0;
  else if (n == "commSkip") a->app.commSkip = v != 0;
  else if (n == "commMasks") a->app.commMasks = v != 0;
  else { g_err = "unknown int parameter " + n; return -2; }
  return 0;
}
int sgh_app_set_float(sgh_app* a, const char* name, float v) {
  if (!a || !name) return -1;
  std::string n(name);
  if (n == "shadowIntensity") a->app.shadowParams.shadowIntensity = v;
  else if (n == "depthThreshold") a->app.shadowParams.depthThreshold = v;
  else if (n == "animation") a->app.animation = v;
  else { g_err = "unknown float parameter " + n; return -2; }
  return 0;
}
int sgh_app_upload_scene(sgh_app* a) { return a ? a->app.uploadScene() : -1; }
// decoded pixels of the texture an `m` directive names (RGB8, row 0 = t 0): texture<index> of GBuffer.frag; NULL unbinds
int sgh_app_set_texture(sgh_app* a, int32_t index, const uint8_t* rgb, int32_t width, int32_t height) {
  if (!a) return -1;
  int rc = a->app.setTexture(index, rgb, width, height);
  if (rc) g_err = a->app.error();
  return rc;
}
int sgh_app_render_shadow_map(sgh_app* a) { return a ? a->app.renderShadowMap() : -1; }
int sgh_app_render_gbuffer(sgh_app* a) { return a ? a->app.renderGBuffer() : -1; }
int sgh_app_filter_shadow_map(sgh_app* a) { return a ? a->app.filterShadowMap() : -1; }
int sgh_app_compute_hard_shadows(sgh_app* a) { return a ? a->app.computeHardShadows() : -1; }
int sgh_app_render_soft_shadows(sgh_app* a) { return a ? a->app.renderSoftShadows() : -1; }
int sgh_app_render_monte_carlo(sgh_app* a) { return a ? a->app.renderMonteCarlo() : -1; }
int sgh_app_shade_scene(sgh_app* a) { return a ? a->app.shadeScene() : -1; }
// shadeScene() + the frame written as a PNG (the file-output counterpart of glutSwapBuffers / glReadPixels)
int sgh_app_save_image(sgh_app* a, const char* path) {
  if (!a || !path) return -1;
  int rc = a->app.shadeScene();
  if (rc) return rc;
  const int W = a->app.windowWidth, H = a->app.windowHeight;
  std::vector<float> img((size_t)W * H * 4);
  if ((rc = sgi_read(a->app.context(), SGI_BUF_SHADED, img.data(), img.size() * sizeof(float)))) { g_err = sgi_last_error(a->app.context()); return rc; }
  const std::vector<uint8_t> px = sgh::toRGBA8TopDown(img.data(), W, H);
  if (!sgh::writePNG(path, px.data(), W, H)) { g_err = std::string("cannot write ") + path; return -3; }
  return 0;
}
// the encoder alone (8-bit RGBA, row 0 = top): usable without a GPU
int sgh_write_png(const char* path, const uint8_t* rgba, int32_t W, int32_t H) { return (path && sgh::writePNG(path, rgba, W, H)) ? 0 : -1; }
int sgh_app_render_shadow_volumes(sgh_app* a) { return a ? a->app.renderShadowVolumes() : -1; }
int sgh_app_display(sgh_app* a, int32_t program) {
  if (!a) return -1;
  switch (program) {
    case SGH_PROGRAM_SHADOW_MAPPING: return a->app.display();
    case SGH_PROGRAM_SOFT_SHADOW_MAPPING: return a->app.displaySoft();
    case SGH_PROGRAM_SHADOW_VOLUMES: return a->app.displayShadowVolumes();
  }
  return -2;
}
// One frame the way the reference's display() spends it: geometry re-uploaded from host memory
// (loadVBOs on every draw), all passes, result read back to the caller's host buffer.
int sgh_app_display_e2e(sgh_app* a, int32_t program, int32_t result_buffer, void* host_dst, size_t bytes) {
  if (!a) return -1;
  int rc = 0;
  for (int attempt = 0; attempt < 4; attempt++) {     // a frame that outgrew the tile lists is run again (they have been re-sized)
    if ((rc = a->app.uploadScene())) return rc;
    if ((rc = sgh_app_display(a, program))) return rc;
    rc = sgi_read(a->app.context(), result_buffer, host_dst, bytes);
    if (rc != SGI_ERR_OVERFLOW) break;
  }
  if (rc) g_err = sgi_last_error(a->app.context());
  return rc;
}
// Pipelined form of the above: returns once the frame is queued; the result lands in host_dst when
// sgh_app_e2e_wait(ticket) returns.  The next frame can be issued in between (the GPU then copies frame k out while it
// renders frame k+1; the geometry upload is double-buffered on the device).
int sgh_app_display_e2e_async(sgh_app* a, int32_t program, int32_t result_buffer, void* host_dst, size_t bytes, int32_t* ticket) {
  if (!a || !ticket) return -1;
  int rc = a->app.uploadScene();
  if (rc) return rc;
  if ((rc = sgh_app_display(a, program))) return rc;
  rc = sgi_read_async(a->app.context(), result_buffer, host_dst, bytes, ticket);
  if (rc) g_err = sgi_last_error(a->app.context());
  return rc;
}
int sgh_app_e2e_wait(sgh_app* a, int32_t ticket) {
  if (!a) return -1;
  int rc = sgi_read_wait(a->app.context(), ticket);
  if (rc) g_err = sgi_last_error(a->app.context());
  return rc;
}
int sgh_app_step_animation(sgh_app* a, float delta) {   // idle(): animation += 6 (ShadowMapping/src/main.cpp:481)
  if (!a) return -1;
  a->app.animation += delta;
  return 0;
}

int sgh_procedural(const char* spec, float** xyz, int32_t* nv, int32_t** idx, int32_t* nt) {
  Mesh m;
  std::string e;
  if (!makeProcedural(spec, &m, &e)) { g_err = e; return -1; }
  *nv = m.getPointCloudSize() / 3; *nt = m.getNumberOfTriangles();
  *xyz = (float*)malloc(sizeof(float) * m.getPointCloudSize());
  *idx = (int32_t*)malloc(sizeof(int) * m.getIndicesSize());
  std::memcpy(*xyz, m.getPointCloud(), sizeof(float) * m.getPointCloudSize());
  std::memcpy(*idx, m.getIndices(), sizeof(int) * m.getIndicesSize());
  return 0;
}
void sgh_free(void* p) { free(p); }

}  // extern "C"
