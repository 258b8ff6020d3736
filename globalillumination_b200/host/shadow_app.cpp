// shadow_app.cpp — see shadow_app.h.
#include "shadow_app.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace sgh {

ShadowApp::ShadowApp(int device) {
  int rc = sgi_create(&ctx, device);
  if (rc != SGI_OK) { ctx = nullptr; err = "sgi_create failed (no CUDA device; this path has no CPU fallback)"; }
  normalMatrix = Mat3{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
}
ShadowApp::~ShadowApp() {
  if (ctx) sgi_synchronize(ctx);
  unpinSceneArrays();
  if (ctx) sgi_destroy(ctx);
}

// The Mesh arrays are re-uploaded on every frame of the end-to-end loop (the reference's loadVBOs per draw): page-lock
// them in place once per scene so that the uploads are DMA transfers without a staging copy on the host.
void ShadowApp::pinSceneArrays() {
  if (!ctx || !pinned.empty()) return;
  struct { void* p; size_t n; } arr[] = {
      {scene.getPointCloud(), sizeof(float) * (size_t)scene.getPointCloudSize()}, {scene.getNormalVector(), sizeof(float) * (size_t)scene.getPointCloudSize()},
      {scene.getIndices(), sizeof(int) * 3 * (size_t)scene.getNumberOfTriangles()}, {uploadColors.empty() ? nullptr : uploadColors.data(), sizeof(float) * uploadColors.size()}};
  for (auto& a : arr)
    if (a.p && a.n && sgi_register_host(a.p, a.n) == SGI_OK) pinned.push_back(a.p);
  sgi_set_option(ctx, "borrow_pinned", 1);
}
void ShadowApp::unpinSceneArrays() {
  if (pinned.empty()) return;
  if (ctx) sgi_synchronize(ctx);                    // no DMA may still be reading them
  for (void* p : pinned) sgi_unregister_host(p);
  pinned.clear();
}

int ShadowApp::fail(int rc, const char* where) {
  err = std::string(where) + ": " + (ctx ? sgi_last_error(ctx) : "no context");
  return rc;
}

int ShadowApp::loadScene(const char* config, const char* base_dir) {
  unpinSceneArrays(); uploadColors.clear(); uploadColorsSrc = nullptr;
  scene = Mesh();
  SceneLoader loader(config, &scene);
  int rc = loader.load(base_dir ? base_dir : "");
  if (rc) { err = loader.error(); return rc; }
  cameraEye = Vec3{loader.getCameraPosition()[0], loader.getCameraPosition()[1], loader.getCameraPosition()[2]};   // main.cpp:861-865
  cameraAt = Vec3{loader.getCameraAt()[0], loader.getCameraAt()[1], loader.getCameraAt()[2]};
  lightPositionConfig = Vec3{loader.getLightPosition()[0], loader.getLightPosition()[1], loader.getLightPosition()[2]};
  lightAt = Vec3{loader.getLightAt()[0], loader.getLightAt()[1], loader.getLightAt()[2]};
  shadowParams.depthThreshold = loader.getDepthThreshold();
  uploaded = false; uvUploaded = false; normalMatrixSet = false;
  return 0;
}

int ShadowApp::setScene(const float* xyz, const float* nrm, int nv, const int* idx, int nt, const float camEye[3], const float camAt_[3],
                        const float lightEyeCfg[3], const float lightAt_[3], float depthThreshold) {
  if (nv < 0 || nt < 0 || (nv > 0 && !xyz) || (nt > 0 && !idx)) { err = "setScene: bad arguments"; return SGI_ERR_INVALID; }
  for (long long k = 0; k < 3LL * nt; k++)                 // Mesh::computeNormals indexes the vertex arrays with these
    if (idx[k] < 0 || idx[k] >= nv) { err = "setScene: vertex index out of range"; return SGI_ERR_INVALID; }
  unpinSceneArrays(); uploadColors.clear(); uploadColorsSrc = nullptr;
  scene = Mesh();
  scene.setGeometry(xyz, nv, idx, nt);
  scene.computeNormals();
  if (nrm) std::memcpy(scene.getNormalVector(), nrm, sizeof(float) * 3 * (size_t)nv);
  cameraEye = Vec3{camEye[0], camEye[1], camEye[2]}; cameraAt = Vec3{camAt_[0], camAt_[1], camAt_[2]};
  lightPositionConfig = Vec3{lightEyeCfg[0], lightEyeCfg[1], lightEyeCfg[2]}; lightAt = Vec3{lightAt_[0], lightAt_[1], lightAt_[2]};
  shadowParams.depthThreshold = depthThreshold;
  uploaded = false; uvUploaded = false; normalMatrixSet = false;
  return 0;
}

int ShadowApp::uploadScene() {
  if (!ctx) return SGI_ERR_NO_DEVICE;
  if (scene.getColorsSize() > 0 && (uploadColors.size() != (size_t)scene.getPointCloudSize() || uploadColorsSrc != scene.getColors())) {
    // per-vertex colours (`c` / `cf` directives): objects without a colour directive leave the reference's colour array
    // short (Mesh::addObject); those vertices are shaded white here.  Built once per scene.
    unpinSceneArrays();
    uploadColors.assign((size_t)scene.getPointCloudSize(), 1.0f);
    std::memcpy(uploadColors.data(), scene.getColors(), sizeof(float) * (size_t)std::min(scene.getColorsSize(), scene.getPointCloudSize()));
    uploadColorsSrc = scene.getColors();
  }
  pinSceneArrays();
  int rc = sgi_set_mesh(ctx, scene.getPointCloud(), scene.getNormalVector(), scene.getPointCloudSize() / 3, scene.getIndices(),
                        scene.getNumberOfTriangles());
  if (rc) return fail(rc, "uploadScene");
  if (scene.getColorsSize() > 0) {                 // feeds the albedo target and shadeScene()
    if ((rc = sgi_set_mesh_colors(ctx, uploadColors.data()))) return fail(rc, "sgi_set_mesh_colors");
  } else if ((rc = sgi_set_mesh_colors(ctx, nullptr))) return fail(rc, "sgi_set_mesh_colors");
  // texture coordinates (u, v, texture id) feed the texture select of GBuffer.frag once a texture is bound (setTexture)
  // (once per scene: they do not change from frame to frame, and the per-frame re-upload of the end-to-end loop stays asynchronous)
  if (!uvUploaded) {
    if (scene.getTextureCoordsSize() == scene.getPointCloudSize() && scene.getTextureCoordsSize() > 0) {
      if ((rc = sgi_set_mesh_uv(ctx, scene.getTextureCoords()))) return fail(rc, "sgi_set_mesh_uv");
    } else if ((rc = sgi_set_mesh_uv(ctx, nullptr))) return fail(rc, "sgi_set_mesh_uv");
    uvUploaded = true;
  }
  uploaded = true;
  return 0;
}

int ShadowApp::setTexture(int index, const unsigned char* rgb, int width, int height) {
  if (!ctx) return SGI_ERR_NO_DEVICE;
  int rc = sgi_set_texture(ctx, index, rgb, width, height);
  return rc ? fail(rc, "sgi_set_texture") : 0;
}

// main.cpp:209-219
void ShadowApp::updateLight() {
  lightEye.x = lightPositionConfig.x + lightTranslationVector[0];
  lightEye.y = lightPositionConfig.y + lightTranslationVector[1];
  lightEye.z = lightPositionConfig.z + lightTranslationVector[2];
  if (animationOn) lightEye = mul3(rotate((float)animation / 10, Vec3{0, 1, 0}), lightEye);
}

// main.cpp:190-195
Mat4 ShadowApp::modelMatrix() const {
  Mat4 model = identity();
  model = mul(model, translate(Vec3{translationVector[0], translationVector[1], translationVector[2]}));
  model = mul(model, rotate(rotationAngles[0], Vec3{1, 0, 0}));
  model = mul(model, rotate(rotationAngles[1], Vec3{0, 1, 0}));
  model = mul(model, rotate(rotationAngles[2], Vec3{0, 0, 1}));
  return model;
}

static FrameMatrices composeFrame(Vec3 lightEye, Vec3 lightAt, Vec3 lightUp, Vec3 camEye, Vec3 camAt, Vec3 camUp, const Mat4& model,
                                  int W, int H, int SW, int SH) {
  const float fov = 45.f, zNear = 1.0f, zFar = 1000.0f;                 // MyGLGeometryViewer.cpp:6-8
  FrameMatrices f;
  Mat4 projection = perspective(fov, (float)SW / SH, zNear, zFar);      // configureAmbient :17
  Mat4 view = lookAt(lightEye, lightAt, lightUp);
  f.lightMVP = mul(mul(projection, view), model);                       // main.cpp:264
  f.lightMVPBiased = mul(biasMatrix(), f.lightMVP);                     // configureShadow :145
  f.lightPositionShading = mul3(rotate(180.0f, Vec3{0, 1, 0}), lightEye);   // main.cpp:283
  projection = perspective(fov, (float)W / H, zNear, zFar);
  view = lookAt(camEye, camAt, camUp);
  f.cameraMVP = mul(mul(projection, view), model);                      // configurePhong :111-112
  f.cameraMV = mul(view, model);
  f.normalMatrix = inverseTranspose3(f.cameraMV);                       // :115
  return f;
}

FrameMatrices ShadowApp::frameMatrices() {
  updateLight();
  FrameMatrices f = composeFrame(lightEye, lightAt, lightUp, cameraEye, cameraAt, cameraUp, modelMatrix(), windowWidth, windowHeight,
                                 shadowParams.shadowMapWidth, shadowParams.shadowMapHeight);
  if (!normalMatrixSet) { normalMatrix = f.normalMatrix; normalMatrixSet = true; }   // frozen (:114-117)
  f.normalMatrix = normalMatrix;
  return f;
}

// computeHardShadows' program selection (main.cpp:406-411) + the branch each program takes on its uniforms
int ShadowApp::technique() const {
  const ShadowParams& p = shadowParams;
  if (p.EDTSM) return p.conservative ? SGI_TECH_EDTSM_CONS : SGI_TECH_EDTSM_NONCONS;   // + filterHardShadowsUsingEDT (main.cpp:466)
  if (p.SMSR || p.RPCFPlusSMSR) {
    bool smsr = p.SMSR;                                                 // NonConservativeSMSR.frag:381
    if (p.conservative) return smsr ? SGI_TECH_RBSM_CONS : SGI_TECH_RPCF_CONS;
    return smsr ? SGI_TECH_RBSM_NONCONS : SGI_TECH_RPCF_NONCONS;
  }
  if (p.RSMSS || p.RPCFPlusRSMSS) return SGI_TECH_RSMSS;
  if (p.naive) return SGI_TECH_HARD;                                    // Shadow.frag:253
  if (p.VSM) return SGI_TECH_VSM;                                       // Shadow.frag:257-264, in the shader's order
  if (p.ESM) return SGI_TECH_ESM;
  if (p.EVSM) return SGI_TECH_EVSM;
  if (p.MSM) return SGI_TECH_MSM;
  if (p.tricubicPCF && !p.bilinearPCF) return SGI_TECH_PCF_TRICUBIC;    // Shadow.frag:101-104: the bilinear assignment comes second and wins
  return SGI_TECH_PCF;
}

int ShadowApp::pushParams(int tech) {
  sgi_params q;
  sgi_default_params(&q);
  const ShadowParams& p = shadowParams;
  q.technique = tech;
  curTech = tech;
  q.shadow_map_width = p.shadowMapWidth; q.shadow_map_height = p.shadowMapHeight;
  q.shadow_intensity = p.shadowIntensity;
  q.kernel_order = p.kernelOrder; q.penumbra_size = p.penumbraSize;
  q.blocker_search_size = p.blockerSearchSize; q.kernel_size = p.kernelSize; q.light_source_radius = p.lightSourceRadius;
  q.max_search = p.maxSearch; q.depth_threshold = p.depthThreshold;
  q.sv_depth_func = svDepthFunc; q.sv_infinity = svInfinity;
  q.sv_silhouette = svSilhouette ? 1 : 0; q.sv_zfail = svZfail ? 1 : 0;
  q.rect_x0 = rect[0]; q.rect_y0 = rect[1]; q.rect_x1 = rect[2]; q.rect_y1 = rect[3];
  // light shards exchange lit masks (1 bit per light and pixel) when the set has at most 32 lights and the fused pass is on,
  // un-normalised float sums otherwise
  const bool masks = commOn && commMasks && fusedMonteCarlo && shadowParams.numberOfSamples <= 32;
  q.multi_partial = (tech == SGI_TECH_MULTI_HARD && lightShardWorld > 1) ? (masks ? 2 : 1) : 0;
  q.multi_fused = (tech == SGI_TECH_MULTI_HARD && fusedMonteCarlo) ? 1 : 0;
  int rc = sgi_set_params(ctx, &q);
  return rc ? fail(rc, "sgi_set_params") : 0;
}

int ShadowApp::renderShadowMap() {
  if (!ctx) return SGI_ERR_NO_DEVICE;
  if (!uploaded) { int rc = uploadScene(); if (rc) return rc; }
  FrameMatrices f = frameMatrices();
  shadowParams.lightMVP = f.lightMVP;
  int rc = sgi_set_lights(ctx, 1, f.lightMVP.m, f.lightMVPBiased.m, &f.lightPositionShading.x, shadowParams.shadowMapWidth,
                          shadowParams.shadowMapHeight);
  if (rc) return fail(rc, "sgi_set_lights");
  if ((rc = pushParams(technique() < 0 ? SGI_TECH_HARD : technique()))) return rc;
  rc = sgi_render_shadow_map(ctx);
  return rc ? fail(rc, "sgi_render_shadow_map") : 0;
}

int ShadowApp::renderGBuffer() {
  if (!ctx) return SGI_ERR_NO_DEVICE;
  if (!uploaded) { int rc = uploadScene(); if (rc) return rc; }
  FrameMatrices f = frameMatrices();
  int rc = pushParams(curTech);                              // the screen rectangle travels in the params block
  if (rc) return rc;
  rc = sgi_set_camera(ctx, f.cameraMVP.m, f.cameraMV.m, f.normalMatrix.m, windowWidth, windowHeight);
  if (rc) return fail(rc, "sgi_set_camera");
  rc = sgi_render_gbuffer(ctx);
  return rc ? fail(rc, "sgi_render_gbuffer") : 0;
}

int ShadowApp::computeHardShadows() {
  if (!ctx) return SGI_ERR_NO_DEVICE;
  int tech = technique();
  if (tech < 0) { err = "computeHardShadows: no technique selected"; return SGI_ERR_INVALID; }
  int rc = pushParams(tech);
  if (rc) return rc;
  rc = sgi_compute_visibility(ctx);
  return rc ? fail(rc, "sgi_compute_visibility") : 0;
}

int ShadowApp::shadeScene() {
  if (!ctx) return SGI_ERR_NO_DEVICE;
  const float clear[4] = {0.63f, 0.82f, 0.96f, 1.0f};        // glClearColor in shadeScene (main.cpp:453)
  int rc = sgi_shade_phong(ctx, clear);
  return rc ? fail(rc, "sgi_shade_phong") : 0;
}

// filterShadowMap(), main.cpp:374-398: the blurred maps are window-sized, so the window goes down with the camera uniforms first
int ShadowApp::filterShadowMap() {
  if (!ctx) return SGI_ERR_NO_DEVICE;
  FrameMatrices f = frameMatrices();
  int rc = sgi_set_camera(ctx, f.cameraMVP.m, f.cameraMV.m, f.normalMatrix.m, windowWidth, windowHeight);
  if (rc) return fail(rc, "sgi_set_camera");
  rc = sgi_filter_shadow_map(ctx);
  return rc ? fail(rc, "sgi_filter_shadow_map") : 0;
}

int ShadowApp::display() {                                  // main.cpp:459-472 without shadeScene/swap
  int rc;
  if ((rc = renderShadowMap())) return rc;
  if (shadowParams.VSM || shadowParams.ESM || shadowParams.EVSM || shadowParams.MSM) { if ((rc = filterShadowMap())) return rc; }   // :463
  if ((rc = renderGBuffer())) return rc;
  return computeHardShadows();
}

int ShadowApp::renderSoftShadows() {                        // SoftShadowMapping/src/main.cpp:925-1022: G-buffer first, then the map
  int rc;
  if ((rc = renderGBuffer())) return rc;
  if ((rc = renderShadowMap())) return rc;
  if (!shadowParams.PCSS && !shadowParams.RBSSM) { err = "renderSoftShadows: only the PCSS and RBSSM branches are built (SURVEY.md C17/C18)"; return SGI_ERR_INVALID; }
  if ((rc = pushParams(shadowParams.RBSSM ? SGI_TECH_RBSSM : SGI_TECH_PCSS))) return rc;      // main.cpp:1012-1013
  rc = sgi_compute_visibility(ctx);
  return rc ? fail(rc, "sgi_compute_visibility") : 0;
}

// UniformSampledLightSource::computeUniformSampling (SoftShadowMapping/src/Scene/LightSource/UniformSampledLightSource.cpp:27-38)
static Vec3 uniformSample(Vec3 sample, int size, int numberOfPointLights, int sampleIndex) {
  float halfSize = (float)((float)size / 2.0);
  float factor = sqrtf((float)numberOfPointLights);
  float sampleSize = (float)((factor - 1) / 2.0);
  sample.x += (((sampleIndex % (int)factor) - sampleSize) / sampleSize) * halfSize;
  sample.y += (((int)(sampleIndex / factor) - sampleSize) / sampleSize) * halfSize;
  return sample;
}

int ShadowApp::renderMonteCarlo() {                          // SoftShadowMapping/src/main.cpp:756-811
  if (!ctx) return SGI_ERR_NO_DEVICE;
  int rc;
  // the reference re-uses the previous technique's G-buffer; here it is rendered explicitly
  if (!fusedMonteCarlo) { if ((rc = renderGBuffer())) return rc; }
  else {
    // AccurateSoftShadow.frag reads only the vertex map: the camera pass stores the winning primitive per pixel and the
    // accumulation kernel interpolates the position itself.  On a light shard every rank rasterises its own screen strip and
    // the id strips are all-gathered (4 B/pixel over NVLink) while the rank's depth passes run.
    if (!uploaded) { if ((rc = uploadScene())) return rc; }
    FrameMatrices fc = frameMatrices();
    if ((rc = sgi_set_camera(ctx, fc.cameraMVP.m, fc.cameraMV.m, fc.normalMatrix.m, windowWidth, windowHeight))) return fail(rc, "sgi_set_camera");
    int keep[4] = {rect[0], rect[1], rect[2], rect[3]};
    if (commOn && lightShardWorld > 1) {
      int32_t r0 = 0, r1 = 0;
      if ((rc = sgi_comm_strip(ctx, lightShardRank, &r0, &r1))) return fail(rc, "sgi_comm_strip");
      rect[0] = 0; rect[1] = r0; rect[2] = windowWidth; rect[3] = r1;
    }
    rc = pushParams(SGI_TECH_MULTI_HARD);
    for (int k = 0; k < 4; k++) rect[k] = keep[k];
    if (rc) return rc;
    if ((rc = sgi_render_prim_ids(ctx))) return fail(rc, "sgi_render_prim_ids");
    if (commOn && lightShardWorld > 1 && !commSkip) { if ((rc = sgi_gather(ctx, SGI_BUF_PRIM_ID))) return fail(rc, "sgi_gather"); }
  }
  updateLight();
  int n = shadowParams.numberOfSamples;
  if (n <= 0 || n > 1024) { err = "renderMonteCarlo: numberOfSamples must be in 1..1024"; return SGI_ERR_INVALID; }
  // light shard: this process builds and samples only lights s = rank (mod world); the shader's common
  // term still comes from the LAST light of the whole set, as in the reference (main.cpp:790,806)
  std::vector<float> mvp, mvpb;
  std::vector<int32_t> mine_ids;
  Mat4 model = modelMatrix();
  FrameMatrices f;
  for (int s = 0; s < n; s++) {
    bool mine = ((int)lightOwner.size() == n ? lightOwner[s] : s % lightShardWorld) == lightShardRank;
    if (!mine && s != n - 1) continue;
    Vec3 e = uniformSample(lightEye, shadowParams.lightSourceSize, n, s), a = uniformSample(lightAt, shadowParams.lightSourceSize, n, s);
    f = composeFrame(e, a, lightUp, cameraEye, cameraAt, cameraUp, model, windowWidth, windowHeight, shadowParams.shadowMapWidth,
                     shadowParams.shadowMapHeight);
    if (mine) {
      mvp.insert(mvp.end(), f.lightMVP.m, f.lightMVP.m + 16);
      mvpb.insert(mvpb.end(), f.lightMVPBiased.m, f.lightMVPBiased.m + 16);
      mine_ids.push_back(s);
    }
  }
  if (mvp.empty()) { err = "renderMonteCarlo: this rank owns no light (more ranks than lights)"; return SGI_ERR_INVALID; }
  Vec3 shading = mul3(rotate(180.0f, Vec3{0, 1, 0}), lightEye);
  rc = sgi_set_lights(ctx, (int)(mvp.size() / 16), mvp.data(), mvpb.data(), &shading.x, shadowParams.shadowMapWidth, shadowParams.shadowMapHeight);
  if (rc) return fail(rc, "sgi_set_lights");
  if (n <= 32) { if ((rc = sgi_set_light_ids(ctx, (int32_t)mine_ids.size(), mine_ids.data(), n))) return fail(rc, "sgi_set_light_ids"); }
  rc = sgi_set_multi_light_common(ctx, lightShardWorld > 1 ? f.lightMVPBiased.m : nullptr);   // f = light n-1 here
  if (rc) return fail(rc, "sgi_set_multi_light_common");
  if ((rc = pushParams(SGI_TECH_MULTI_HARD))) return rc;
  if ((rc = sgi_render_shadow_map(ctx))) return fail(rc, "sgi_render_shadow_map");
  if ((rc = sgi_compute_visibility(ctx))) return fail(rc, "sgi_compute_visibility");
  // partial sums of the ranks -> final visibility of this rank's strip (reduce-scatter + the division of AccurateSoftShadow.frag:127)
  if (commOn && lightShardWorld > 1 && !commSkip) { if ((rc = sgi_reduce_lights(ctx, n))) return fail(rc, "sgi_reduce_lights"); }
  return 0;
}

// Cost of each light's depth pass (they differ with what the light sees): the caller balances the light shards with them.
int ShadowApp::measureLightCosts(float* ms, int n) {
  if (!ctx) return SGI_ERR_NO_DEVICE;
  if (n <= 0 || n > 1024) { err = "measureLightCosts: n must be in 1..1024"; return SGI_ERR_INVALID; }
  int rc;
  if (!uploaded) { if ((rc = uploadScene())) return rc; }
  updateLight();
  Mat4 model = modelMatrix();
  Vec3 shading = mul3(rotate(180.0f, Vec3{0, 1, 0}), lightEye);
  if ((rc = pushParams(SGI_TECH_HARD))) return rc;
  for (int pass = 0; pass < 2; pass++)                 // first round: allocations and list sizing; second round: timed
    for (int s = 0; s < n; s++) {
      Vec3 e = uniformSample(lightEye, shadowParams.lightSourceSize, n, s), a = uniformSample(lightAt, shadowParams.lightSourceSize, n, s);
      FrameMatrices f = composeFrame(e, a, lightUp, cameraEye, cameraAt, cameraUp, model, windowWidth, windowHeight, shadowParams.shadowMapWidth,
                                     shadowParams.shadowMapHeight);
      if ((rc = sgi_set_lights(ctx, 1, f.lightMVP.m, f.lightMVPBiased.m, &shading.x, shadowParams.shadowMapWidth, shadowParams.shadowMapHeight))) return fail(rc, "sgi_set_lights");
      if (pass) { sgi_enable_timing(ctx, 1); sgi_reset_timing(ctx); }
      if ((rc = sgi_render_shadow_map(ctx))) return fail(rc, "sgi_render_shadow_map");
      rc = sgi_synchronize(ctx);
      if (rc && rc != SGI_ERR_OVERFLOW) return fail(rc, "sgi_synchronize");
      if (pass) {
        double t = 0; int64_t calls = 0;
        sgi_pass_time_ms(ctx, SGI_PASS_SHADOW_MAP, &t, &calls);
        ms[s] = (float)t;
        sgi_enable_timing(ctx, 0);
      }
    }
  return 0;
}

int ShadowApp::commInit(const void* id128, size_t bytes, int rank, int world) {
  if (!ctx) return SGI_ERR_NO_DEVICE;
  int rc = sgi_comm_init(ctx, id128, bytes, rank, world);
  if (rc) return fail(rc, "sgi_comm_init");
  lightShardRank = rank; lightShardWorld = world; commOn = true;
  return 0;
}

int ShadowApp::displaySoft() { return shadowParams.monteCarlo ? renderMonteCarlo() : renderSoftShadows(); }

int ShadowApp::renderShadowVolumes() {                       // ShadowVolumes/src/main.cpp:126-172
  if (!ctx) return SGI_ERR_NO_DEVICE;
  int rc;
  if ((rc = renderGBuffer())) return rc;                     // depth pre-pass (:154-158)
  if ((rc = pushParams(SGI_TECH_HARD))) return rc;
  updateLight();
  rc = sgi_compute_shadow_volume(ctx, &lightEye.x);          // shadowVolume->update(scene, lightEye) + stencil pass
  return rc ? fail(rc, "sgi_compute_shadow_volume") : 0;
}
int ShadowApp::displayShadowVolumes() { return renderShadowVolumes(); }

}  // namespace sgh
