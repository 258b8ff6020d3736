// shadow_app.h — the reference's frame orchestration (layer L4) over the C ABI instead of OpenGL.
//
// Mirrors, with the same names and the same argument meaning:
//   struct ShadowParams     ShadowMapping/include/Viewers/ShadowParams.h:6-37 (+ the SoftShadowMapping fields used
//                           here, SoftShadowMapping/include/Viewers/ShadowParams.h:6-69)
//   updateLight             ShadowMapping/src/main.cpp:209-219
//   renderShadowMap         :350-361  (displaySceneFromLightPOV :221-274)
//   renderGBuffer           :363-372  (displaySceneFromCameraPOV :276-300)
//   filterShadowMap         :374-398  (separable blur of the moment map; VSM / ESM / EVSM / MSM)
//   computeHardShadows      :400-414  (displaySceneFromGBuffer :302-348, configureShadow/configureRevectorization)
//   display                 :459-472
//   renderSoftShadows       SoftShadowMapping/src/main.cpp:925-1022 (PCSS branch)
//   renderMonteCarlo        SoftShadowMapping/src/main.cpp:756-811
//   SV display              ShadowVolumes/src/main.cpp:120-206 (update + stencil pass)
// The GL state, FBOs and shader programs of the original are replaced by one sgi_ctx; matrices are composed
// with glmath.h exactly as MyGLGeometryViewer does (configureAmbient :14-19, configurePhong :108-134,
// configureShadow :136-186).
#pragma once
#include <string>
#include <vector>

#include "../../include/shadowgi.h"
#include "glmath.h"
#include "mesh.h"
#include "scene_loader.h"

namespace sgh {

struct ShadowParams {
  Mat4 lightMVP, lightMV, lightP;
  int shadowMapWidth = 2048, shadowMapHeight = 2048;
  int maxSearch = 16;             // SMSR
  int kernelOrder = 7;
  int penumbraSize = 1;
  float depthThreshold = 0.0f;    // SMSR
  float shadowIntensity = 0.25f;
  bool tricubicPCF = false, bilinearPCF = true;
  bool VSM = false, ESM = false, EVSM = false, MSM = false;
  bool naive = false;
  bool SMSR = false, RPCFPlusSMSR = false, RSMSS = false, RPCFPlusRSMSS = false, EDTSM = false;
  bool useHardShadowMap = false, conservative = false;
  // SoftShadowMapping additions
  bool PCSS = true, monteCarlo = false, RBSSM = false;   // SoftShadowMapping/include/Viewers/ShadowParams.h:45
  int blockerSearchSize = 7, kernelSize = 15, lightSourceRadius = 8, numberOfSamples = 289;
  int lightSourceSize = 16;       // LightSource::size of the area light
  int windowWidth = 1024, windowHeight = 1024;
};

struct FrameMatrices {            // what the passes upload as uniforms for the current frame
  Mat4 lightMVP, lightMVPBiased, cameraMVP, cameraMV;
  Mat3 normalMatrix;
  Vec3 lightPositionShading;      // light eye rotated 180 deg about Y (main.cpp:283)
};

class ShadowApp {
 public:
  explicit ShadowApp(int device);
  ~ShadowApp();
  bool ok() const { return ctx != nullptr; }
  const std::string& error() const { return err; }

  // initGL(): scene + sizes (ShadowMapping/src/main.cpp:839-953)
  int loadScene(const char* config, const char* base_dir);
  int setScene(const float* xyz, const float* nrm, int nv, const int* idx, int nt, const float camEye[3], const float camAt[3],
               const float lightEyeCfg[3], const float lightAt_[3], float depthThreshold);
  void setWindowSize(int w, int h) { windowWidth = w; windowHeight = h; normalMatrixSet = false; }
  void setShadowMapSize(int w, int h) { shadowParams.shadowMapWidth = w; shadowParams.shadowMapHeight = h; }
  // scene textures (Mesh::loadTexture -> loadRGBTexture): the host side has no image decoder (the reference's is OpenCV); the caller
  // hands over decoded RGB8 pixels for texture<index> of GBuffer.frag (index = the `m` directive's running number - 1)
  int setTexture(int index, const unsigned char* rgb, int width, int height);
  int uploadScene();              // MyGLGeometryViewer::loadVBOs (:383-405); the reference calls it on every draw

  // per-frame passes
  int renderShadowMap();
  int renderGBuffer();
  int filterShadowMap();          // ShadowMapping/src/main.cpp:374-398 (VSM / ESM / EVSM / MSM)
  int computeHardShadows();
  int renderSoftShadows();
  int renderMonteCarlo();
  int renderShadowVolumes();
  int shadeScene();               // ShadowMapping/src/main.cpp:449-457 (deferred Phong into SGI_BUF_SHADED)
  int display();                  // ShadowMapping
  int displaySoft();              // SoftShadowMapping (PCSS or Monte-Carlo by shadowParams.monteCarlo)
  int displayShadowVolumes();     // ShadowVolumes

  void updateLight();
  FrameMatrices frameMatrices();  // for the current light/camera/animation state (single light)
  int technique() const;          // which sgi_technique the ShadowParams bools select (-1: out of scope)

  sgi_ctx* context() { return ctx; }
  Mesh* getScene() { return &scene; }

  ShadowParams shadowParams;
  Vec3 cameraEye{0, 0, 0}, cameraAt{0, 0, 0}, cameraUp{0, 0, 1};
  Vec3 lightEye{0, 0, 0}, lightAt{0, 0, 0}, lightUp{0, 0, 1};
  Vec3 lightPositionConfig{0, 0, 0};
  float translationVector[3] = {0, 0, 0}, lightTranslationVector[3] = {0, 0, 0}, rotationAngles[3] = {0, 0, 0};
  bool animationOn = false;
  float animation = -1800;        // main.cpp:143
  int windowWidth = 1024, windowHeight = 1024;
  int svInfinity = 100;           // ShadowVolumes/src/main.cpp:469
  int svDepthFunc = SGI_DEPTH_LEQUAL;
  bool svSilhouette = false;      // extrude silhouette / boundary edges only (interior side quads cancel in pairs); false = the reference's per-triangle prisms
  bool svZfail = false;           // depth-fail counting over capped volumes (robust when the eye is inside a volume); false = the reference's depth-pass stencil ops
  int rect[4] = {0, 0, 0, 0};     // multi-GPU screen tile (empty = whole window)
  int lightShardRank = 0, lightShardWorld = 1;   // multi-GPU many-light: this process owns lights l = rank (mod world)
  std::vector<int> lightOwner;    // many-light shards: rank that owns light s (empty: s mod world); set from measured per-light costs
  int measureLightCosts(float* ms, int n);   // depth-pass time of each of the n lights of renderMonteCarlo, one at a time (CUDA events)
  bool fusedMonteCarlo = false;   // renderMonteCarlo: camera pass reduced to primitive ids (sgi_render_prim_ids), positions resolved inside
                                  // the accumulation kernel (sgi_params.multi_fused); identical visibility, no vertex map materialised
  bool commMasks = false;         // light shards of at most 32 lights exchange lit masks (1 B/pixel per 8 lights) instead of float sums: the
                                  // un-sharded bits for EVERY shadow intensity, but measured slower at 8 GPUs (1.13 vs 1.04 ms)
  bool commSkip = false;          // measurement aid: run the sharded frame without its exchanges (what the collectives cost = the difference)
  bool commOn = false;            // sgi_comm_init done (commInit): renderMonteCarlo exchanges id strips / partial sums over NCCL itself
  int commInit(const void* id128, size_t bytes, int rank, int world);   // joins the NCCL communicator; light shard = (rank, world)

 private:
  int fail(int rc, const char* where);
  int pushParams(int technique);
  Mat4 modelMatrix() const;
  sgi_ctx* ctx = nullptr;
  Mesh scene;
  std::string err;
  Mat3 normalMatrix;              // frozen on the first camera-view pass (MyGLGeometryViewer.cpp:114-117)
  bool normalMatrixSet = false;
  bool uploaded = false, uvUploaded = false;
  std::vector<void*> pinned;        // scene arrays page-locked in place for DMA uploads (released before the arrays change)
  void pinSceneArrays(); void unpinSceneArrays();
  std::vector<float> uploadColors; const float* uploadColorsSrc = nullptr;   // colour array padded to the vertex count (uploadScene)
  int curTech = SGI_TECH_HARD;
};

}  // namespace sgh
