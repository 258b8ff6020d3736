// ref_host_glue.cpp — C API over the reference's OWN host classes (compiled from /root/reference by
// build_ref.py into oracle/_ref/libref_host.so).  TEST INFRASTRUCTURE ONLY: used to pin oracle/ and the
// product's host code (scene loader, matrices, prisms, light samples) against the reference's arithmetic.
//
// Matrix sequences follow ShadowMapping/src/main.cpp:221-348 and
// ShadowMapping/src/Viewers/MyGLGeometryViewer.cpp:14-19,108-134,136-186 using the vendored GLM 0.9.3.1.
#include <unistd.h>
#include <cstdlib>
#include <cstring>
#include "Mesh.h"
#include "IO/SceneLoader.h"
#include "ShadowVolume.h"
#include "glm/gtc/matrix_inverse.hpp"
#include "glm/gtx/transform.hpp"
#ifdef SSM_INCLUDE
#include "Scene/LightSource/UniformSampledLightSource.h"
#endif

// Image.cpp needs OpenCV (absent); textures only feed colour, so a no-op stands in.
Image::Image(int width, int height, int channels) {
  data = (unsigned char*)malloc((size_t)width * height * channels + 1);
  this->width = width;
  this->height = height;
}
Image::Image(char*) { data = (unsigned char*)malloc(1); width = 0; height = 0; }
Image::~Image() { free(data); }

struct RefScene { Mesh* mesh; SceneLoader* loader; };

static void put(const glm::mat4& m, float* o) { memcpy(o, &m[0][0], 64); }

extern "C" {

void* ref_scene_load(const char* root, const char* config) {
  if (chdir(root) != 0) return nullptr;
  if (access(config, R_OK) != 0) return nullptr;
  RefScene* s = new RefScene;
  s->mesh = new Mesh();
  s->loader = new SceneLoader((char*)config, s->mesh);
  s->loader->load();
  return s;
}
void ref_scene_counts(void* h, int* nv, int* nt) {
  RefScene* s = (RefScene*)h;
  *nv = s->mesh->getPointCloudSize() / 3;
  *nt = s->mesh->getNumberOfTriangles();
}
void ref_scene_copy(void* h, float* xyz, float* nrm, int* idx) {
  RefScene* s = (RefScene*)h;
  memcpy(xyz, s->mesh->getPointCloud(), sizeof(float) * s->mesh->getPointCloudSize());
  memcpy(nrm, s->mesh->getNormalVector(), sizeof(float) * s->mesh->getPointCloudSize());
  memcpy(idx, s->mesh->getIndices(), sizeof(int) * s->mesh->getIndicesSize());
}
void ref_scene_views(void* h, float* cam_eye, float* cam_at, float* light_eye, float* light_at, float* depth_threshold) {
  RefScene* s = (RefScene*)h;
  for (int a = 0; a < 3; a++) {
    cam_eye[a] = s->loader->getCameraPosition()[a];
    cam_at[a] = s->loader->getCameraAt()[a];
    light_eye[a] = s->loader->getLightPosition()[a];
    light_at[a] = s->loader->getLightAt()[a];
  }
  *depth_threshold = s->loader->getDepthThreshold();
}

void ref_perspective(float fov, float aspect, float zn, float zf, float* out) { put(glm::perspective(fov, aspect, zn, zf), out); }
void ref_look_at(const float* e, const float* a, const float* u, float* out) {
  put(glm::lookAt(glm::vec3(e[0], e[1], e[2]), glm::vec3(a[0], a[1], a[2]), glm::vec3(u[0], u[1], u[2])), out);
}
void ref_rotate(float angle, const float* ax, float* out) { put(glm::rotate(angle, glm::vec3(ax[0], ax[1], ax[2])), out); }

// One frame's uniforms exactly as display() derives them (no user translation/rotation, no animation).
void ref_frame_matrices(const float* cam_eye, const float* cam_at, const float* light_eye_in, const float* light_at, int W,
                        int H, int SW, int SH, float* cam_mvp, float* cam_mv, float* normal_matrix9, float* light_mvp,
                        float* light_mvp_biased, float* light_pos_shading) {
  const float fov = 45.f, zNear = 1.0f, zFar = 1000.0f;             // MyGLGeometryViewer.cpp:6-8
  glm::vec3 up(0, 0, 1);                                            // main.cpp:866-867
  glm::vec3 lightEye(light_eye_in[0], light_eye_in[1], light_eye_in[2]);
  // displaySceneFromLightPOV (main.cpp:249-266)
  glm::mat4 projection = glm::perspective(fov, (float)SW / SH, zNear, zFar);
  glm::mat4 view = glm::lookAt(lightEye, glm::vec3(light_at[0], light_at[1], light_at[2]), up);
  glm::mat4 model = glm::mat4(1.0f);
  model *= glm::translate(glm::vec3(0.0f, 0.0f, 0.0f));
  model *= glm::rotate(0.0f, glm::vec3(1, 0, 0));
  model *= glm::rotate(0.0f, glm::vec3(0, 1, 0));
  model *= glm::rotate(0.0f, glm::vec3(0, 0, 1));
  glm::mat4 lightMVP = projection * view * model;
  put(lightMVP, light_mvp);
  // displaySceneFromCameraPOV / displaySceneFromGBuffer (main.cpp:282-283,290-297,336-344)
  glm::vec3 shadingLight = glm::mat3(glm::rotate((float)180.0, glm::vec3(0, 1, 0))) * lightEye;
  light_pos_shading[0] = shadingLight[0]; light_pos_shading[1] = shadingLight[1]; light_pos_shading[2] = shadingLight[2];
  projection = glm::perspective(fov, (float)W / H, zNear, zFar);
  view = glm::lookAt(glm::vec3(cam_eye[0], cam_eye[1], cam_eye[2]), glm::vec3(cam_at[0], cam_at[1], cam_at[2]), up);
  model = glm::mat4(1.0f);
  model *= glm::translate(glm::vec3(0.0f, 0.0f, 0.0f));
  model *= glm::rotate(0.0f, glm::vec3(1, 0, 0));
  model *= glm::rotate(0.0f, glm::vec3(0, 1, 0));
  model *= glm::rotate(0.0f, glm::vec3(0, 0, 1));
  glm::mat4 mvp = projection * view * model;                         // configurePhong :111-112
  glm::mat4 mv = view * model;
  glm::mat3 normalMatrix = glm::inverseTranspose(glm::mat3(mv));     // :115
  put(mvp, cam_mvp);
  put(mv, cam_mv);
  memcpy(normal_matrix9, &normalMatrix[0][0], 36);
  glm::mat4 bias;                                                    // configureShadow :139-145
  bias[0][0] = 0.5; bias[0][1] = 0;   bias[0][2] = 0;   bias[0][3] = 0.0;
  bias[1][0] = 0;   bias[1][1] = 0.5; bias[1][2] = 0;   bias[1][3] = 0.0;
  bias[2][0] = 0;   bias[2][1] = 0;   bias[2][2] = 0.5; bias[2][3] = 0.0;
  bias[3][0] = 0.5; bias[3][1] = 0.5; bias[3][2] = 0.5; bias[3][3] = 1.0;
  put(bias * lightMVP, light_mvp_biased);
}

// ShadowVolume::build on caller-provided arrays (ShadowVolumes/src/ShadowVolume.cpp:15-114)
void ref_sv_prisms(const float* xyz, const float* nrm, int V, const int* idx, int T, const float* light, int infinity,
                   float* prism_xyz, int* prism_idx) {
  Mesh* scene = new Mesh(V, T);
  memcpy(scene->getPointCloud(), xyz, sizeof(float) * 3 * V);
  memcpy(scene->getNormalVector(), nrm, sizeof(float) * 3 * V);
  memcpy(scene->getIndices(), idx, sizeof(int) * 3 * T);
  ShadowVolume* sv = new ShadowVolume(infinity);
  glm::vec3 L(light[0], light[1], light[2]);
  sv->build(scene, L);
  sv->update(scene, L);
  memcpy(prism_xyz, sv->getData()->getPointCloud(), sizeof(float) * 18 * T);
  memcpy(prism_idx, sv->getData()->getIndices(), sizeof(int) * 18 * T);
  // leaked on purpose: the reference's destructors mix malloc/delete[]
}

// MyGLGeometryViewer::configureMoments (ShadowMapping/src/Viewers/MyGLGeometryViewer.cpp:191-199): the 16 numbers are typed
// into mQuantization[i][j] (passed in here in that order), then glm::transpose and glm::inverse of the vendored GLM run.
void ref_moment_quantization(const float typed16[16], float m_out[16], float minv_out[16]) {
  glm::mat4 mQuantization, mQuantizationInverse;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) mQuantization[i][j] = typed16[i * 4 + j];
  mQuantization = glm::transpose(mQuantization);
  mQuantizationInverse = glm::inverse(mQuantization);
  put(mQuantization, m_out);
  put(mQuantizationInverse, minv_out);
}

#ifdef SSM_INCLUDE
void ref_uniform_sample(const float* p, int size, int n_lights, int index, float* eye_out) {
  LightSource base;
  base.setEye(glm::vec3(p[0], p[1], p[2]));
  base.setAt(glm::vec3(p[0], p[1], p[2]));
  base.setUp(glm::vec3(0, 0, 1));
  base.setSize(size);
  UniformSampledLightSource u(&base, n_lights);
  glm::vec3 e = u.getEye(index);
  eye_out[0] = e[0]; eye_out[1] = e[1]; eye_out[2] = e[2];
}
#endif
}
