// scene_loader.h — the reference's Configs/*.txt reader (ShadowMapping/src/IO/SceneLoader.cpp:11-95, with the
// SoftShadowMapping `ha`/`hb` keys, SoftShadowMapping/src/IO/SceneLoader.cpp:93-96).  Same grammar, same
// order of application, same quirk (an empty line re-executes the previous directive with the stale value).
// Differences, all on the error path: a missing OBJ returns an error (the reference exit(1)s) unless a
// procedural stand-in is registered for that file name (procedural.h; the mount lacks several assets,
// SURVEY.md F10); textures are not decoded.
#pragma once
#include <string>
#include <vector>

#include "mesh.h"

namespace sgh {

class SceneLoader {
 public:
  SceneLoader(const char* filename, Mesh* mesh);
  // base_dir: directory the relative `o`/`m`/`cf` paths are resolved against (the reference uses the CWD)
  int load(const std::string& base_dir = "");
  float* getCameraPosition() { return cameraPosition; }
  float* getCameraAt() { return cameraAt; }
  float* getLightPosition() { return lightPosition; }
  float* getLightAt() { return lightAt; }
  float getDepthThreshold() const { return depthThreshold; }
  float getHSMAlpha() const { return HSMAlpha; }
  float getHSMBeta() const { return HSMBeta; }
  const std::string& error() const { return err; }
  const std::vector<std::string>& substitutions() const { return substituted; }

 private:
  std::string filename;
  Mesh* mesh;
  float cameraPosition[3] = {0, 0, 0}, cameraAt[3] = {0, 0, 0}, lightPosition[3] = {0, 0, 0}, lightAt[3] = {0, 0, 0};
  float depthThreshold = 0.0f, HSMAlpha = 0.0f, HSMBeta = 0.0f;
  std::string err;
  std::vector<std::string> substituted;
};

}  // namespace sgh
