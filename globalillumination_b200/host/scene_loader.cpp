// scene_loader.cpp — see scene_loader.h.  Line references: ShadowMapping/src/IO/SceneLoader.cpp.
#include "scene_loader.h"

#include <cstdlib>
#include <fstream>
#include <memory>
#include <sstream>

#include "procedural.h"

namespace sgh {

SceneLoader::SceneLoader(const char* fn, Mesh* m) : filename(fn), mesh(m) {}

static std::string join(const std::string& base, const std::string& p) {
  if (base.empty() || (!p.empty() && p[0] == '/')) return p;
  return base + "/" + p;
}

int SceneLoader::load(const std::string& base_dir) {
  std::ifstream file(filename);
  if (!file) { err = "SceneLoader: can't open \"" + filename + "\""; return -1; }
  std::string line, key, value;                       // persist across lines, as in the reference (:14)
  std::unique_ptr<Mesh> temp;
  float scale[3], translate[3], rotate[3], color[3];
  int numberOfTextures = 0;
  auto need_temp = [&]() { if (!temp) { err = "SceneLoader: directive before any `o` line in " + filename; return false; } return true; };
  while (!file.eof()) {                               // :22
    std::getline(file, line);
    std::istringstream split(line);
    split >> key;                                     // an empty line leaves `key` unchanged (:27)
    if (key.empty()) continue;
    if (key[0] == 'o') {                              // :29-33
      split >> value;
      temp.reset(new Mesh());
      std::string e;
      int rc = temp->loadOBJFile(join(base_dir, value).c_str(), &e);
      if (rc == -1 && value.rfind("procedural:", 0) == 0) rc = makeProcedural(value.substr(11), temp.get(), &e) ? 0 : -4;
      else if (rc == -1 && substituteMissingAsset(value, temp.get())) { substituted.push_back(value); rc = 0; }
      if (rc) { err = e; return rc; }
      temp->computeNormals();
    } else if (key[0] == 'm') {                       // :34-37
      split >> value;
      if (!need_temp()) return -5;
      numberOfTextures++;
      temp->loadTexture(value.c_str(), numberOfTextures);
    } else if (key[0] == 's') {                       // :38-43
      for (int a = 0; a < 3; a++) { split >> value; scale[a] = (float)atof(value.c_str()); }
      if (!need_temp()) return -5;
      temp->scale(scale[0], scale[1], scale[2]);
    } else if (key[0] == 't') {                       // :44-49
      for (int a = 0; a < 3; a++) { split >> value; translate[a] = (float)atof(value.c_str()); }
      if (!need_temp()) return -5;
      temp->translate(translate[0], translate[1], translate[2]);
    } else if (key[0] == 'r') {                       // :50-55
      for (int a = 0; a < 3; a++) { split >> value; rotate[a] = (float)atof(value.c_str()); }
      if (!need_temp()) return -5;
      temp->rotate(rotate[0], rotate[1], rotate[2]);
    } else if (key[0] == '+') {                       // :56-58
      if (!need_temp()) return -5;
      mesh->addObject(temp.get());
      temp.reset();
    } else if (key[0] == 'v') {                       // :59-64
      for (int a = 0; a < 3; a++) {
        split >> value;
        if (key.size() > 1 && key[1] == 'e') cameraPosition[a] = (float)atof(value.c_str());
        else cameraAt[a] = (float)atof(value.c_str());
      }
    } else if (key[0] == 'l') {                       // :65-70
      for (int a = 0; a < 3; a++) {
        split >> value;
        if (key.size() > 1 && key[1] == 'e') lightPosition[a] = (float)atof(value.c_str());
        else lightAt[a] = (float)atof(value.c_str());
      }
    } else if (key[0] == 'c') {                       // :71-86
      if (!need_temp()) return -5;
      if (key.size() > 1 && key[1] == 'f') {
        split >> value;
        std::string e;
        if (temp->loadColorFromOBJFile(join(base_dir, value).c_str(), &e) != 0) temp->setBaseColor(1.0f, 1.0f, 1.0f);
      } else {
        for (int a = 0; a < 3; a++) { split >> value; color[a] = (float)atof(value.c_str()); }
        temp->setBaseColor(color[0], color[1], color[2]);
      }
    } else if (key[0] == 'd') {                       // :88-90
      split >> value;
      depthThreshold = (float)atof(value.c_str());
    } else if (key[0] == 'h') {                       // SoftShadowMapping/src/IO/SceneLoader.cpp:93-96
      split >> value;
      if (key.size() > 1 && key[1] == 'a') HSMAlpha = (float)atof(value.c_str());
      else HSMBeta = (float)atof(value.c_str());
    }
  }
  return 0;
}

}  // namespace sgh
