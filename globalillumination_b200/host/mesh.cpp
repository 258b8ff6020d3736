// mesh.cpp — see mesh.h.  Line references are to ShadowMapping/src/Mesh.cpp in the reference.
#include "mesh.h"

#include <cstdio>
#include <cstring>

#include "glmath.h"
#include "obj_loader.h"

namespace sgh {

Mesh::Mesh(int numberOfPoints, int numberOfTriangles) {
  pointCloud.assign((size_t)numberOfPoints * 3, 0.0f);
  normalVector.assign((size_t)numberOfPoints * 3, 0.0f);
  colors.assign((size_t)numberOfPoints * 3, 0.0f);
  indices.assign((size_t)numberOfTriangles * 3, 0);
}

// :47-193 — concatenate, re-basing the appended indices by the previous vertex count
void Mesh::addObject(const Mesh* mesh) {
  int prevPoints = (int)pointCloud.size() / 3;
  pointCloud.insert(pointCloud.end(), mesh->pointCloud.begin(), mesh->pointCloud.end());
  // the reference copies pointCloudSize normals from the other mesh; pad if it had none
  size_t want = pointCloud.size();
  normalVector.insert(normalVector.end(), mesh->normalVector.begin(), mesh->normalVector.end());
  normalVector.resize(want, 0.0f);
  textureCoords.insert(textureCoords.end(), mesh->textureCoords.begin(), mesh->textureCoords.end());
  colors.insert(colors.end(), mesh->colors.begin(), mesh->colors.end());
  for (int v : mesh->indices) indices.push_back(v + prevPoints);
  numberOfTextures += mesh->numberOfTextures;
  if (numberOfTextures > 0) isTextureFromImage = true;
}

// :195-234 — running mean of face normals in triangle order; double arithmetic narrowed on store
void Mesh::computeNormals() {
  if (normalVector.size() != pointCloud.size()) normalVector.resize(pointCloud.size(), 0.0f);
  std::vector<int> nb_seen(pointCloud.size() / 3, 0);
  for (size_t i = 0; i + 2 < indices.size(); i += 3) {
    int a = indices[i], b = indices[i + 1], c = indices[i + 2];
    Vec3 pa{pointCloud[a * 3], pointCloud[a * 3 + 1], pointCloud[a * 3 + 2]};
    Vec3 pb{pointCloud[b * 3], pointCloud[b * 3 + 1], pointCloud[b * 3 + 2]};
    Vec3 pc{pointCloud[c * 3], pointCloud[c * 3 + 1], pointCloud[c * 3 + 2]};
    Vec3 normal = normalize(cross(Vec3{pb.x - pa.x, pb.y - pa.y, pb.z - pa.z}, Vec3{pc.x - pa.x, pc.y - pa.y, pc.z - pa.z}));
    int v[3] = {a, b, c};
    for (int j = 0; j < 3; j++) {
      int cur = v[j];
      nb_seen[cur]++;
      if (nb_seen[cur] == 1) {
        normalVector[cur * 3 + 0] = normal.x;
        normalVector[cur * 3 + 1] = normal.y;
        normalVector[cur * 3 + 2] = normal.z;
      } else {
        const float n[3] = {normal.x, normal.y, normal.z};
        for (int k = 0; k < 3; k++)
          normalVector[cur * 3 + k] = (float)(normalVector[cur * 3 + k] * (1.0 - 1.0 / nb_seen[cur]) + n[k] * 1.0 / nb_seen[cur]);
      }
    }
  }
}

void Mesh::computeCentroid(float* centroid) const {
  for (int a = 0; a < 3; a++) centroid[a] = 0;
  for (size_t p = 0; p < pointCloud.size() / 3; p++)
    for (int a = 0; a < 3; a++) centroid[a] += pointCloud[p * 3 + a];
  for (int a = 0; a < 3; a++) centroid[a] /= (int)(pointCloud.size() / 3);
}

// :251-313
int Mesh::loadOBJFile(const char* filename, std::string* err) {
  ObjModel model;
  int rc = readOBJ(filename, &model, err);
  if (rc) return rc;
  size_t nv = model.numvertices;
  pointCloud.assign(nv * 3, 0.0f);
  indices.assign((size_t)model.numtriangles * 3, 0);
  textureCoords.assign(nv * 3, 0.0f);            // last coordinate selects the texture
  for (size_t p = 0; p < nv; p++)
    for (int k = 0; k < 3; k++) pointCloud[p * 3 + k] = model.vertices[(p + 1) * 3 + k];
  normalVector.clear();
  if (model.numnormals > 0) {                    // :273-283 (always overwritten by computeNormals afterwards)
    normalVector.assign(nv * 3, 0.0f);
    for (size_t n = 0; n < nv && n < model.numnormals; n++)
      for (int k = 0; k < 3; k++) normalVector[n * 3 + k] = model.normals[(n + 1) * 3 + k];
  }
  for (size_t t = 0; t < model.numtriangles; t++)
    for (int k = 0; k < 3; k++) {
      int vi = (int)model.triangles[t].vindices[k] - 1;
      if (vi < 0 || (size_t)vi >= nv) { if (err) *err = std::string("loadOBJFile: face index out of range in ") + filename; return -3; }
      indices[t * 3 + k] = vi;
    }
  if (model.numtexcoords > 0) {
    for (size_t t = 0; t < model.numtriangles; t++)
      for (int k = 0; k < 3; k++) {
        uint32_t tc = model.triangles[t].tindices[k];
        uint32_t vc = model.triangles[t].vindices[k];
        if (tc > model.numtexcoords) continue;
        textureCoords[(vc - 1) * 3 + 0] = model.texcoords[tc * 2 + 0];
        textureCoords[(vc - 1) * 3 + 1] = model.texcoords[tc * 2 + 1];
        textureCoords[(vc - 1) * 3 + 2] = 0;
      }
  }
  return 0;
}

void Mesh::loadTexture(const char*, int ID) {      // :315-325 without the cv::imread
  numberOfTextures++;
  isTextureFromImage = true;
  for (size_t c = 0; c < textureCoords.size() / 3; c++) textureCoords[c * 3 + 2] = (float)ID;
}

// :327-363 — "v x y z r g b" lines
int Mesh::loadColorFromOBJFile(const char* filename, std::string* err) {
  FILE* file = fopen(filename, "r");
  if (!file) { if (err) *err = std::string("loadColorFromOBJFile: can't open \"") + filename + "\""; return -1; }
  colors.assign(pointCloud.size(), 0.0f);
  char buf[128];
  float temp[3], col[3];
  size_t nv = 0;
  while (fscanf(file, "%127s", buf) != EOF) {
    if (buf[0] == 'v' && buf[1] == '\0') {
      if (fscanf(file, "%f %f %f %f %f %f", &temp[0], &temp[1], &temp[2], &col[0], &col[1], &col[2]) == 6 && nv * 3 + 2 < colors.size()) {
        colors[nv * 3 + 0] = col[0]; colors[nv * 3 + 1] = col[1]; colors[nv * 3 + 2] = col[2];
      }
      nv++;
    }
  }
  fclose(file);
  return 0;
}

void Mesh::setBaseColor(float r, float g, float b) {
  colors.assign(pointCloud.size(), 0.0f);
  for (size_t c = 0; c < colors.size() / 3; c++) { colors[c * 3] = r; colors[c * 3 + 1] = g; colors[c * 3 + 2] = b; }
}

void Mesh::translate(float x, float y, float z) {
  for (size_t p = 0; p < pointCloud.size() / 3; p++) { pointCloud[p * 3] += x; pointCloud[p * 3 + 1] += y; pointCloud[p * 3 + 2] += z; }
}
void Mesh::scale(float x, float y, float z) {
  for (size_t p = 0; p < pointCloud.size() / 3; p++) { pointCloud[p * 3] *= x; pointCloud[p * 3 + 1] *= y; pointCloud[p * 3 + 2] *= z; }
}

// :402-430 — Rx*Ry*Rz transposed, applied as a row-vector product to positions AND normals
void Mesh::rotate(float x, float y, float z) {
  Mat4 R = sgh::rotate(x, Vec3{1, 0, 0});
  R = mul(R, sgh::rotate(y, Vec3{0, 1, 0}));
  R = mul(R, sgh::rotate(z, Vec3{0, 0, 1}));
  R = transpose(R);
  auto at = [&](int c, int r) { return R.m[c * 4 + r]; };
  auto apply = [&](float* v) {
    float rx = v[0] * at(0, 0) + v[1] * at(0, 1) + v[2] * at(0, 2);
    float ry = v[0] * at(1, 0) + v[1] * at(1, 1) + v[2] * at(1, 2);
    float rz = v[0] * at(2, 0) + v[1] * at(2, 1) + v[2] * at(2, 2);
    v[0] = rx; v[1] = ry; v[2] = rz;
  };
  if (normalVector.size() != pointCloud.size()) normalVector.resize(pointCloud.size(), 0.0f);
  for (size_t p = 0; p < pointCloud.size() / 3; p++) { apply(&pointCloud[p * 3]); apply(&normalVector[p * 3]); }
}

void Mesh::setGeometry(const float* xyz, int nv, const int* idx, int nt) {
  pointCloud.assign(xyz, xyz + (size_t)nv * 3);
  indices.assign(idx, idx + (size_t)nt * 3);
  textureCoords.assign((size_t)nv * 3, 0.0f);
  normalVector.clear();
}

}  // namespace sgh
