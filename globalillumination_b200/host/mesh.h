// mesh.h — the reference's Mesh container, re-typed as portable C++17
// (ShadowMapping/include/Mesh.h:16-66, ShadowMapping/src/Mesh.cpp).  Same public interface and the same
// arithmetic (so vertex/normal arrays are bit-identical to the reference loader's), std::vector storage
// instead of malloc/delete[], no OpenCV: loadTexture only records the texture ID in uv.z (the images feed
// colour, not visibility — SURVEY.md C4).  The arrays returned by the getters are the C-ABI geometry input
// (sgi_set_mesh).
#pragma once
#include <string>
#include <vector>

namespace sgh {

class Mesh {
 public:
  Mesh() = default;
  Mesh(int numberOfPoints, int numberOfTriangles);

  void addObject(const Mesh* mesh);                          // Mesh.cpp:47-193
  void computeNormals();                                     // :195-234
  void computeCentroid(float* centroid) const;               // :236-249
  int loadOBJFile(const char* filename, std::string* err);   // :251-313 (returns <0 instead of exit(1))
  void loadTexture(const char* filename, int ID);            // :315-325 (ID bookkeeping only)
  int loadColorFromOBJFile(const char* filename, std::string* err);  // :327-363
  void translate(float x, float y, float z);                 // :378-388
  void scale(float x, float y, float z);                     // :390-400
  void rotate(float x, float y, float z);                    // :402-430
  void setBaseColor(float r, float g, float b);              // :365-376
  void setGeometry(const float* xyz, int nv, const int* idx, int nt);   // procedural stand-ins for missing assets

  float* getPointCloud() { return pointCloud.data(); }
  float* getNormalVector() { return normalVector.data(); }
  float* getTextureCoords() { return textureCoords.data(); }
  float* getColors() { return colors.data(); }
  int* getIndices() { return indices.data(); }
  const float* getPointCloud() const { return pointCloud.data(); }
  const float* getNormalVector() const { return normalVector.data(); }
  const int* getIndices() const { return indices.data(); }

  int getPointCloudSize() const { return (int)pointCloud.size(); }
  int getIndicesSize() const { return (int)indices.size(); }
  int getTextureCoordsSize() const { return (int)textureCoords.size(); }
  int getColorsSize() const { return (int)colors.size(); }
  int getNumberOfTextures() const { return numberOfTextures; }
  int getNumberOfTriangles() const { return (int)indices.size() / 3; }
  bool textureFromImage() const { return isTextureFromImage; }

 private:
  std::vector<float> pointCloud, normalVector, textureCoords, colors;
  std::vector<int> indices;
  int numberOfTextures = 0;
  bool isTextureFromImage = false;
};

}  // namespace sgh
