// obj_loader.h — Wavefront OBJ reader with the tokenisation rules of the reader the reference uses
// (Nate Robins' glm.c as vendored in ShadowMapping/src/IO/OBJLoader.cpp: glmFirstPass :441-571,
// glmSecondPass :581-764, glmReadOBJ :1309-1378).  What matters for parity with the reference's Mesh:
//   * the file is consumed as a whitespace-separated TOKEN stream (fscanf "%s"), not line by line: after
//     "v x y z" any extra tokens on the line (Meshlab's per-vertex r g b) are each treated as an unknown
//     keyword whose handler eats the rest of the line;
//   * faces accept v, v/t, v//n, v/t/n; polygons are fan-triangulated (v0, v_prev, v_new);
//   * negative indices are relative to the count so far (+1 because arrays are 1-based);
//   * vertices/normals/texcoords are stored 1-based (slot 0 unused), exactly like GLMmodel;
//   * a missing .mtl is tolerated (materials do not influence geometry).
// No OpenGL types, no drawing code, no global state; errors are returned, never exit().
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sgh {

struct ObjTriangle { uint32_t vindices[3]; uint32_t nindices[3]; uint32_t tindices[3]; };

struct ObjModel {
  uint32_t numvertices = 0, numnormals = 0, numtexcoords = 0, numtriangles = 0;
  std::vector<float> vertices;    // 3*(numvertices+1)
  std::vector<float> normals;     // 3*(numnormals+1)   (empty if none)
  std::vector<float> texcoords;   // 2*(numtexcoords+1) (empty if none)
  std::vector<ObjTriangle> triangles;
};

// Returns 0 on success; on failure a negative code and a message in err.
int readOBJ(const std::string& filename, ObjModel* model, std::string* err);

}  // namespace sgh
