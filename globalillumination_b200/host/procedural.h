// procedural.h — seeded procedural stand-ins for the assets that are absent from the reference mount
// (/root/reference/.MISSING_LARGE_BLOBS): OBJ/Sponza/sponza.obj (config c2, the headline
// scene), OBJ/SanDiego/sphere.obj (c5), OBJ/TreeWithLeaves/TreeSub1.obj (c4).  They are NOT the reference's
// meshes; results obtained on them say so (bench.py reports `scene: procedural`).  Deterministic: integer LCG,
// no libm calls whose rounding could differ between machines except sin/cos on exact table angles.
#pragma once
#include <string>

#include "mesh.h"

namespace sgh {

// spec = "<name>[?seed=N][&tris=N]"; names: sponza_like, sphere, leaves, plane
bool makeProcedural(const std::string& spec, Mesh* out, std::string* err);
// If `path` names one of the known-missing assets, fill `out` with its stand-in and return true.
bool substituteMissingAsset(const std::string& path, Mesh* out);

}  // namespace sgh
