// Scene ingest: Wavefront OBJ -> device Tri array, with the semantics of the
// reference's loader (src/load_obj.cpp:78-239) and triangle setup (load_model, src/main.cpp:246-275).
#pragma once

#include <string>
#include <vector>

#include "hgb_types.h"

namespace hagrid {

/// Geometry of an OBJ file as the reference's front end consumes it: positions (index 0 is the dummy
/// vertex of src/load_obj.cpp:96) and one index triple per fan triangle, in file order.
struct ObjGeometry {
    std::vector<vec3> vertices;
    std::vector<int>  indices;      // 3 per triangle
    std::string error;              // non-empty: the reference's loader would have refused the file
};

/// Parses `path` with `threads` worker threads (0 = hardware concurrency). Returns false when the file cannot
/// be read or contains what the reference counts as an error (unknown command, invalid face or index).
bool parse_obj(const std::string& path, int threads, ObjGeometry& out);

/// e1 = v0 - v1, e2 = v2 - v0, n = e1 x e2 for every index triple, on the device, in the host's arithmetic
/// (IEEE, no contraction): `tris` (device, indices.size() / 3 records) equals what load_model builds on the CPU.
/// `dev_vertices` / `dev_indices` are device copies of the arrays above. Asynchronous on the legacy stream.
void setup_triangles(const vec3* dev_vertices, const int* dev_indices, int num_tris, Tri* tris);

} // namespace hagrid
