// Traversal API — src/traverse.h:11-14 plus one extra entry point.
#ifndef TRAVERSE_H
#define TRAVERSE_H

#include "hgb_types.h"

namespace hagrid {

/// Captures the grid constants used by traverse_grid (once per grid).
void setup_traversal(const Grid& grid);

/// Closest hit of every ray; asynchronous on the legacy default stream.
/// Reference-verbatim result: Hit::id holds the traversal step count
/// (src/traverse.cu:80,93), Hit::t the hit distance (ray.tmax if none).
void traverse_grid(const Grid& grid, const Tri* tris, const Ray* rays, Hit* hits, int num_rays);

/// Same traversal, but Hit::id is the primitive index (-1 = no hit) as
/// documented in src/ray.h:22. Not part of the reference API.
void traverse_grid_prim_ids(const Grid& grid, const Tri* tris, const Ray* rays, Hit* hits, int num_rays);

/// Tuning switches ("traverse_variant": 0 = one thread per ray, 1 = persistent
/// phase-scheduled warps). Returns false for unknown keys.
bool set_traversal_option(const char* key, int value);

} // namespace hagrid
#endif
