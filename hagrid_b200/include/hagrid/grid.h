// Compatibility name: the reference front end includes "grid.h" (src/grid.h);
// in this library every public type of the path lives in hgb_types.h.
#ifndef GRID_H
#define GRID_H
#include "hgb_types.h"
#include "hgb_inline.h"     // lookup_entry, foreach_ref, intersect_prim_ray, load_ray, ... (the reference's inline helpers)
#endif
