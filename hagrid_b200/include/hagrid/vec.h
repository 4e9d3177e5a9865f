// Compatibility name: the reference front end includes "vec.h" (src/vec.h);
// in this library every public type of the path lives in hgb_types.h.
#ifndef VEC_H
#define VEC_H
#include "hgb_types.h"
#endif
