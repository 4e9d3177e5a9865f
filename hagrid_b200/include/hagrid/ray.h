// Compatibility name: the reference front end includes "ray.h" (src/ray.h);
// in this library every public type of the path lives in hgb_types.h.
#ifndef RAY_H
#define RAY_H
#include "hgb_types.h"
#endif
