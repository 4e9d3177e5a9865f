// Compatibility name: the reference front end includes "bbox.h" (src/bbox.h);
// in this library every public type of the path lives in hgb_types.h.
#ifndef BBOX_H
#define BBOX_H
#include "hgb_types.h"
#endif
