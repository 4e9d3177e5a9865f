// Compatibility name: the reference's front end includes "build.h" (src/build.h); the construction entry
// points of this library are declared in hgb_api.h.
#ifndef BUILD_H
#define BUILD_H
#include "hgb_api.h"
#endif
