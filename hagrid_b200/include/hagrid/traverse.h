// Compatibility name: the reference's front end includes "traverse.h" (src/traverse.h); the traversal entry
// points of this library are declared in hgb_api.h.
#ifndef TRAVERSE_H
#define TRAVERSE_H
#include "hgb_api.h"
#endif
