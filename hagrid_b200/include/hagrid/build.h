// Construction API of the irregular grid — same five entry points, argument
// meaning and call order as the reference (src/build.h:17-31):
//   build_grid -> merge_grid -> flatten_grid -> expand_grid -> [compress_grid]
// All array pointers are device pointers; `grid.entries/cells/ref_ids/
// small_cells` are always allocated from the MemManager passed in, so the
// caller may mem.free() them before the next build (src/main.cpp:496-498).
#ifndef BUILD_H
#define BUILD_H

#include "mem_manager.h"
#include "hgb_types.h"

namespace hagrid {

/// Two-level build: uniform top grid of density `top_density`, per-top-cell
/// octree whose depth follows `snd_density`; voxel map = that octree.
void build_grid(MemManager& mem, const Tri* tris, int num_tris, Grid& grid, float top_density, float snd_density);

/// SAH-driven merging of face-aligned neighbour cells, x/y/z passes repeated
/// while the cell count shrinks below `alpha` x previous; alpha <= 0 disables.
void merge_grid(MemManager& mem, Grid& grid, float alpha);

/// Collapses uniform octree nodes and fuses up to 3 octree levels per voxel-map node.
void flatten_grid(MemManager& mem, Grid& grid);

/// Grows cell boxes over neighbours whose reference set is a subset, `iters` x (x,y,z).
void expand_grid(MemManager& mem, Grid& grid, const Tri* tris, int iters);

/// 16-byte cells + sentinel-terminated reference lists. False (grid untouched)
/// when a virtual dimension does not fit 16 bits.
bool compress_grid(MemManager& mem, Grid& grid);

} // namespace hagrid
#endif
