// hagrid_b200 — entry points of the irregular-grid path, C++ side.
//
// The reference's front end includes "build.h" and "traverse.h" (src/build.h:17-31, src/traverse.h:11-14);
// in this library both names forward to this header, which declares the same functions with the same
// signatures and call order
//   build_grid -> merge_grid -> flatten_grid -> expand_grid -> [compress_grid] -> setup_traversal -> traverse_grid*
// plus the entry points the reference does not have (primitive-id hits, host-buffer frames, camera frames).
// All array pointers are device pointers; `grid.entries/cells/ref_ids/small_cells` are always allocated from
// the MemManager passed in, so the caller may mem.free() them before the next build (src/main.cpp:496-498).
#ifndef HGB_API_H
#define HGB_API_H

#include "mem_manager.h"
#include "hgb_types.h"

namespace hagrid {

// ------------------------------------------------------------------ traversal
/// Captures the grid constants used by traverse_grid (once per grid).
void setup_traversal(const Grid& grid);

/// Closest hit of every ray; asynchronous on the legacy default stream (the first call for a ray buffer not seen
/// before waits about 10 us for a look at its layout: camera raster or incoherent).
/// Reference-verbatim result: Hit::id holds the traversal step count
/// (src/traverse.cu:80,93), Hit::t the hit distance (ray.tmax if none).
void traverse_grid(const Grid& grid, const Tri* tris, const Ray* rays, Hit* hits, int num_rays);

/// Same traversal, but Hit::id is the primitive index (-1 = no hit) as
/// documented in src/ray.h:22. Not part of the reference API.
void traverse_grid_prim_ids(const Grid& grid, const Tri* tris, const Ray* rays, Hit* hits, int num_rays);

/// One interactive frame with HOST buffers (the loop body of src/main.cpp:599-613): uploads the rays,
/// traces them, downloads the hits; returns when `host_hits` is complete. Upload, traversal and download
/// are pipelined in chunks over several streams (fastest with page-locked host buffers; pageable ones
/// work). `dev_rays`/`dev_hits` are device staging buffers of `num_rays` elements owned by the caller.
/// Not part of the reference API.
void traverse_grid_host(const Grid& grid, const Tri* tris, const Ray* host_rays, Hit* host_hits, int num_rays,
                        Ray* dev_rays, Hit* dev_hits, bool prim_ids);

/// One two-wave frame (BASELINE config C5) with everything resident, in one call: primary rays -> `hits_primary`
/// (primitive ids), their bounce rays (generate_bounce_rays with `keys`) -> `bounce`, those traced -> `hits_bounce`;
/// `counters` (two uint64, device, zeroed by the caller, may be null) receives count_hits of both hit buffers.
/// Ordered after earlier work on the legacy default stream and joined back into it. Same results as the five separate
/// calls. Not part of the reference API.
void trace_two_waves(const Grid& grid, const Tri* tris, int num_tris, const Ray* rays, int num_rays, const int* keys,
                     float offset, float tmax, unsigned seed, Hit* hits_primary, Ray* bounce, Hit* hits_bounce,
                     unsigned long long* counters);

/// Device-resident rays whose hits are wanted in host memory (the second wave of a frame: its rays were made on the
/// device by generate_bounce_rays): traced in chunks, every chunk's hits travel to the host while the next chunk is
/// traced. Returns when `host_hits` is complete. Not part of the reference API.
void traverse_grid_to_host(const Grid& grid, const Tri* tris, const Ray* dev_rays, Hit* dev_hits, Hit* host_hits, int num_rays,
                           bool prim_ids);

/// Camera of a frame: what gen_camera of the reference's front end produces (src/main.cpp:18-23,42-50).
struct FrameCamera { vec3 eye, right, up, dir; };

/// gen_camera (src/main.cpp:42-50), host arithmetic.
FrameCamera make_camera(const vec3& eye, const vec3& center, const vec3& up, float fov, float ratio);

/// gen_rays (src/main.cpp:52-66) on the device: width x height rays in scan-line order into the device
/// buffer `rays`, bit-identical to the host loop. Asynchronous on the legacy default stream.
void generate_rays(const FrameCamera& cam, float clip, int width, int height, Ray* rays);

/// Second wave (BASELINE config C5; the reference stops at primary rays): for every ray whose `hits[i].id`
/// names a triangle (hits from traverse_grid_prim_ids) emit a cosine-weighted diffuse bounce -- origin =
/// hit point + `offset` x unit normal facing the ray, direction from a counter-based generator keyed by
/// (`seed`, i), tmin 0, tmax `tmax`; rays that missed are emitted again unchanged. `out` may be `rays`.
/// IEEE arithmetic only: oracle/hagrid_oracle.c og_bounce_rays produces the same bits on the CPU.
/// `keys` (device, one int per ray, may be null = the ray's index in this buffer) names each ray's random stream: a
/// shard of a frame passes the rays' indices in the whole frame and gets the rays the unsharded frame would get.
/// Asynchronous on the legacy default stream. Not part of the reference API.
void generate_bounce_rays(const Tri* tris, int num_tris, const Ray* rays, const Hit* hits, int num_rays,
                          float offset, float tmax, unsigned seed, Ray* out, const int* keys = nullptr);

/// Per-frame counters of a hit buffer (primitive-id hits): counters[0] += hits with id >= 0, counters[1] += sum of
/// (id + 1); `counters` is device memory the caller zeroes. What the ranks of a sharded frame all-reduce.
/// Asynchronous on the legacy default stream. Not part of the reference API.
void count_hits(const Hit* hits, int num_hits, unsigned long long* counters);

/// One frame of the reference's viewer (src/main.cpp:591-625) fused into one launch: primary rays are
/// generated, traced and coloured on the device; `pixels` (device, width * height BGRA words) receives
/// what update_surface (src/main.cpp:90-111) would write: mode 0 = depth, 1 = step count as grey,
/// 2 = step count as heat map. Asynchronous on the legacy default stream. Not part of the reference API.
void render_frame(const Grid& grid, const Tri* tris, const FrameCamera& cam, float clip, int width, int height,
                  int mode, unsigned* pixels);

/// Tuning switches: "traverse_variant" (0 = one thread per ray, 1 = persistent phase-scheduled warps,
/// 2 = one thread per ray re-tiled 8x4 on rasters, 4 = resident warps pulling 8x4 tiles, 3 = automatic) and
/// "host_frame_chunk_rays" (chunk size of traverse_grid_host). Returns false for unknown keys.
bool set_traversal_option(const char* key, int value);

// ------------------------------------------------------------------ construction
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

// ------------------------------------------------------------------ device-wide primitives
// What the reference's `Parallel` wrapper offers on top of CUB (src/parallel.cuh:12-89), on this library's own kernels;
// all pointers are device pointers, scratch comes from `mem`, work runs on the legacy default stream.
/// out[i] = sum of in[0, i) for i in [0, n]; `out` has n + 1 elements and may be `in`. elem_bytes 4: int32; 8: two packed
/// 32-bit counters per element (every prefix of each half below 2^32).
void prim_exclusive_scan(MemManager& mem, const void* in, int n, int elem_bytes, void* out);
/// *out = reduction of in[0, n): op 0 = sum of int32, 1 = max of int32, 2 = min of float, 3 = max of float (4 bytes).
void prim_reduce(MemManager& mem, const void* in, int n, int op, void* out);
/// cub::DevicePartition::Flagged's order: items with a non-zero flag first, in input order, the others behind them in
/// REVERSE input order (the order the reference's build relies on, src/build.cu:568-569). Returns the number kept.
int prim_partition(MemManager& mem, const int* in, const int* flags, int n, int* out);
/// Stable sort of (key, value) pairs by the low `bits` bits of the non-negative keys, in place.
void prim_sort_pairs(MemManager& mem, int* keys, int* vals, int n, int bits);

} // namespace hagrid
#endif
