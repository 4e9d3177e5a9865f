/* hagrid_b200 — C ABI of the B200-native irregular-grid build + traversal path.
 *
 * The reference (cg-saarland/hagrid) exposes this path as C++ symbols in
 * `namespace hagrid` (src/build.h:17-31, src/traverse.h:11-14) driven by
 * src/main.cpp.  There is no FFI in the reference; this header is the plain-C
 * boundary a foreign host (ctypes, cgo, JNI ...) binds instead.  Every entry
 * point cites the reference interface it stands for.  All pointers are raw
 * host or device addresses, all sizes are element counts or bytes; no C++ or
 * torch types cross this boundary.
 *
 * The same source that implements this ABI (hagrid_b200/csrc/c_api.cpp) also
 * compiles, unchanged, against the reference's own headers and objects
 * (oracle/build_ref.sh -> oracle/_ref/libhagrid_ref.so); that is how the
 * parity tests drive the reference and this library through one interface.
 *
 * Error behaviour: the reference aborts the process on any CUDA error
 * (src/common.h:101-108).  The C++ API of this library keeps that behaviour;
 * the C ABI functions return 0 on success and a negative code on argument
 * errors (the message is available from hgb_last_error()).
 */
#ifndef HAGRID_B200_H
#define HAGRID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGB_MAX_LEVELS 32

#if defined(__GNUC__)
#define HGB_API __attribute__((visibility("default")))
#else
#define HGB_API
#endif

/* Opaque scene handle: one MemManager (src/mem_manager.h:34), one device Tri
 * array (src/prims.h:13) and one Grid (src/grid.h:48), as main.cpp:471-478
 * sets them up. */
typedef struct hgb_scene hgb_scene;

/* Host-visible part of `hagrid::Grid` (src/grid.h:48-62). */
typedef struct hgb_grid_info {
    float   bbox_min[3];
    float   bbox_max[3];
    int32_t dims[3];        /* top-level dimensions                         */
    int32_t shift;          /* log2(virtual dims / top-level dims)          */
    int32_t num_cells;
    int32_t num_entries;
    int32_t num_refs;
    int32_t compressed;     /* 1 when small_cells != nullptr                */
    int32_t num_offsets;
    int32_t offsets[HGB_MAX_LEVELS];
} hgb_grid_info;

/* What traverse_grid stores in Hit::id. The reference kernel overwrites the
 * primitive id with its step counter (src/traverse.cu:93); HGB_HIT_STEPS is
 * that verbatim behaviour, HGB_HIT_PRIM_ID keeps the id documented in
 * src/ray.h:22 (-1 = no hit). */
enum { HGB_HIT_STEPS = 0, HGB_HIT_PRIM_ID = 1 };

/* Device arrays that can be downloaded / uploaded. */
enum {
    HGB_ARRAY_ENTRIES     = 0,  /* Entry[num_entries]       4 B each  (grid.h:12)  */
    HGB_ARRAY_CELLS       = 1,  /* Cell[num_cells]         32 B each  (grid.h:23)  */
    HGB_ARRAY_SMALL_CELLS = 2,  /* SmallCell[num_cells]    16 B each  (grid.h:36)  */
    HGB_ARRAY_REFS        = 3,  /* int[num_refs]                                     */
    HGB_ARRAY_TRIS        = 4   /* Tri[num_tris]           48 B each  (prims.h:13) */
};

/* Identification. `hgb_impl()` returns "hagrid_b200" for this library and
 * "reference" for the reference build of the same ABI. */
HGB_API const char* hgb_impl(void);
HGB_API const char* hgb_last_error(void);
HGB_API int  hgb_device_count(void);
/* Tuning/diagnostic switches of this library (no reference counterpart; the
 * reference build of this ABI accepts and ignores them). Keys:
 *   "traverse_variant"       0 = one thread per ray, 1 = persistent phase-scheduled warps, 2 = one thread per
 *                            ray re-tiled 8x4 when the buffer is a raster, 4 = resident warps pulling 8x4 tiles,
 *                            3 = automatic (default: 4 for rasters, 1 otherwise)
 *   "host_frame_chunk_rays"  rays per full-size chunk of hgb_traverse_grid_host's pipeline
 *   "tile_min_rays"          rasters below this many rays go to variant 2 instead of 4 under "automatic" (default
 *                            128 K); "tile_cold_min_rays" is the same line for launches without a ticket list
 *                            ("tile_order" 0, or the second launch on a buffer; default 1280 K)
 *   "vote_min_rays"          incoherent buffers below this many rays go to variant 0 instead of 1
 *   "tile_order"             variant 4 hands the tiles of a ray buffer out by what they cost on earlier launches
 *                            of the same buffer, longest first, in 2^value cost classes (default 8; 0 = buffer order)
 *   "tile_split"             ... and up to this many of the most expensive tiles in parts (default 256; 0 = none)
 *   "tile_split_log"         ... 2^value parts per tile (default 2; at most 5 = single rays)
 *   "tile_split_share"       ... only tiles that cost at least value % of one resident warp's share of the launch
 *                            (default 50)
 *   "two_wave_chunks"        pieces hgb_trace_two_waves cuts a frame into (default 1)
 *   "ray_sort"               1 = bin incoherent rays by octant and entry cell before tracing (default 0)
 *   "merge_one_launch_max_cells"  hgb_merge_grid runs all its passes in one cooperative launch on grids of up to this
 *                            many cells before merging (default 768 K; 0 = one launch per kernel and pass)
 * None of them can change a hit. Returns 0 when the key is known. */
HGB_API int  hgb_set_option(const char* key, int value);
/* Diagnosis: the per-tile times variant 4 last recorded for the device ray buffer (`dev_rays`, `num_rays`), in SM
 * clock ticks / 64 per tile of 32 rays (8x4 pixels of a raster), copied to `host_costs` (at most `capacity`
 * entries). Returns the number of entries written, 0 when nothing is recorded for that buffer (always 0 in the
 * reference build). */
HGB_API int  hgb_tile_costs(const void* dev_rays, int num_rays, unsigned short* host_costs, int capacity);

/* Scene life cycle: MemManager(keep) + Tri upload, main.cpp:471-478. */
HGB_API hgb_scene* hgb_scene_create(int device, int keep_alive);
HGB_API void       hgb_scene_destroy(hgb_scene* scene);
/* `host_tris`: num_tris x 48 B {v0,nx,e1,ny,e2,nz} (prims.h:13-16). */
HGB_API int  hgb_scene_set_tris(hgb_scene* scene, const void* host_tris, int num_tris);
/* Scene ingest: load_model of the front end (src/main.cpp:246-275) = ObjLoader::load_obj (src/load_obj.cpp:78-239)
 * + fan triangulation + triangle setup. This library parses the file with `threads` host threads (0 = all) and
 * builds the 48-byte records on the device; the reference build of this ABI runs the reference's own
 * single-threaded load_model and uploads the result. Returns the number of triangles, negative on error
 * (unreadable file, or anything the reference's loader counts as an error). Two inputs the reference mishandles are
 * errors here: a position index beyond the file's vertices (the reference reads out of bounds) and a line of 1023
 * characters or more, comments included (the reference's getline fails there and it silently keeps the geometry
 * parsed so far, src/load_obj.cpp:103-105). */
HGB_API int  hgb_scene_load_obj(hgb_scene* scene, const char* path, int threads);
/* The host half of the ingest on its own (no device needed): positions (index 0 = the loader's dummy vertex,
 * src/load_obj.cpp:96) and three position indices per fan triangle, in file order. */
typedef struct hgb_obj hgb_obj;
HGB_API hgb_obj*     hgb_obj_parse(const char* path, int threads);       /* NULL on error, see hgb_last_error() */
HGB_API int          hgb_obj_num_vertices(const hgb_obj* obj);
HGB_API int          hgb_obj_num_tris(const hgb_obj* obj);
HGB_API const float* hgb_obj_vertices(const hgb_obj* obj);               /* 3 floats per vertex */
HGB_API const int*   hgb_obj_indices(const hgb_obj* obj);                /* 3 ints per triangle */
HGB_API void         hgb_obj_free(hgb_obj* obj);
HGB_API int  hgb_scene_num_tris(const hgb_scene* scene);
/* Peak bytes handed out by the scene's MemManager (mem_manager.h:104). */
HGB_API size_t hgb_scene_peak_bytes(const hgb_scene* scene);

/* The five construction stages (build.h:17,20,25,28,31).  hgb_build_grid
 * first releases the previous grid arrays like main.cpp:496-498 does. */
HGB_API int  hgb_build_grid(hgb_scene* scene, float top_density, float snd_density);
HGB_API int  hgb_merge_grid(hgb_scene* scene, float alpha);
HGB_API int  hgb_flatten_grid(hgb_scene* scene);
HGB_API int  hgb_expand_grid(hgb_scene* scene, int iters);
/* Returns 1 when compressed, 0 when refused (virtual dims >= 65536,
 * compress.cu:41-44), negative on error. */
HGB_API int  hgb_compress_grid(hgb_scene* scene);

/* main.cpp:494-508: one event-timed (profile(), profile.cu:5-18) pass of
 * build+merge+flatten+expand[+compress], `iters` times after `warmup`
 * untimed passes; ms_out[iters] receives each pass' milliseconds. */
HGB_API int  hgb_build_pipeline(hgb_scene* scene, float top_density, float snd_density,
                        float alpha, int exp_iters, int compress,
                        int warmup, int iters, float* ms_out);

/* traverse.h:11 and :14.  Rays: 32 B {org,tmin,dir,tmax} (ray.h:9-20), hits:
 * 16 B {id,t,u,v} (ray.h:23-33); `dev_*` are device pointers. The launch is
 * asynchronous on the legacy default stream like the reference's. */
HGB_API int  hgb_setup_traversal(hgb_scene* scene);
HGB_API int  hgb_traverse_grid(hgb_scene* scene, const void* dev_rays, void* dev_hits,
                       int num_rays, int hit_mode);
/* main.cpp:414-425: `warmup` untimed launches then `iters` launches each
 * timed with profile(); ms_out[iters]. */
HGB_API int  hgb_traverse_timed(hgb_scene* scene, const void* dev_rays, void* dev_hits,
                        int num_rays, int hit_mode, int warmup, int iters,
                        float* ms_out);
/* Interactive-frame shape (main.cpp:599-613): H2D rays, traverse, D2H hits,
 * everything inside the call; host buffers may be pageable or pinned. */
HGB_API int  hgb_traverse_grid_host(hgb_scene* scene, const void* host_rays,
                            void* host_hits, int num_rays, int hit_mode);

/* Camera frames: the interactive loop of the reference's front end (src/main.cpp:591-625).
 * `cam` is 12 floats in the member order of its Camera struct (src/main.cpp:18-23): eye, right, up, dir.
 *   hgb_make_camera    gen_camera  (src/main.cpp:42-50), host arithmetic
 *   hgb_generate_rays  gen_rays    (src/main.cpp:52-66): width*height rays, scan-line order, into a DEVICE buffer
 *   hgb_render_frame   gen_rays + traverse_grid + update_surface (src/main.cpp:90-111, 598-621) for one frame;
 *                      `host_bgra` receives width*height BGRA words; display_mode 0 = depth, 1 = steps as grey,
 *                      2 = steps as heat map (DisplayMode, src/main.cpp:34-38). This library does all three
 *                      steps in one launch on the device; the reference build of this ABI runs the reference's own
 *                      host loop (CPU ray generation, upload, trace, download, CPU colouring). */
HGB_API int  hgb_make_camera(const float eye[3], const float center[3], const float up[3], float fov, float ratio,
                             float cam_out[12]);
HGB_API int  hgb_generate_rays(hgb_scene* scene, const float cam[12], float clip, int width, int height, void* dev_rays);
HGB_API int  hgb_render_frame(hgb_scene* scene, const float cam[12], float clip, int width, int height,
                              int display_mode, void* host_bgra);

/* Second wave of BASELINE.json's config C5. The reference's front end stops at primary rays
 * (gen_rays, src/main.cpp:52-66); this entry point defines the next stage on the device so that the hits never
 * leave HBM between the two waves: ray i with a primitive-id hit (HGB_HIT_PRIM_ID) becomes a cosine-weighted
 * diffuse bounce (origin = hit point + offset x unit normal turned towards the ray, tmin 0, tmax `tmax`,
 * direction from a counter-based generator keyed by (seed, i)); a ray that missed is emitted again unchanged.
 * `dev_out` may equal `dev_rays`. IEEE arithmetic only; oracle/hagrid_oracle.c (og_bounce_rays) gives the same
 * bits on the CPU. The reference build of this ABI reports an error. */
HGB_API int  hgb_generate_bounce_rays(hgb_scene* scene, const void* dev_rays, const void* dev_hits, int num_rays,
                                      float offset, float tmax, unsigned seed, void* dev_out);
/* Same, for a shard of a frame: `dev_keys` (device, one int32 per ray, NULL = the ray's index in this buffer) names each
 * ray's random stream. A rank that passes the indices its rays have in the whole frame gets exactly the second-wave
 * rays the unsharded frame would get, so the gathered frame does not depend on the number of GPUs. */
HGB_API int  hgb_generate_bounce_rays_keyed(hgb_scene* scene, const void* dev_rays, const void* dev_hits, int num_rays,
                                            float offset, float tmax, unsigned seed, const void* dev_keys, void* dev_out);
/* Per-frame counters of a buffer of primitive-id hits, the only thing the ranks of a sharded frame exchange
 * (one all-reduce per frame): dev_counters[0] += hits with id >= 0, dev_counters[1] += sum of (id + 1);
 * two uint64 in device memory, zeroed by the caller. Asynchronous on the legacy default stream. */
HGB_API int  hgb_count_hits(hgb_scene* scene, const void* dev_hits, int num_hits, void* dev_counters);
/* One two-wave frame (config C5) with everything resident in HBM, in one call: primary rays -> dev_hits_primary
 * (primitive ids), their bounce rays (hgb_generate_bounce_rays_keyed) -> dev_bounce_rays, those traced ->
 * dev_hits_bounce; dev_counters (two uint64, zeroed by the caller, may be NULL) receives hgb_count_hits of both hit
 * buffers. Same results as the five separate calls. Ordered after earlier work on the legacy default stream and
 * joined back into it; asynchronous. The reference build of this ABI reports an error. */
HGB_API int  hgb_trace_two_waves(hgb_scene* scene, const void* dev_rays, int num_rays, const void* dev_keys,
                                 float offset, float tmax, unsigned seed, void* dev_hits_primary,
                                 void* dev_bounce_rays, void* dev_hits_bounce, void* dev_counters);
/* One two-wave frame (config C5) with HOST buffers: uploads `num_rays` primary rays, traces them, makes the bounce
 * rays on the device from the resident rays and hits (hgb_generate_bounce_rays_keyed), traces those, and returns
 * when both hit buffers (primitive ids) are complete in host memory. Upload, traversals and downloads overlap in
 * chunks; the second wave never crosses PCIe as rays (the reference's front end would download the hits, generate on
 * one CPU thread and upload 32 B per ray). The reference build of this ABI reports an error. */
HGB_API int  hgb_trace_two_waves_host(hgb_scene* scene, const void* host_rays, int num_rays, const void* dev_keys,
                                      float offset, float tmax, unsigned seed, void* host_hits_primary,
                                      void* host_hits_bounce);

/* Device-wide primitives on their own: what the reference's `Parallel` wrapper offers on top of CUB
 * (src/parallel.cuh:12-89 -- scan :29-41, reduce :43-55, partition :57-71, sort_pairs :73-86), served by this library's
 * hand-written kernels. The construction pipeline uses them fused with its own functors; these entry points make them
 * testable and measurable in isolation (tests/test_primitives.py, tools/gpu_primitives_bench.py vs the toolkit's CUB).
 * All pointers are device pointers; scratch comes from the scene's pool; the reference build reports an error.
 *   hgb_prim_exclusive_scan  dev_out[i] = sum of dev_in[0, i), i in [0, n] (n + 1 outputs, may alias the input);
 *                            elem_bytes 4 = int32, 8 = two packed 32-bit counters per element
 *   hgb_prim_reduce          op 0 = sum of int32, 1 = max of int32, 2 = min of float, 3 = max of float -> 4 bytes
 *   hgb_prim_partition       cub::DevicePartition::Flagged order: flagged items first in input order, the others behind
 *                            them in REVERSE input order (src/build.cu:568-569 relies on it); returns the number kept
 *   hgb_prim_sort_pairs      stable sort of int32 (key, value) pairs by the low `bits` bits of the keys, in place */
HGB_API int  hgb_prim_exclusive_scan(hgb_scene* scene, const void* dev_in, int n, int elem_bytes, void* dev_out);
HGB_API int  hgb_prim_reduce(hgb_scene* scene, const void* dev_in, int n, int op, void* dev_out);
HGB_API int  hgb_prim_partition(hgb_scene* scene, const void* dev_in, const void* dev_flags, int n, void* dev_out);
HGB_API int  hgb_prim_sort_pairs(hgb_scene* scene, void* dev_keys, void* dev_vals, int n, int bits);

/* Headless stand-in for the viewer's SDL window (src/main.cpp:558-625): a frame of BGRA words as written by
 * hgb_render_frame / update_surface goes to a binary PPM file (P6). */
HGB_API int  hgb_save_image(const char* path, const void* host_bgra, int width, int height);

/* On-disk formats. `.rays` is the reference's ray file (6 float32 per ray: org, dir; tmin / tmax come from the
 * caller; load_rays, src/main.cpp:277-300). hgb_load_rays fills a DEVICE buffer of hgb_rays_file_count() rays:
 * this library uploads the 24-byte records and expands them on the device, the reference build of this ABI runs
 * the reference's own load_rays and uploads 32-byte rays. `.hgrid` is this library's grid cache (the reference
 * has no serialisation); hgb_grid_load replaces the scene's grid, arrays come from its MemManager. */
HGB_API long long hgb_rays_file_count(const char* path);                 /* -1: cannot open */
HGB_API long long hgb_load_rays(hgb_scene* scene, const char* path, float tmin, float tmax, void* dev_rays);
HGB_API int  hgb_save_rays(hgb_scene* scene, const char* path, const void* dev_rays, long long count);
HGB_API int  hgb_grid_save(hgb_scene* scene, const char* path);
HGB_API int  hgb_grid_load(hgb_scene* scene, const char* path);

/* Grid inspection / transplant (parity tests move a grid between the
 * reference build and this library through host memory). */
HGB_API int  hgb_grid_get_info(const hgb_scene* scene, hgb_grid_info* info);
HGB_API int  hgb_grid_download(const hgb_scene* scene, int which, void* host_dst, size_t bytes);
/* `cells` is Cell[] or, when info->compressed, SmallCell[]. Arrays are taken
 * from the scene's MemManager so the next build can free them. */
HGB_API int  hgb_grid_upload(hgb_scene* scene, const hgb_grid_info* info,
                     const void* host_entries, const void* host_cells,
                     const void* host_refs);

/* Raw device buffers from the scene's MemManager (mem_manager.h:46,75,84). */
HGB_API void* hgb_device_alloc(hgb_scene* scene, size_t bytes);
HGB_API void  hgb_device_free(hgb_scene* scene, void* dev_ptr);
HGB_API int   hgb_copy_to_device(hgb_scene* scene, void* dev_dst, const void* host_src, size_t bytes);
HGB_API int   hgb_copy_to_host(hgb_scene* scene, void* host_dst, const void* dev_src, size_t bytes);
HGB_API int   hgb_device_synchronize(void);
/* Kernels launched by this library since load (0 for the reference build, which
 * does not count). bench.py reports the difference over its timed region. */
HGB_API unsigned long long hgb_kernel_launches(void);

#ifdef __cplusplus
}
#endif

#endif /* HAGRID_B200_H */
