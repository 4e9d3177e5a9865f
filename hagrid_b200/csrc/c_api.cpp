// C ABI (include/hagrid_b200.h) over the `namespace hagrid` C++ API.
//
// This file only uses what the reference's own front end uses
// (src/main.cpp:471-533: MemManager, Grid, the five build stages,
// setup_traversal/traverse_grid and profile()), so it builds unchanged against
//   * this repository's headers + kernels  -> libhagrid_b200.so  (the product)
//   * /root/reference/src headers + objects -> oracle/_ref/libhagrid_ref.so
//     (with -DHGB_REFERENCE_BUILD; test oracle and reference bench arm only).
// Host-only translation unit: compiled with g++ -DHOST= -DDEVICE= exactly like
// the reference compiles main.cpp (src/CMakeLists.txt:41-43).

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include "build.h"
#include "traverse.h"
#include "mem_manager.h"

#include "hagrid_b200.h"

#ifdef HGB_REFERENCE_BUILD
// The reference's own gen_camera / gen_rays / update_surface (oracle/ref_frontend.cpp includes src/main.cpp)
extern "C" void hgb_ref_gen_camera(const float* eye, const float* center, const float* up, float fov, float ratio, float* cam12);
extern "C" const void* hgb_ref_gen_rays(const float* cam12, float clip, int w, int h);
extern "C" void hgb_ref_update_surface(int mode, const void* hits, float clip, int w, int h, void* bgra);
extern "C" const void* hgb_ref_load_model(const char* path, int* num_tris);
extern "C" const void* hgb_ref_load_rays(const char* path, float tmin, float tmax, long long* count);
#else
#include "formats.h"
#include "scene_ingest.h"
#endif

namespace hagrid {
#ifdef HGB_REFERENCE_BUILD
// Second build of the reference's traverse.cu with the `hit.id = steps` line
// (src/traverse.cu:93) removed; see oracle/build_ref.sh.
void setup_traversal_pid(const Grid& grid);
void traverse_grid_pid(const Grid& grid, const Tri* tris, const Ray* rays, Hit* hits, int num_rays);
#else
// Extra entry points of this library (hagrid_b200/include/hagrid/traverse.h).
void traverse_grid_prim_ids(const Grid& grid, const Tri* tris, const Ray* rays, Hit* hits, int num_rays);
void traverse_grid_host(const Grid& grid, const Tri* tris, const Ray* host_rays, Hit* host_hits, int num_rays,
                        Ray* dev_rays, Hit* dev_hits, bool prim_ids);
void traverse_grid_to_host(const Grid& grid, const Tri* tris, const Ray* dev_rays, Hit* dev_hits, Hit* host_hits, int num_rays,
                           bool prim_ids);
void trace_two_waves(const Grid& grid, const Tri* tris, int num_tris, const Ray* rays, int num_rays, const int* keys,
                     float offset, float tmax, unsigned seed, Hit* hits_primary, Ray* bounce, Hit* hits_bounce,
                     unsigned long long* counters);
void prim_exclusive_scan(MemManager& mem, const void* in, int n, int elem_bytes, void* out);
void prim_reduce(MemManager& mem, const void* in, int n, int op, void* out);
int prim_partition(MemManager& mem, const int* in, const int* flags, int n, int* out);
void prim_sort_pairs(MemManager& mem, int* keys, int* vals, int n, int bits);
bool set_traversal_option(const char* key, int value);
bool set_merge_option(const char* key, int value);
int debug_tile_costs(const void* rays, int num_rays, unsigned short* out, int capacity);
unsigned long long kernel_launch_count();
void trim_device_pool();
#endif
}

using namespace hagrid;

struct hgb_scene {
    MemManager mem;
    Grid grid;
    Tri* tris;
    int num_tris;
    int device;
    Ray* frame_rays;
    Hit* frame_hits;
    int frame_capacity;
    unsigned* frame_pixels;
    int pixel_capacity;
    unsigned long long grid_epoch;      // bumped by everything that changes the grid or the triangles
    unsigned long long setup_epoch;     // grid_epoch at the last hgb_setup_traversal (~0: never)

    hgb_scene(int dev, bool keep)
        : mem(keep), tris(nullptr), num_tris(0), device(dev),
          frame_rays(nullptr), frame_hits(nullptr), frame_capacity(0), frame_pixels(nullptr), pixel_capacity(0), grid_epoch(0), setup_epoch(~0ull)
    {
        grid.entries = nullptr;
        grid.ref_ids = nullptr;
        grid.cells = nullptr;
        grid.small_cells = nullptr;
        grid.num_cells = grid.num_entries = grid.num_refs = grid.shift = 0;
        grid.dims = ivec3(0, 0, 0);
        grid.bbox = BBox(vec3(0, 0, 0), vec3(0, 0, 0));
    }
};

static thread_local std::string g_error;

static int fail(const char* msg) {
    g_error = msg;
    return -1;
}

static bool bind(const hgb_scene* scene) {
    if (!scene) { g_error = "null scene"; return false; }
    if (cudaSetDevice(scene->device) != cudaSuccess) {
        g_error = "cudaSetDevice failed";
        cudaGetLastError();
        return false;
    }
    return true;
}

static void release_grid(hgb_scene* s) {
    s->grid_epoch++;
    // main.cpp:496-498 frees exactly these three; small_cells is leaked by the
    // reference across rebuilds (build.cu:755). Freeing it here is harmless for
    // both builds because it always comes from the same MemManager.
    s->mem.free(s->grid.entries);
    s->mem.free(s->grid.cells);
    s->mem.free(s->grid.ref_ids);
    s->mem.free(s->grid.small_cells);
    s->grid.entries = nullptr;
    s->grid.cells = nullptr;
    s->grid.ref_ids = nullptr;
    s->grid.small_cells = nullptr;
}

// hgb_setup_traversal must follow every change of a scene's grid or triangles (the reference's call order,
// src/main.cpp:536-549): tracing a scene whose setup is stale is an error code. The state is per scene --
// any number of scenes, on any devices, can be set up and traced side by side.
#ifdef HGB_REFERENCE_BUILD
// ... except in the reference build, whose traversal constants are per process (src/traverse.cu:7-12)
static const hgb_scene* g_ref_setup_scene = nullptr;
static bool setup_matches(const hgb_scene* s) { return g_ref_setup_scene == s && s->setup_epoch == s->grid_epoch; }
#else
static bool setup_matches(const hgb_scene* s) { return s->setup_epoch == s->grid_epoch; }
#endif

/// Device staging buffers of a host-buffer frame, from the scene's pool
static void reserve_frame(hgb_scene* s, int num_rays) {
    if (num_rays <= s->frame_capacity) return;
    s->mem.free(s->frame_rays);
    s->mem.free(s->frame_hits);
    s->frame_rays = s->mem.alloc<Ray>(num_rays);
    s->frame_hits = s->mem.alloc<Hit>(num_rays);
    s->frame_capacity = num_rays;
}

static void run_traverse(hgb_scene* s, const Ray* rays, Hit* hits, int n, int hit_mode) {
    if (hit_mode == HGB_HIT_PRIM_ID) {
#ifdef HGB_REFERENCE_BUILD
        traverse_grid_pid(s->grid, s->tris, rays, hits, n);
#else
        traverse_grid_prim_ids(s->grid, s->tris, rays, hits, n);
#endif
    } else {
        traverse_grid(s->grid, s->tris, rays, hits, n);
    }
}

extern "C" {

const char* hgb_impl(void) {
#ifdef HGB_REFERENCE_BUILD
    return "reference";
#else
    return "hagrid_b200";
#endif
}

const char* hgb_last_error(void) { return g_error.c_str(); }

int hgb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int hgb_set_option(const char* key, int value) {
    if (!key) return fail("set_option: null key");
#ifdef HGB_REFERENCE_BUILD
    (void)value;
    return 0;
#else
    return set_traversal_option(key, value) || set_merge_option(key, value) ? 0 : fail("set_option: unknown key");
#endif
}

int hgb_tile_costs(const void* dev_rays, int num_rays, unsigned short* host_costs, int capacity) {
#ifdef HGB_REFERENCE_BUILD
    (void)dev_rays; (void)num_rays; (void)host_costs; (void)capacity;
    return 0;
#else
    if (!dev_rays || !host_costs || num_rays <= 0 || capacity <= 0) return 0;
    return debug_tile_costs(dev_rays, num_rays, host_costs, capacity);
#endif
}

hgb_scene* hgb_scene_create(int device, int keep_alive) {
    if (device < 0 || device >= hgb_device_count()) {
        g_error = "no such CUDA device (this library has no CPU fallback)";
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { g_error = "cudaSetDevice failed"; return nullptr; }
    return new hgb_scene(device, keep_alive != 0);
}

void hgb_scene_destroy(hgb_scene* s) {
    if (!s) return;
#ifdef HGB_REFERENCE_BUILD
    if (g_ref_setup_scene == s) g_ref_setup_scene = nullptr;
#endif
    const bool bound = bind(s);
    if (bound) {
        cudaDeviceSynchronize();
        release_grid(s);
        s->mem.free(s->tris);
        s->mem.free(s->frame_rays);
        s->mem.free(s->frame_hits);
        s->mem.free(s->frame_pixels);
    }
    delete s;               // ~MemManager hands every slot back, also what keep-alive mode retained
#ifndef HGB_REFERENCE_BUILD
    if (bound) trim_device_pool();
#endif
}

int hgb_scene_set_tris(hgb_scene* s, const void* host_tris, int num_tris) {
    if (!bind(s)) return -1;
    if (!host_tris || num_tris <= 0) return fail("set_tris: empty triangle array");
    s->grid_epoch++;            // the grid indexes the old triangle array: stale until rebuilt and set up again
    s->mem.free(s->tris);
    s->tris = s->mem.alloc<Tri>(num_tris);
    s->mem.copy<Copy::HST_TO_DEV>(s->tris, static_cast<const Tri*>(host_tris), num_tris);
    s->num_tris = num_tris;
    return 0;
}

int hgb_scene_load_obj(hgb_scene* s, const char* path, int threads) {
    if (!bind(s)) return -1;
    if (!path) return fail("load_obj: null path");
    s->grid_epoch++;
#ifdef HGB_REFERENCE_BUILD
    (void)threads;
    int n = 0;
    const Tri* host = static_cast<const Tri*>(hgb_ref_load_model(path, &n));
    if (!host || n <= 0) return fail("load_obj: the reference's load_model refused the file");
    s->mem.free(s->tris);
    s->tris = s->mem.alloc<Tri>(n);
    s->mem.copy<Copy::HST_TO_DEV>(s->tris, host, n);
    s->num_tris = n;
    return n;
#else
    ObjGeometry geo;
    if (!parse_obj(path, threads, geo)) { g_error = "load_obj: " + geo.error; return -1; }
    const int n = int(geo.indices.size() / 3);
    if (n <= 0) return fail("load_obj: no triangles");
    vec3* dev_vertices = s->mem.alloc<vec3>(geo.vertices.size());
    int* dev_indices = s->mem.alloc<int>(geo.indices.size());
    s->mem.copy<Copy::HST_TO_DEV>(dev_vertices, geo.vertices.data(), geo.vertices.size());
    s->mem.copy<Copy::HST_TO_DEV>(dev_indices, geo.indices.data(), geo.indices.size());
    s->mem.free(s->tris);
    s->tris = s->mem.alloc<Tri>(n);
    setup_triangles(dev_vertices, dev_indices, n, s->tris);
    cudaDeviceSynchronize();
    s->mem.free(dev_vertices);
    s->mem.free(dev_indices);
    s->num_tris = n;
    return n;
#endif
}

#ifdef HGB_REFERENCE_BUILD
struct hgb_obj { int unused; };
hgb_obj* hgb_obj_parse(const char*, int) { g_error = "obj_parse: not part of the reference build"; return nullptr; }
int hgb_obj_num_vertices(const hgb_obj*) { return 0; }
int hgb_obj_num_tris(const hgb_obj*) { return 0; }
const float* hgb_obj_vertices(const hgb_obj*) { return nullptr; }
const int* hgb_obj_indices(const hgb_obj*) { return nullptr; }
void hgb_obj_free(hgb_obj*) {}
#else
struct hgb_obj { ObjGeometry geo; };
hgb_obj* hgb_obj_parse(const char* path, int threads) {
    if (!path) { g_error = "obj_parse: null path"; return nullptr; }
    hgb_obj* obj = new hgb_obj;
    if (!parse_obj(path, threads, obj->geo)) { g_error = "obj_parse: " + obj->geo.error; delete obj; return nullptr; }
    return obj;
}
int hgb_obj_num_vertices(const hgb_obj* obj) { return obj ? int(obj->geo.vertices.size()) : 0; }
int hgb_obj_num_tris(const hgb_obj* obj) { return obj ? int(obj->geo.indices.size() / 3) : 0; }
const float* hgb_obj_vertices(const hgb_obj* obj) { return obj ? reinterpret_cast<const float*>(obj->geo.vertices.data()) : nullptr; }
const int* hgb_obj_indices(const hgb_obj* obj) { return obj ? obj->geo.indices.data() : nullptr; }
void hgb_obj_free(hgb_obj* obj) { delete obj; }
#endif

int hgb_scene_num_tris(const hgb_scene* s) { return s ? s->num_tris : 0; }
size_t hgb_scene_peak_bytes(const hgb_scene* s) { return s ? s->mem.max_usage() : 0; }

int hgb_build_grid(hgb_scene* s, float top_density, float snd_density) {
    if (!bind(s)) return -1;
    if (!s->tris) return fail("build_grid: no triangles");
    release_grid(s);
    build_grid(s->mem, s->tris, s->num_tris, s->grid, top_density, snd_density);
    return 0;
}

int hgb_merge_grid(hgb_scene* s, float alpha) {
    if (!bind(s)) return -1;
    if (!s->grid.cells) return fail("merge_grid: no uncompressed grid");
    s->grid_epoch++;
    merge_grid(s->mem, s->grid, alpha);
    return 0;
}

int hgb_flatten_grid(hgb_scene* s) {
    if (!bind(s)) return -1;
    if (!s->grid.entries) return fail("flatten_grid: no grid");
    s->grid_epoch++;
    flatten_grid(s->mem, s->grid);
    return 0;
}

int hgb_expand_grid(hgb_scene* s, int iters) {
    if (!bind(s)) return -1;
    if (!s->grid.cells) return fail("expand_grid: no uncompressed grid");
    s->grid_epoch++;
    expand_grid(s->mem, s->grid, s->tris, iters);
    return 0;
}

int hgb_compress_grid(hgb_scene* s) {
    if (!bind(s)) return -1;
    if (!s->grid.cells) return fail("compress_grid: no uncompressed grid");
    s->grid_epoch++;
    return compress_grid(s->mem, s->grid) ? 1 : 0;
}

int hgb_build_pipeline(hgb_scene* s, float top_density, float snd_density,
                       float alpha, int exp_iters, int compress,
                       int warmup, int iters, float* ms_out) {
    if (!bind(s)) return -1;
    if (!s->tris) return fail("build_pipeline: no triangles");
    for (int i = 0; i < warmup + iters; i++) {
        release_grid(s);
        float ms = profile([&] {
            build_grid(s->mem, s->tris, s->num_tris, s->grid, top_density, snd_density);
            merge_grid(s->mem, s->grid, alpha);
            flatten_grid(s->mem, s->grid);
            expand_grid(s->mem, s->grid, s->tris, exp_iters);
            if (compress) compress_grid(s->mem, s->grid);
        });
        if (i >= warmup && ms_out) ms_out[i - warmup] = ms;
    }
    return 0;
}

int hgb_setup_traversal(hgb_scene* s) {
    if (!bind(s)) return -1;
    if (!s->grid.entries) return fail("setup_traversal: no grid");
    setup_traversal(s->grid);
#ifdef HGB_REFERENCE_BUILD
    setup_traversal_pid(s->grid);
#endif
#ifdef HGB_REFERENCE_BUILD
    g_ref_setup_scene = s;
#endif
    s->setup_epoch = s->grid_epoch;
    return 0;
}

int hgb_traverse_grid(hgb_scene* s, const void* dev_rays, void* dev_hits, int num_rays, int hit_mode) {
    if (!bind(s)) return -1;
    if (!s->grid.entries) return fail("traverse_grid: no grid");
    if (!setup_matches(s)) return fail("traverse_grid: hgb_setup_traversal was not called for this grid");
    if (num_rays <= 0) return 0;
    run_traverse(s, static_cast<const Ray*>(dev_rays), static_cast<Hit*>(dev_hits), num_rays, hit_mode);
    return 0;
}

int hgb_traverse_timed(hgb_scene* s, const void* dev_rays, void* dev_hits, int num_rays,
                       int hit_mode, int warmup, int iters, float* ms_out) {
    if (!bind(s)) return -1;
    if (!s->grid.entries) return fail("traverse_timed: no grid");
    if (!setup_matches(s)) return fail("traverse_timed: hgb_setup_traversal was not called for this grid");
    if (num_rays <= 0) return fail("traverse_timed: no rays");
    auto rays = static_cast<const Ray*>(dev_rays);
    auto hits = static_cast<Hit*>(dev_hits);
    for (int i = 0; i < warmup; i++) run_traverse(s, rays, hits, num_rays, hit_mode);
    for (int i = 0; i < iters; i++) {
        float ms = profile([&] { run_traverse(s, rays, hits, num_rays, hit_mode); });
        if (ms_out) ms_out[i] = ms;
    }
    return 0;
}

int hgb_traverse_grid_host(hgb_scene* s, const void* host_rays, void* host_hits, int num_rays, int hit_mode) {
    if (!bind(s)) return -1;
    if (!s->grid.entries) return fail("traverse_grid_host: no grid");
    if (!setup_matches(s)) return fail("traverse_grid_host: hgb_setup_traversal was not called for this grid");
    if (num_rays <= 0) return 0;
    reserve_frame(s, num_rays);
#ifdef HGB_REFERENCE_BUILD
    // the reference's frame, verbatim: blocking upload, launch, blocking download (src/main.cpp:599-613)
    s->mem.copy<Copy::HST_TO_DEV>(s->frame_rays, static_cast<const Ray*>(host_rays), num_rays);
    run_traverse(s, s->frame_rays, s->frame_hits, num_rays, hit_mode);
    s->mem.copy<Copy::DEV_TO_HST>(static_cast<Hit*>(host_hits), s->frame_hits, num_rays);
#else
    traverse_grid_host(s->grid, s->tris, static_cast<const Ray*>(host_rays), static_cast<Hit*>(host_hits), num_rays,
                       s->frame_rays, s->frame_hits, hit_mode == HGB_HIT_PRIM_ID);
#endif
    return 0;
}

int hgb_make_camera(const float eye[3], const float center[3], const float up[3], float fov, float ratio, float cam_out[12]) {
    if (!eye || !center || !up || !cam_out) return fail("make_camera: null argument");
#ifdef HGB_REFERENCE_BUILD
    hgb_ref_gen_camera(eye, center, up, fov, ratio, cam_out);
#else
    const FrameCamera cam = make_camera(vec3(eye[0], eye[1], eye[2]), vec3(center[0], center[1], center[2]),
                                        vec3(up[0], up[1], up[2]), fov, ratio);
    const vec3 v[4] = {cam.eye, cam.right, cam.up, cam.dir};
    for (int i = 0; i < 4; i++) { cam_out[3 * i] = v[i].x; cam_out[3 * i + 1] = v[i].y; cam_out[3 * i + 2] = v[i].z; }
#endif
    return 0;
}

#ifndef HGB_REFERENCE_BUILD
static FrameCamera camera_of(const float* c) {
    FrameCamera cam;
    cam.eye = vec3(c[0], c[1], c[2]); cam.right = vec3(c[3], c[4], c[5]);
    cam.up = vec3(c[6], c[7], c[8]); cam.dir = vec3(c[9], c[10], c[11]);
    return cam;
}
#endif

int hgb_generate_rays(hgb_scene* s, const float cam[12], float clip, int width, int height, void* dev_rays) {
    if (!bind(s)) return -1;
    if (!cam || !dev_rays || width <= 0 || height <= 0) return fail("generate_rays: bad argument");
#ifdef HGB_REFERENCE_BUILD
    // the reference generates on the host and uploads (src/main.cpp:598-599)
    const Ray* host = static_cast<const Ray*>(hgb_ref_gen_rays(cam, clip, width, height));
    s->mem.copy<Copy::HST_TO_DEV>(static_cast<Ray*>(dev_rays), host, size_t(width) * height);
#else
    generate_rays(camera_of(cam), clip, width, height, static_cast<Ray*>(dev_rays));
#endif
    return 0;
}

int hgb_generate_bounce_rays(hgb_scene* s, const void* dev_rays, const void* dev_hits, int num_rays, float offset,
                             float tmax, unsigned seed, void* dev_out) {
    if (!bind(s)) return -1;
    if (num_rays < 0 || (num_rays > 0 && (!dev_rays || !dev_hits || !dev_out))) return fail("generate_bounce_rays: bad argument");
    if (!s->tris) return fail("generate_bounce_rays: the scene has no triangles");
#ifdef HGB_REFERENCE_BUILD
    return fail("generate_bounce_rays: the reference has no second-wave ray generation");
#else
    generate_bounce_rays(s->tris, s->num_tris, static_cast<const Ray*>(dev_rays), static_cast<const Hit*>(dev_hits),
                         num_rays, offset, tmax, seed, static_cast<Ray*>(dev_out));
    return 0;
#endif
}

int hgb_generate_bounce_rays_keyed(hgb_scene* s, const void* dev_rays, const void* dev_hits, int num_rays, float offset,
                                   float tmax, unsigned seed, const void* dev_keys, void* dev_out) {
    if (!bind(s)) return -1;
    if (num_rays < 0 || (num_rays > 0 && (!dev_rays || !dev_hits || !dev_out))) return fail("generate_bounce_rays_keyed: bad argument");
    if (!s->tris) return fail("generate_bounce_rays_keyed: the scene has no triangles");
#ifdef HGB_REFERENCE_BUILD
    (void)offset; (void)tmax; (void)seed; (void)dev_keys;
    return fail("generate_bounce_rays_keyed: the reference has no second-wave ray generation");
#else
    generate_bounce_rays(s->tris, s->num_tris, static_cast<const Ray*>(dev_rays), static_cast<const Hit*>(dev_hits),
                         num_rays, offset, tmax, seed, static_cast<Ray*>(dev_out), static_cast<const int*>(dev_keys));
    return 0;
#endif
}

int hgb_count_hits(hgb_scene* s, const void* dev_hits, int num_hits, void* dev_counters) {
    if (!bind(s)) return -1;
    if (num_hits < 0 || (num_hits > 0 && !dev_hits) || !dev_counters) return fail("count_hits: bad argument");
#ifdef HGB_REFERENCE_BUILD
    return fail("count_hits: not part of the reference");
#else
    count_hits(static_cast<const Hit*>(dev_hits), num_hits, static_cast<unsigned long long*>(dev_counters));
    return 0;
#endif
}

int hgb_trace_two_waves(hgb_scene* s, const void* dev_rays, int num_rays, const void* dev_keys, float offset, float tmax,
                        unsigned seed, void* dev_hits_primary, void* dev_bounce_rays, void* dev_hits_bounce, void* dev_counters) {
    if (!bind(s)) return -1;
    if (!s->grid.entries) return fail("trace_two_waves: no grid");
    if (!setup_matches(s)) return fail("trace_two_waves: hgb_setup_traversal was not called for this grid");
    if (num_rays < 0 || (num_rays > 0 && (!dev_rays || !dev_hits_primary || !dev_bounce_rays || !dev_hits_bounce)))
        return fail("trace_two_waves: bad argument");
#ifdef HGB_REFERENCE_BUILD
    (void)dev_keys; (void)offset; (void)tmax; (void)seed; (void)dev_counters;
    return fail("trace_two_waves: the reference has no second-wave ray generation");
#else
    trace_two_waves(s->grid, s->tris, s->num_tris, static_cast<const Ray*>(dev_rays), num_rays, static_cast<const int*>(dev_keys),
                    offset, tmax, seed, static_cast<Hit*>(dev_hits_primary), static_cast<Ray*>(dev_bounce_rays),
                    static_cast<Hit*>(dev_hits_bounce), static_cast<unsigned long long*>(dev_counters));
    return 0;
#endif
}

int hgb_trace_two_waves_host(hgb_scene* s, const void* host_rays, int num_rays, const void* dev_keys, float offset, float tmax,
                             unsigned seed, void* host_hits_primary, void* host_hits_bounce) {
    if (!bind(s)) return -1;
    if (!s->grid.entries) return fail("trace_two_waves_host: no grid");
    if (!setup_matches(s)) return fail("trace_two_waves_host: hgb_setup_traversal was not called for this grid");
    if (num_rays < 0 || (num_rays > 0 && (!host_rays || !host_hits_primary || !host_hits_bounce))) return fail("trace_two_waves_host: bad argument");
    if (num_rays == 0) return 0;
#ifdef HGB_REFERENCE_BUILD
    (void)dev_keys; (void)offset; (void)tmax; (void)seed;
    return fail("trace_two_waves_host: the reference has no second-wave ray generation");
#else
    reserve_frame(s, num_rays);
    // first wave: upload, trace, download in overlapping chunks; the rays and their hits stay in the staging buffers
    traverse_grid_host(s->grid, s->tris, static_cast<const Ray*>(host_rays), static_cast<Hit*>(host_hits_primary), num_rays,
                       s->frame_rays, s->frame_hits, true);
    // second wave: made on the device from what is resident (in place), traced, downloaded chunk by chunk
    generate_bounce_rays(s->tris, s->num_tris, s->frame_rays, s->frame_hits, num_rays, offset, tmax, seed, s->frame_rays,
                         static_cast<const int*>(dev_keys));
    traverse_grid_to_host(s->grid, s->tris, s->frame_rays, s->frame_hits, static_cast<Hit*>(host_hits_bounce), num_rays, true);
    return 0;
#endif
}

int hgb_render_frame(hgb_scene* s, const float cam[12], float clip, int width, int height, int display_mode, void* host_bgra) {
    if (!bind(s)) return -1;
    if (!s->grid.entries) return fail("render_frame: no grid");
    if (!setup_matches(s)) return fail("render_frame: hgb_setup_traversal was not called for this grid");
    if (!cam || !host_bgra || width <= 0 || height <= 0) return fail("render_frame: bad argument");
    if (display_mode < 0 || display_mode > 2) return fail("render_frame: display_mode must be 0, 1 or 2");
    const int n = width * height;
#ifdef HGB_REFERENCE_BUILD
    // one iteration of the reference's viewer loop, with its own functions (src/main.cpp:598-621)
    static std::vector<Hit> host_hits;
    host_hits.resize(n);
    reserve_frame(s, n);
    const Ray* host_rays = static_cast<const Ray*>(hgb_ref_gen_rays(cam, clip, width, height));
    s->mem.copy<Copy::HST_TO_DEV>(s->frame_rays, host_rays, n);
    traverse_grid(s->grid, s->tris, s->frame_rays, s->frame_hits, n);
    s->mem.copy<Copy::DEV_TO_HST>(host_hits.data(), s->frame_hits, n);
    hgb_ref_update_surface(display_mode, host_hits.data(), clip, width, height, host_bgra);
#else
    if (n > s->pixel_capacity) {
        s->mem.free(s->frame_pixels);
        s->frame_pixels = s->mem.alloc<unsigned>(n);
        s->pixel_capacity = n;
    }
    render_frame(s->grid, s->tris, camera_of(cam), clip, width, height, display_mode, s->frame_pixels);
    s->mem.copy<Copy::DEV_TO_HST>(static_cast<unsigned*>(host_bgra), s->frame_pixels, n);
#endif
    return 0;
}

long long hgb_rays_file_count(const char* path) {
    if (!path) return -1;
    std::FILE* fp = std::fopen(path, "rb");
    if (!fp) return -1;
    std::fseek(fp, 0, SEEK_END);
    const long long n = std::ftell(fp) / 24;                  // src/main.cpp:282
    std::fclose(fp);
    return n;
}

long long hgb_load_rays(hgb_scene* s, const char* path, float tmin, float tmax, void* dev_rays) {
    if (!bind(s)) return -1;
    if (!path || !dev_rays) return fail("load_rays: null argument");
#ifdef HGB_REFERENCE_BUILD
    long long n = 0;
    const Ray* host = static_cast<const Ray*>(hgb_ref_load_rays(path, tmin, tmax, &n));
    if (!host) return fail("load_rays: cannot load ray file");
    if (n > 0) s->mem.copy<Copy::HST_TO_DEV>(static_cast<Ray*>(dev_rays), host, size_t(n));
    return n;
#else
    const long long n = load_rays_to_device(s->mem, path, tmin, tmax, static_cast<Ray*>(dev_rays));
    if (n < 0) return fail("load_rays: cannot load ray file");
    return n;
#endif
}

int hgb_save_rays(hgb_scene* s, const char* path, const void* dev_rays, long long count) {
    if (!bind(s)) return -1;
    if (!path || (!dev_rays && count > 0) || count < 0) return fail("save_rays: bad argument");
#ifdef HGB_REFERENCE_BUILD
    return fail("save_rays: the reference has no ray file writer");
#else
    return save_rays_from_device(s->mem, path, static_cast<const Ray*>(dev_rays), count) ? 0 : fail("save_rays: cannot write file");
#endif
}

int hgb_save_image(const char* path, const void* host_bgra, int width, int height) {
    if (!path || !host_bgra || width <= 0 || height <= 0) return fail("save_image: bad argument");
#ifdef HGB_REFERENCE_BUILD
    return fail("save_image: the reference only draws into an SDL window");
#else
    return save_image_ppm(path, static_cast<const unsigned char*>(host_bgra), width, height) ? 0 : fail("save_image: cannot write file");
#endif
}

int hgb_grid_save(hgb_scene* s, const char* path) {
    if (!bind(s)) return -1;
    if (!path) return fail("grid_save: null path");
#ifdef HGB_REFERENCE_BUILD
    return fail("grid_save: the reference has no grid serialisation");
#else
    std::string err;
    if (!save_grid(s->mem, path, s->grid, err)) { g_error = "grid_save: " + err; return -1; }
    return 0;
#endif
}

int hgb_grid_load(hgb_scene* s, const char* path) {
    if (!bind(s)) return -1;
    if (!path) return fail("grid_load: null path");
#ifdef HGB_REFERENCE_BUILD
    return fail("grid_load: the reference has no grid serialisation");
#else
    Grid loaded;
    loaded.entries = nullptr; loaded.cells = nullptr; loaded.small_cells = nullptr; loaded.ref_ids = nullptr;
    std::string err;
    if (!load_grid(s->mem, path, loaded, err)) { g_error = "grid_load: " + err; return -1; }
    release_grid(s);
    s->grid = loaded;
    return 0;
#endif
}

int hgb_grid_get_info(const hgb_scene* s, hgb_grid_info* info) {
    if (!s || !info) return fail("grid_get_info: null argument");
    const Grid& g = s->grid;
    std::memset(info, 0, sizeof(*info));
    info->bbox_min[0] = g.bbox.min.x; info->bbox_min[1] = g.bbox.min.y; info->bbox_min[2] = g.bbox.min.z;
    info->bbox_max[0] = g.bbox.max.x; info->bbox_max[1] = g.bbox.max.y; info->bbox_max[2] = g.bbox.max.z;
    info->dims[0] = g.dims.x; info->dims[1] = g.dims.y; info->dims[2] = g.dims.z;
    info->shift = g.shift;
    info->num_cells = g.num_cells;
    info->num_entries = g.num_entries;
    info->num_refs = g.num_refs;
    info->compressed = g.small_cells ? 1 : 0;
    if (g.offsets.size() > HGB_MAX_LEVELS) return fail("grid_get_info: too many levels");
    info->num_offsets = static_cast<int32_t>(g.offsets.size());
    for (size_t i = 0; i < g.offsets.size(); i++) info->offsets[i] = g.offsets[i];
    return 0;
}

int hgb_grid_download(const hgb_scene* cs, int which, void* host_dst, size_t bytes) {
    auto s = const_cast<hgb_scene*>(cs);
    if (!bind(s)) return -1;
    if (!host_dst) return fail("grid_download: null destination");
    const Grid& g = s->grid;
    const void* src = nullptr;
    size_t have = 0;
    switch (which) {
        case HGB_ARRAY_ENTRIES:     src = g.entries;     have = size_t(g.num_entries) * sizeof(Entry); break;
        case HGB_ARRAY_CELLS:       src = g.cells;       have = size_t(g.num_cells) * sizeof(Cell); break;
        case HGB_ARRAY_SMALL_CELLS: src = g.small_cells; have = size_t(g.num_cells) * sizeof(SmallCell); break;
        case HGB_ARRAY_REFS:        src = g.ref_ids;     have = size_t(g.num_refs) * sizeof(int); break;
        case HGB_ARRAY_TRIS:        src = s->tris;       have = size_t(s->num_tris) * sizeof(Tri); break;
        default: return fail("grid_download: unknown array");
    }
    if (!src) return fail("grid_download: array not present");
    if (bytes != have) return fail("grid_download: size mismatch");
    if (bytes) s->mem.copy<Copy::DEV_TO_HST>(static_cast<char*>(host_dst), static_cast<const char*>(src), bytes);
    return 0;
}

int hgb_grid_upload(hgb_scene* s, const hgb_grid_info* info,
                    const void* host_entries, const void* host_cells, const void* host_refs) {
    if (!bind(s)) return -1;
    if (!info || !host_entries || !host_cells) return fail("grid_upload: null argument");
    if (info->num_offsets < 0 || info->num_offsets > HGB_MAX_LEVELS) return fail("grid_upload: bad offsets");
    // validate before the scene is touched: a bad header must leave the old grid in place
    if (info->num_cells <= 0 || info->num_entries <= 0 || info->num_refs < 0) return fail("grid_upload: bad counts");
    if (info->num_refs > 0 && !host_refs) return fail("grid_upload: null reference array");
    if (info->shift < 0 || info->shift > 20) return fail("grid_upload: bad shift");
    long long top = 1;
    for (int k = 0; k < 3; k++) {
        if (info->dims[k] <= 0 || ((long long)info->dims[k] << info->shift) > (1ll << 30)) return fail("grid_upload: bad dims");
        if (!(info->bbox_max[k] > info->bbox_min[k])) return fail("grid_upload: empty bounding box");
        top *= info->dims[k];
    }
    if (top > info->num_entries) return fail("grid_upload: fewer entries than top-level cells");
    for (int i = 0; i < info->num_offsets; i++)
        if (info->offsets[i] < 0 || info->offsets[i] > info->num_entries || (i > 0 && info->offsets[i] < info->offsets[i - 1]))
            return fail("grid_upload: bad offsets");
    release_grid(s);
    Grid& g = s->grid;
    g.bbox = BBox(vec3(info->bbox_min[0], info->bbox_min[1], info->bbox_min[2]),
                  vec3(info->bbox_max[0], info->bbox_max[1], info->bbox_max[2]));
    g.dims = ivec3(info->dims[0], info->dims[1], info->dims[2]);
    g.shift = info->shift;
    g.num_cells = info->num_cells;
    g.num_entries = info->num_entries;
    g.num_refs = info->num_refs;
    g.offsets.assign(info->offsets, info->offsets + info->num_offsets);

    g.entries = s->mem.alloc<Entry>(g.num_entries);
    s->mem.copy<Copy::HST_TO_DEV>(g.entries, static_cast<const Entry*>(host_entries), g.num_entries);
    // One spare element so an empty reference array still owns a slot.
    g.ref_ids = s->mem.alloc<int>(size_t(g.num_refs) + 1);
    if (g.num_refs) s->mem.copy<Copy::HST_TO_DEV>(g.ref_ids, static_cast<const int*>(host_refs), g.num_refs);
    if (info->compressed) {
        g.small_cells = s->mem.alloc<SmallCell>(g.num_cells);
        s->mem.copy<Copy::HST_TO_DEV>(g.small_cells, static_cast<const SmallCell*>(host_cells), g.num_cells);
    } else {
        g.cells = s->mem.alloc<Cell>(g.num_cells);
        s->mem.copy<Copy::HST_TO_DEV>(g.cells, static_cast<const Cell*>(host_cells), g.num_cells);
    }
    return 0;
}

void* hgb_device_alloc(hgb_scene* s, size_t bytes) {
    if (!bind(s) || bytes == 0) return nullptr;
    return s->mem.alloc<char>(bytes);
}

void hgb_device_free(hgb_scene* s, void* dev_ptr) {
    if (!bind(s) || !dev_ptr) return;
    s->mem.free(static_cast<char*>(dev_ptr));
}

int hgb_copy_to_device(hgb_scene* s, void* dev_dst, const void* host_src, size_t bytes) {
    if (!bind(s)) return -1;
    if (bytes) s->mem.copy<Copy::HST_TO_DEV>(static_cast<char*>(dev_dst), static_cast<const char*>(host_src), bytes);
    return 0;
}

int hgb_copy_to_host(hgb_scene* s, void* host_dst, const void* dev_src, size_t bytes) {
    if (!bind(s)) return -1;
    if (bytes) s->mem.copy<Copy::DEV_TO_HST>(static_cast<char*>(host_dst), static_cast<const char*>(dev_src), bytes);
    return 0;
}

#ifdef HGB_REFERENCE_BUILD
int hgb_prim_exclusive_scan(hgb_scene*, const void*, int, int, void*) { return fail("prim: the reference build wraps CUB (src/parallel.cuh), nothing of its own to call"); }
int hgb_prim_reduce(hgb_scene*, const void*, int, int, void*) { return fail("prim: not part of the reference build"); }
int hgb_prim_partition(hgb_scene*, const void*, const void*, int, void*) { return fail("prim: not part of the reference build"); }
int hgb_prim_sort_pairs(hgb_scene*, void*, void*, int, int) { return fail("prim: not part of the reference build"); }
#else
int hgb_prim_exclusive_scan(hgb_scene* s, const void* dev_in, int n, int elem_bytes, void* dev_out) {
    if (!bind(s)) return -1;
    if (n < 0 || !dev_out || (n > 0 && !dev_in) || (elem_bytes != 4 && elem_bytes != 8)) return fail("prim_exclusive_scan: bad argument");
    prim_exclusive_scan(s->mem, dev_in, n, elem_bytes, dev_out);
    return 0;
}

int hgb_prim_reduce(hgb_scene* s, const void* dev_in, int n, int op, void* dev_out) {
    if (!bind(s)) return -1;
    if (n < 0 || !dev_out || (n > 0 && !dev_in) || op < 0 || op > 3) return fail("prim_reduce: bad argument");
    prim_reduce(s->mem, dev_in, n, op, dev_out);
    return 0;
}

int hgb_prim_partition(hgb_scene* s, const void* dev_in, const void* dev_flags, int n, void* dev_out) {
    if (!bind(s)) return -1;
    if (n < 0 || (n > 0 && (!dev_in || !dev_flags || !dev_out))) return fail("prim_partition: bad argument");
    return prim_partition(s->mem, static_cast<const int*>(dev_in), static_cast<const int*>(dev_flags), n, static_cast<int*>(dev_out));
}

int hgb_prim_sort_pairs(hgb_scene* s, void* dev_keys, void* dev_vals, int n, int bits) {
    if (!bind(s)) return -1;
    if (n < 0 || bits < 0 || bits > 31 || (n > 0 && (!dev_keys || !dev_vals))) return fail("prim_sort_pairs: bad argument");
    prim_sort_pairs(s->mem, static_cast<int*>(dev_keys), static_cast<int*>(dev_vals), n, bits);
    return 0;
}
#endif

unsigned long long hgb_kernel_launches(void) {
#ifdef HGB_REFERENCE_BUILD
    return 0;
#else
    return kernel_launch_count();
#endif
}

int hgb_device_synchronize(void) {
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : fail("cudaDeviceSynchronize failed");
}

} // extern "C"
