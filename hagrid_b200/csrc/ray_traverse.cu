// Traversal of the irregular grid for sm_100a: setup_traversal / traverse_grid
// (API of src/traverse.h:11-14; semantics of src/traverse.cu:28-95).
//
// Per-ray semantics are exactly the reference's: the same cells are visited in
// the same order, the references of a cell are tested in array order with the
// shrinking tmax, and the ray stops when `hit.t <= texit` or the voxel leaves
// the grid. What is different is how rays are mapped onto the machine:
//
//   * coherent buffers (a W x H raster of camera rays, recognised on the device or on the host): a warp
//     traces an 8x4 pixel tile instead of a 32x1 strip, and resident warps pull tiles from a global
//     counter (kernel A2) so that no warp slot idles behind the slowest warp of its block;
//   * incoherent buffers: persistent warps whose lanes are refilled from a ray counter as their rays
//     end; each iteration the warp serves the larger of two groups of lanes, those that need a cell
//     step and those that own references to test (kernel B, `__ballot_sync` population counts);
//   * host-buffer frames are cut into chunks whose upload, traversal and download overlap.
//
// ncu (profiles/): the coherent kernel issues at 70 % of the SM's peak rate (1 756 warp instructions per 32
// rays at 22 of 32 lanes active) and is bound by instruction issue and dependent-load latency, the incoherent
// one (59 %) by the latency of its dependent L2 loads under scattered access. The grid, the references and
// the triangles a view touches live in the 126 MB L2; compulsory HBM traffic is 32 B/ray in + 16 B/hit out.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "device_math.cuh"
#include "primitives.cuh"
#include "runtime.h"
#include "traverse.h"

// Build-time tuning knobs of the tile kernel (tools/gpu_tile_variants.py builds the alternatives)
#ifndef HGB_TILE_BLOCKS
#define HGB_TILE_BLOCKS 12
#endif
#ifndef HGB_REF_UNROLL
#define HGB_REF_UNROLL 1                 // 0: ptxas decides (it unrolls four-fold)
#endif
#define HGB_STR2(x) #x
#define HGB_STR(x) HGB_STR2(x)
#if HGB_REF_UNROLL > 0
#define HGB_REF_LOOP_PRAGMA _Pragma(HGB_STR(unroll HGB_REF_UNROLL))
#else
#define HGB_REF_LOOP_PRAGMA
#endif

namespace hagrid {

namespace {

/// Grid constants captured by setup_traversal (src/traverse.cu:97-108). All
/// floating-point values are computed on the host in IEEE arithmetic.
struct TraversalParams {
    int   dims_x, dims_y, dims_z;       // virtual dims = top dims << shift
    int   top_x, top_y;                 // top-level dims (x, y)
    int   shift;
    float min_x, min_y, min_z;          // grid box
    float max_x, max_y, max_z;
    float cell_x, cell_y, cell_z;       // extents / virtual dims
    float inv_x, inv_y, inv_z;          // virtual dims / extents
};

/// The constants are a pure function of the host-side Grid (src/traverse.cu:97-101 computes them once per
/// setup_traversal and keeps them in __constant__ memory, one set per process). Here every launch derives them
/// from the grid it is given and passes them as a __grid_constant__ kernel parameter: no process-wide state,
/// any number of scenes, devices and host threads side by side.
TraversalParams params_of(const Grid& grid) {
    // Host IEEE arithmetic, same expressions as src/traverse.cu:97-101.
    const vec3 extents = grid.bbox.extents();
    const ivec3 dims = grid.dims << grid.shift;
    const vec3 inv = vec3(dims) / extents;
    const vec3 cell = extents / vec3(dims);
    TraversalParams P;
    P.dims_x = dims.x; P.dims_y = dims.y; P.dims_z = dims.z;
    P.top_x = dims.x >> grid.shift; P.top_y = dims.y >> grid.shift;
    P.shift = grid.shift;
    P.min_x = grid.bbox.min.x; P.min_y = grid.bbox.min.y; P.min_z = grid.bbox.min.z;
    P.max_x = grid.bbox.max.x; P.max_y = grid.bbox.max.y; P.max_z = grid.bbox.max.z;
    P.cell_x = cell.x; P.cell_y = cell.y; P.cell_z = cell.z;
    P.inv_x = inv.x; P.inv_y = inv.y; P.inv_z = inv.z;
    return P;
}

struct RayState {
    float ox, oy, oz, tmin;
    float dx, dy, dz;
    float ix, iy, iz;                   // safe_rcp(dir)
    float hit_t;
    int   hit_id;
    int   steps;
    int   vx, vy, vz;                   // current voxel
};

/// Ray/triangle test, expression shapes as in the reference SASS:
///   c   = v0 - org                                    FADD x3
///   r   = dir x c      r.x = fma(dy, cz, -(dz*cy)) .. FMUL+FFMA x3
///   det = fma(nz, dz, fma(nx, dx, ny*dy))             FMUL+FFMA x2
///   u,v = prodsign(dot(r, e2 | e1), det)  w = (|det| - u) - v
///   t   = prodsign(dot(n, c), det)
///   accept: u,v,w >= -1e-9,  t >= |det|*tmin,  |det|*tmax > t ; hit.t = t * rcp(|det|)
/// (src/prims.h:266-295)
__device__ __forceinline__ void intersect_tri(RayState& r, const Tri* __restrict__ tris, int ref) {
    using namespace dev;
    const float4 t0 = ldg4(reinterpret_cast<const float4*>(tris + ref) + 0);   // v0, nx
    const float4 t1 = ldg4(reinterpret_cast<const float4*>(tris + ref) + 1);   // e1, ny
    const float4 t2 = ldg4(reinterpret_cast<const float4*>(tris + ref) + 2);   // e2, nz
    const float cx = sub(t0.x, r.ox), cy = sub(t0.y, r.oy), cz = sub(t0.z, r.oz);
    const float rx = diff_of_products(r.dy, cz, r.dz, cy);
    const float ry = diff_of_products(r.dz, cx, r.dx, cz);
    const float rz = diff_of_products(r.dx, cy, r.dy, cx);
    const float det = dot3(t0.w, t1.w, t2.w, r.dx, r.dy, r.dz);
    const float abs_det = fabsf(det);
    const float u = prodsign(dot3(rx, ry, rz, t2.x, t2.y, t2.z), det);
    const float v = prodsign(dot3(rx, ry, rz, t1.x, t1.y, t1.z), det);
    const float w = sub(sub(abs_det, u), v);
    const float eps = 1e-9f;
    if (u >= -eps && v >= -eps && w >= -eps) {
        const float t = prodsign(dot3(t0.w, t1.w, t2.w, cx, cy, cz), det);
        if (t >= mul(abs_det, r.tmin) && mul(abs_det, r.hit_t) > t) {
            r.hit_t = mul(t, rcp(abs_det));
            r.hit_id = ref;
        }
    }
}

/// Loads ray `id`, clips it against the grid box and finds its first voxel
/// (src/traverse.cu:36-54). Returns false when the ray misses the grid.
__device__ __forceinline__ bool init_ray(RayState& r, const TraversalParams& P, const float4 a, const float4 b) {
    using namespace dev;
    r.ox = a.x; r.oy = a.y; r.oz = a.z; r.tmin = a.w;
    r.dx = b.x; r.dy = b.y; r.dz = b.z;
    r.ix = safe_rcp(b.x); r.iy = safe_rcp(b.y); r.iz = safe_rcp(b.z);
    r.hit_t = b.w;
    r.hit_id = -1;
    r.steps = 0;

    const float lx = mul(sub(P.min_x, r.ox), r.ix), hx = mul(sub(P.max_x, r.ox), r.ix);
    const float ly = mul(sub(P.min_y, r.oy), r.iy), hy = mul(sub(P.max_y, r.oy), r.iy);
    const float lz = mul(sub(P.min_z, r.oz), r.iz), hz = mul(sub(P.max_z, r.oz), r.iz);
    const float t0 = fmaxf(sel_min(lx, hx), fmaxf(sel_min(ly, hy), sel_min(lz, hz)));
    const float t1 = fminf(sel_max(lx, hx), fminf(sel_max(ly, hy), sel_max(lz, hz)));
    const float tstart = fmaxf(t0, r.tmin);
    const float tend = fminf(t1, b.w);
    if (tstart > tend) return false;

    // voxel = clamp(int((t * dir + org - grid_min) * grid_inv), 0, dims - 1):  FFMA, FADD, FMUL, F2I
    r.vx = min(P.dims_x - 1, max(0, trunc_to_int(mul(sub(fma(r.dx, tstart, r.ox), P.min_x), P.inv_x))));
    r.vy = min(P.dims_y - 1, max(0, trunc_to_int(mul(sub(fma(r.dy, tstart, r.oy), P.min_y), P.inv_y))));
    r.vz = min(P.dims_z - 1, max(0, trunc_to_int(mul(sub(fma(r.dz, tstart, r.oz), P.min_z), P.inv_z))));
    return true;
}

__device__ __forceinline__ bool start_ray(RayState& r, const TraversalParams& P, const Ray* __restrict__ rays, int id) {
    const float4 a = dev::ldg4_stream(reinterpret_cast<const float4*>(rays + id) + 0);    // org, tmin
    const float4 b = dev::ldg4_stream(reinterpret_cast<const float4*>(rays + id) + 1);    // dir, tmax
    return init_ray(r, P, a, b);
}

/// Enters the cell owning the current voxel, computes where the ray leaves it
/// and moves `voxel` to the next cell (src/traverse.cu:57-77). Returns the
/// exit distance; `cell` receives the reference range.
/// kOct >= 0: the signs of the direction are known at compile time (bit 0: dx >= 0, bit 1: dy >= 0, bit 2: dz >= 0),
/// which folds the far-plane selection, the exit-plane step and the monotone clamp (16 of ~100 instructions).
template <typename CellT, int kOct = -1>
__device__ __forceinline__ float enter_cell(RayState& r, const TraversalParams& P,
                                            const uint32_t* __restrict__ entries,
                                            const CellT* __restrict__ cells, dev::CellBox& cell) {
    using namespace dev;
    const int cell_id = lookup_cell(entries, P.shift, P.top_x, P.top_y, r.vx, r.vy, r.vz);
    cell = load_cell_box(cells, cell_id);

    const bool px = kOct >= 0 ? (kOct & 1) != 0 : r.dx >= 0.0f;
    const bool py = kOct >= 0 ? (kOct & 2) != 0 : r.dy >= 0.0f;
    const bool pz = kOct >= 0 ? (kOct & 4) != 0 : r.dz >= 0.0f;
    const int cx = px ? cell.max_x : cell.min_x;
    const int cy = py ? cell.max_y : cell.min_y;
    const int cz = pz ? cell.max_z : cell.min_z;
    // tcell = (cell_point * cell_size + grid_min - org) * inv_dir:  I2F, FFMA, FADD, FMUL
    const float tx = mul(sub(fma(int_to_float(cx), P.cell_x, P.min_x), r.ox), r.ix);
    const float ty = mul(sub(fma(int_to_float(cy), P.cell_y, P.min_y), r.oy), r.iy);
    const float tz = mul(sub(fma(int_to_float(cz), P.cell_z, P.min_z), r.oz), r.iz);
    const float texit = fminf(tx, fminf(ty, tz));

    // On the exit axis step across the exact integer plane, elsewhere re-derive the voxel from texit
    const int ex = trunc_to_int(mul(sub(fma(r.dx, texit, r.ox), P.min_x), P.inv_x));
    const int ey = trunc_to_int(mul(sub(fma(r.dy, texit, r.oy), P.min_y), P.inv_y));
    const int ez = trunc_to_int(mul(sub(fma(r.dz, texit, r.oz), P.min_z), P.inv_z));
    const int nx = texit == tx ? cx + (px ? 0 : -1) : ex;
    const int ny = texit == ty ? cy + (py ? 0 : -1) : ey;
    const int nz = texit == tz ? cz + (pz ? 0 : -1) : ez;
    r.vx = px ? max(nx, r.vx) : min(nx, r.vx);
    r.vy = py ? max(ny, r.vy) : min(ny, r.vy);
    r.vz = pz ? max(nz, r.vz) : min(nz, r.vz);
    return texit;
}

__device__ __forceinline__ bool outside(const RayState& r, const TraversalParams& P) {
    return (r.vx < 0) | (r.vx >= P.dims_x) | (r.vy < 0) | (r.vy >= P.dims_y) | (r.vz < 0) | (r.vz >= P.dims_z);
}

template <bool kPrimId>
__device__ __forceinline__ void finish_ray(const RayState& r, Hit* __restrict__ hits, int id) {
    // u = v = 0: the reference never defines COMPUTE_UVS (src/prims.h:285-288)
    dev::stg4_stream(hits + id, make_float4(__int_as_float(kPrimId ? r.hit_id : r.steps), r.hit_t, 0.0f, 0.0f));
}

// ---------------------------------------------------------------------------
// Ray packet layout. Camera rays arrive in raster order (src/main.cpp:52-66):
// a warp of 32 consecutive rays is a 32x1 pixel strip whose rays fan out over
// many cells. When the buffer is recognised as a W x H raster (detect_raster
// below) thread -> ray assignment is re-tiled so that a warp traces an 8x4 pixel
// tile: the lanes then walk (nearly) the same cells and test the same reference
// lists, which is what the SIMT efficiency of this kernel depends on. The mapping
// is a bijection on [0, num_rays) for any W, so it can only affect speed.
// ---------------------------------------------------------------------------
constexpr int kTileW = 8, kTileH = 4;

__device__ __forceinline__ int tiled_ray_index(int thread, int width) {
    const int tile = thread >> 5, lane = thread & 31;
    const int tiles_per_row = width / kTileW;
    const int ty = tile / tiles_per_row, tx = tile - ty * tiles_per_row;
    return (ty * kTileH + (lane >> 3)) * width + tx * kTileW + (lane & 7);
}

/// One block. layout[0] <- W if rays[0..n) look like a W x (n / W) raster of a
/// smoothly varying (org, dir) field with W % 8 == 0 and (n / W) % 4 == 0, else 0.
/// The row length is the first index where the ray-to-ray delta breaks.
__global__ void __launch_bounds__(256) detect_raster(const Ray* __restrict__ rays, int n, int* __restrict__ layout,
                                                     volatile int* __restrict__ layout_host) {
    __shared__ int row_break;
    __shared__ int bad, misses;
    constexpr int kScan = 16384, kSpot = 1024;
    if (threadIdx.x == 0) { row_break = kScan; bad = 0; misses = 0; }
    __syncthreads();
    int width = 0;
    if (n >= 4 * 64) {
        const float4 a0 = dev::ldg4(reinterpret_cast<const float4*>(rays) + 0), b0 = dev::ldg4(reinterpret_cast<const float4*>(rays) + 1);
        const float4 a1 = dev::ldg4(reinterpret_cast<const float4*>(rays) + 2), b1 = dev::ldg4(reinterpret_cast<const float4*>(rays) + 3);
        const float step[6] = {a1.x - a0.x, a1.y - a0.y, a1.z - a0.z, b1.x - b0.x, b1.y - b0.y, b1.z - b0.z};
        float tol = 0.0f;
        for (int k = 0; k < 6; k++) tol = fmaxf(tol, fabsf(step[k]));
        tol *= 0.25f;
        const int limit = min(n, kScan);
        auto continues = [&](int i) {
            const float4 a = dev::ldg4(reinterpret_cast<const float4*>(rays + i)), b = dev::ldg4(reinterpret_cast<const float4*>(rays + i) + 1);
            const float4 pa = dev::ldg4(reinterpret_cast<const float4*>(rays + i - 1)), pb = dev::ldg4(reinterpret_cast<const float4*>(rays + i - 1) + 1);
            const float d[6] = {a.x - pa.x, a.y - pa.y, a.z - pa.z, b.x - pb.x, b.y - pb.y, b.z - pb.z};
            float err = 0.0f;
            for (int k = 0; k < 6; k++) err = fmaxf(err, fabsf(d[k] - step[k]));
            return err <= tol;                       // NaN compares false
        };
        if (tol > 0.0f) {
            // 256 candidates per round; the block stops at the first round that holds a break
            for (int base = 2; base < limit; base += 256) {
                const int i = base + threadIdx.x;
                if (i < limit && !continues(i)) atomicMin(&row_break, i);
                __syncthreads();
                const bool found = row_break < kScan;
                __syncthreads();                      // everyone has read the flag before the next round writes it
                if (found) break;
            }
        }
        __syncthreads();
        width = row_break;
        const bool shape_ok = tol > 0.0f && width < kScan && width >= 64 && width % kTileW == 0 && n % width == 0 &&
                              (n / width) % kTileH == 0;
        if (shape_ok) {
            // the second and the last row must continue with the same step ...
            for (int k = threadIdx.x; k < 62; k += 256) {
                if (!continues(width + 1 + k)) atomicOr(&bad, 1);
                if (!continues(n - width + 1 + k)) atomicOr(&bad, 1);
            }
            // ... and so must the buffer as a whole: kSpot positions spread over it (row starts skipped). A frame of
            // secondary rays often begins and ends with rows of re-emitted camera rays; its interior does not pass.
            for (int k = threadIdx.x; k < kSpot; k += 256) {
                int i = int((long long)n * k / kSpot) + 1 + (k * 37) % 61;
                if (i >= n) i = n - 1;
                if (i % width != 0 && !continues(i)) atomicAdd(&misses, 1);
            }
        }
        __syncthreads();
        if (!shape_ok || bad || misses > kSpot / 16) width = 0;
    }
    if (threadIdx.x == 0) {
        layout[0] = width;
        *layout_host = width;
    }
}

// ---------------------------------------------------------------------------
// The march of one ray (shared by all kernels) and the tile index of the resident-warp kernels.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int tiled_ray_index_nodiv(int tile, int lane, int width) {
    // tile / (width / 8) by a float estimate and one fix-up step each way (tile < 2^24 keeps it within +-1)
    const int per_row = width >> 3;
    int ty = __float2int_rz(__int2float_rn(tile) * dev::rcp(__int2float_rn(per_row)));
    int tx = tile - ty * per_row;
    if (tx < 0) { ty--; tx += per_row; }
    if (tx >= per_row) { ty++; tx -= per_row; }
    return (ty * kTileH + (lane >> 3)) * width + tx * kTileW + (lane & 7);
}

/// One cell of the march (src/traverse.cu:57-88): enter it, test its references in array order, count the steps.
/// Returns the exit distance; the ray's voxel already is the next cell's.
template <typename CellT, int kOct = -1>
__device__ __forceinline__ float visit_cell(RayState& r, const TraversalParams& P, const uint32_t* __restrict__ entries,
                                            const CellT* __restrict__ cells, const int* __restrict__ ref_ids,
                                            const Tri* __restrict__ tris) {
    constexpr bool kSentinel = sizeof(CellT) == sizeof(SmallCell);
    dev::CellBox cell;
    const float texit = enter_cell<CellT, kOct>(r, P, entries, cells, cell);
    // plain loops (HGB_REF_LOOP_PRAGMA: not unrolled, which keeps the tile kernel at 40 registers = 48 warps per SM)
    if (kSentinel) {
        int cur = cell.begin;
        if (cur >= 0)
            for (int ref = __ldg(ref_ids + cur++); ref >= 0; ref = __ldg(ref_ids + cur++)) intersect_tri(r, tris, ref);
        r.steps += 1 + (cur - cell.begin);
    } else {
        HGB_REF_LOOP_PRAGMA
        for (int cur = cell.begin; cur < cell.end; cur++) intersect_tri(r, tris, __ldg(ref_ids + cur));
        r.steps += 1 + (cell.end - cell.begin);
    }
    return texit;
}

__device__ __forceinline__ bool left_grid(const RayState& r, const TraversalParams& P) {
    // unsigned compares fold the < 0 tests
    return (unsigned(r.vx) >= unsigned(P.dims_x)) | (unsigned(r.vy) >= unsigned(P.dims_y)) | (unsigned(r.vz) >= unsigned(P.dims_z));
}

/// The march of one ray through the grid after init_ray (src/traverse.cu:56-90)
template <typename CellT, int kOct = -1>
__device__ __forceinline__ void walk(RayState& r, const TraversalParams& P, const uint32_t* __restrict__ entries,
                                     const CellT* __restrict__ cells, const int* __restrict__ ref_ids,
                                     const Tri* __restrict__ tris) {
    while (true) {
        const float texit = visit_cell<CellT, kOct>(r, P, entries, cells, ref_ids, tris);
        if (r.hit_t <= texit) break;
        if (left_grid(r, P)) break;
    }
}

template <typename CellT, bool kPrimId>
__device__ __forceinline__ void trace_one(const TraversalParams& P, const uint32_t* __restrict__ entries,
                                          const CellT* __restrict__ cells, const int* __restrict__ ref_ids,
                                          const Tri* __restrict__ tris, const Ray* __restrict__ rays,
                                          Hit* __restrict__ hits, int id) {
    RayState r;
    if (start_ray(r, P, rays, id)) walk(r, P, entries, cells, ref_ids, tris);
    finish_ray<kPrimId>(r, hits, id);
}

// ---------------------------------------------------------------------------
// Kernel A: one thread per ray, optionally re-tiled. The plain mapping of the reference (variant 0) and
// its tiled form (variant 2), kept as the baselines the resident-warp kernels are measured against.
// ---------------------------------------------------------------------------
template <typename CellT, bool kPrimId>
__global__ void __launch_bounds__(128)
traverse_per_thread(const __grid_constant__ TraversalParams P,
                    const uint32_t* __restrict__ entries, const CellT* __restrict__ cells,
                    const int* __restrict__ ref_ids, const Tri* __restrict__ tris,
                    const Ray* __restrict__ rays, Hit* __restrict__ hits, int num_rays,
                    const int* __restrict__ layout, int host_width) {
    int id = threadIdx.x + blockDim.x * blockIdx.x;
    if (id >= num_rays) return;
    {   // raster width: found on the device (layout word) or already known to the host
        const int width = layout ? __ldg(layout) : host_width;
        if (width > 0) id = tiled_ray_index(id, width);
    }
    trace_one<CellT, kPrimId>(P, entries, cells, ref_ids, tris, rays, hits, id);
}

/// Direction octant of a ray (bit k set: component k >= 0, the test enter_cell makes)
__device__ __forceinline__ int octant_of(const RayState& r) {
    return (r.dx >= 0.0f ? 1 : 0) | (r.dy >= 0.0f ? 2 : 0) | (r.dz >= 0.0f ? 4 : 0);
}

/// March of the lanes with `ok` set. Called by all 32 lanes of a converged warp: when every marching lane has
/// the same direction octant — the rule for an 8x4 tile of camera rays — the warp takes the loop specialised
/// for it, otherwise the generic one. Same cells, same triangles, same order either way.
template <typename CellT>
__device__ __forceinline__ bool walk_warp(bool ok, RayState& r, const TraversalParams& P, const uint32_t* __restrict__ entries,
                                          const CellT* __restrict__ cells, const int* __restrict__ ref_ids,
                                          const Tri* __restrict__ tris) {
    constexpr unsigned kAll = 0xFFFFFFFFu;
    const unsigned marching = __ballot_sync(kAll, ok);
    if (marching == 0) return true;
    const int oct = octant_of(r);
    const int first = __shfl_sync(kAll, oct, __ffs(marching) - 1);
    const bool uniform = __all_sync(kAll, !ok || oct == first);
    if (!ok) return uniform;
    if (uniform) {
        switch (first) {
            case 0: walk<CellT, 0>(r, P, entries, cells, ref_ids, tris); break;
            case 1: walk<CellT, 1>(r, P, entries, cells, ref_ids, tris); break;
            case 2: walk<CellT, 2>(r, P, entries, cells, ref_ids, tris); break;
            case 3: walk<CellT, 3>(r, P, entries, cells, ref_ids, tris); break;
            case 4: walk<CellT, 4>(r, P, entries, cells, ref_ids, tris); break;
            case 5: walk<CellT, 5>(r, P, entries, cells, ref_ids, tris); break;
            case 6: walk<CellT, 6>(r, P, entries, cells, ref_ids, tris); break;
            default: walk<CellT, 7>(r, P, entries, cells, ref_ids, tris); break;
        }
    } else {
        walk<CellT, -1>(r, P, entries, cells, ref_ids, tris);
    }
    return uniform;
}

// ---------------------------------------------------------------------------
// Kernel A2: resident warps pull 32-ray tiles from a global counter (coherent buffers). Same per-ray
// march as kernel A; what changes is residency — a block of kernel A keeps its four warp slots until its
// slowest warp is done, here a warp that is done takes the next tile — and the march itself, which the
// warp picks specialised for its direction octant when all its rays share one (walk_warp).
// ---------------------------------------------------------------------------
constexpr int kTileBlock = 128;
constexpr int kTileBlocksPerSm = HGB_TILE_BLOCKS;     // 12 blocks = 48 warps per SM at 40 registers (reference loop not unrolled): measured best, profiles/r02_traverse_experiments.md

/// Tile hand-out without a reset: the counter only ever grows, the host passes the value it has at the start
/// of the launch (`base`). Every traced tile is followed by exactly one fetch, so a launch over T tiles advances
/// the counter by exactly T and the host knows the next base without reading anything back (uint32 wrap-around
/// is harmless: only differences are used).
__device__ __forceinline__ int fetch_tile(unsigned* __restrict__ next_tile, unsigned base, int first_dynamic, int lane) {
    int tile = 0;
    if (lane == 0) tile = first_dynamic + int(atomicAdd(next_tile, 1u) - base);
    return __shfl_sync(0xFFFFFFFFu, tile, 0);
}


#ifdef HGB_TILE_TRACE
// Diagnosis build only (tools/gpu_tile_variants.py): per warp [first tile started, last tile finished, tiles traced], ns
__device__ long long g_tile_trace[3 * 8192];
__device__ __forceinline__ long long global_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

/// Longest tiles first, the very longest in parts. A launch ends with a tail: the queue is empty, every warp finishes
/// the tile it holds, and the launch lasts until the slowest of them is done (C2: queue dry at 139 us, last warp back
/// at 160-165 us; C5: dry at 65 us, last warp at 210 us -- tiles of foliage whose rays diverge and take turns). Callers
/// trace the same buffer again and again (a viewer's frame loop, a benchmark's iterations) and a frame resembles the
/// one before, so a launch can time its tiles (`cost`: clock ticks / 64, one 2-byte store per tile) and later launches
/// on the same buffer are handed a ticket list made from those times on a side stream (order_tiles): tiles by cost
/// class, longest first, so that the tiles in flight when the queue runs dry are the cheap ones; and the few tiles
/// that alone last a good part of the whole launch as 2^split_log tickets each, every ticket tracing 32 >> split_log
/// of the tile's rays -- where the lanes of a warp take turns, fewer lanes per warp make a shorter chain. A ticket is
/// tile | (part + 1) << 26 (part bits 0: the whole tile; at most 32 parts, so never all ones) or -1 (nothing: the list
/// has a fixed length the host knows).
/// Every ray is traced exactly once whatever the list says, each in its own lane with its own state: the list cannot
/// change a hit.
struct TileHistory {
    const int* tickets;          // null: tile k is ticket k
    unsigned short* cost;        // written by the kTimed instantiation only (such a launch gets a list without parts)
    int extra_tickets;           // length of the list - number of tiles
    int split_log;
};

template <typename CellT, bool kPrimId, bool kTimed = false>
__global__ void __launch_bounds__(kTileBlock, kTileBlocksPerSm)
traverse_tiles(const __grid_constant__ TraversalParams P,
               const uint32_t* __restrict__ entries, const CellT* __restrict__ cells,
               const int* __restrict__ ref_ids, const Tri* __restrict__ tris,
               const Ray* __restrict__ rays, Hit* __restrict__ hits, int num_rays,
               const int* __restrict__ layout, int host_width, unsigned* __restrict__ next_tile, unsigned ticket_base,
               int* __restrict__ feedback, const __grid_constant__ TileHistory history) {
    const int lane = threadIdx.x & 31;
    // feedback (device memory, may be null): [0] += warps whose rays did not share a direction octant, [1] += 1
    // per launch. A buffer of camera rays has a few such warps along the image axes; a buffer whose warps are
    // mostly mixed is not what this kernel is for: the host copies the words back now and then and moves such a
    // buffer to the incoherent kernel.
    if (feedback && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(feedback + 1, 1);
    const int width = layout ? __ldg(layout) : host_width;
    const int num_tiles = (num_rays + 31) >> 5;
    const int first_dynamic = gridDim.x * (kTileBlock / 32);
#ifdef HGB_TILE_TRACE
    const int trace_slot = blockIdx.x * (kTileBlock / 32) + (threadIdx.x >> 5);
    int traced_tiles = 0;
    if (lane == 0 && trace_slot < 8192) g_tile_trace[3 * trace_slot] = global_ns();
#endif
    const int num_tickets = num_tiles + history.extra_tickets;
    int ticket = blockIdx.x * (kTileBlock / 32) + (threadIdx.x >> 5);
    while (ticket < num_tickets) {
        const int what = history.tickets ? __ldg(history.tickets + ticket) : ticket;
        if (what == -1) { ticket = fetch_tile(next_tile, ticket_base, first_dynamic, lane); continue; }
        const unsigned started = kTimed ? unsigned(clock()) : 0u;
        RayState r;
        bool ok = false;
        {
            const int tile = what & ((1 << 26) - 1), part = int(unsigned(what) >> 26);
            int id = tile * 32 + lane;
            if (id < num_rays && (part == 0 || (lane >> (5 - history.split_log)) == part - 1)) {
                if (width > 0) id = tiled_ray_index_nodiv(tile, lane, width);
                ok = start_ray(r, P, rays, id);
            }
        }
        const bool uniform = walk_warp(ok, r, P, entries, cells, ref_ids, tris);
        if (!uniform && feedback && lane == 0) atomicAdd(feedback, 1);
        {   // the ray's place in the buffer again (cheaper than keeping it in a register across the march)
            const int tile = what & ((1 << 26) - 1), part = int(unsigned(what) >> 26);
            int id = tile * 32 + lane;
            if (id < num_rays && (part == 0 || (lane >> (5 - history.split_log)) == part - 1)) {
                if (width > 0) id = tiled_ray_index_nodiv(tile, lane, width);
                finish_ray<kPrimId>(r, hits, id);
            }
        }
        __syncwarp();
        if (kTimed && lane == 0) history.cost[what & ((1 << 26) - 1)] = (unsigned short)min((unsigned(clock()) - started) >> 6, 0xFFFFu);
#ifdef HGB_TILE_TRACE
        traced_tiles++;
        if (lane == 0 && trace_slot < 8192) { g_tile_trace[3 * trace_slot + 1] = global_ns(); g_tile_trace[3 * trace_slot + 2] = traced_tiles; }
#endif
        ticket = fetch_tile(next_tile, ticket_base, first_dynamic, lane);
    }
}

/// stats[0] = the largest cost, stats[1] = the sum of all costs >> 8 (one block)
__global__ void __launch_bounds__(1024) tile_cost_stats(const unsigned short* __restrict__ cost, int num_tiles, unsigned* __restrict__ stats) {
    __shared__ unsigned warp_max[32];
    __shared__ unsigned long long warp_sum[32];
    unsigned m = 0;
    unsigned long long sum = 0;
    for (int t = threadIdx.x; t < num_tiles; t += 1024) { const unsigned c = cost[t]; m = max(m, c); sum += c; }
    for (int d = 16; d > 0; d >>= 1) { m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, d)); sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d); }
    if ((threadIdx.x & 31) == 0) { warp_max[threadIdx.x >> 5] = m; warp_sum[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x < 32) {
        m = warp_max[threadIdx.x]; sum = warp_sum[threadIdx.x];
        for (int d = 16; d > 0; d >>= 1) { m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, d)); sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d); }
        if (threadIdx.x == 0) { stats[0] = m; stats[1] = unsigned(sum >> 8); }
    }
}

/// keys[t] = the tile's cost class out of 2^bits equal classes between 0 and the largest cost, most expensive class
/// first (the sort is stable: tiles of one class keep their buffer order and with it the cells they share with their
/// neighbours); order[t] = t
__global__ void __launch_bounds__(256) tile_order_keys(const unsigned short* __restrict__ cost, int num_tiles, int bits, const unsigned* __restrict__ stats,
                                                       int* __restrict__ keys, int* __restrict__ order) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= num_tiles) return;
    const unsigned top = __ldg(stats);
    keys[t] = int(((top - unsigned(cost[t])) << bits) / (top + 1u));
    order[t] = t;
}

/// The ticket list (TileHistory) from the sorted tiles: the first `split` of them get 2^split_log tickets each -- the
/// parts of the tile if it alone costs `share` percent or more of what one of the launch's `warps` warps has to do
/// (sum of all costs / warps), otherwise the whole tile and empty tickets -- the others one ticket each.
__global__ void __launch_bounds__(256) tile_tickets(const int* __restrict__ sorted, const unsigned short* __restrict__ cost, const unsigned* __restrict__ stats,
                                                    int num_tiles, int split, int split_log, int share, int warps, int* __restrict__ tickets) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int parts = 1 << split_log;
    if (j >= num_tiles + split * (parts - 1)) return;
    if (j >= split * parts) { tickets[j] = sorted[j - split * (parts - 1)]; return; }
    const int tile = sorted[j >> split_log], part = j & (parts - 1);
    // cost >= share % of sum / warps, the sum being kept >> 8
    const bool in_parts = (unsigned long long)cost[tile] * unsigned(warps) * 100ull >= ((unsigned long long)stats[1] << 8) * unsigned(share);
    tickets[j] = in_parts ? tile | ((part + 1) << 26) : part == 0 ? tile : -1;
}

// ---------------------------------------------------------------------------
// Frames from a camera (the interactive loop of src/main.cpp:591-625 without its host round trips): the
// reference generates the primary rays on one CPU thread (gen_rays, src/main.cpp:52-66), uploads 32 B per
// ray, traces, downloads 16 B per hit and colours the pixels on the CPU (update_surface,
// src/main.cpp:90-111). Here a ray is generated in the registers of the thread that traces it and its
// pixel is written by the same thread: no ray buffer, no hit buffer, 4 B per pixel leave the device.
// The arithmetic is the host's: IEEE multiply / add / divide, no contraction (the front end is compiled
// by g++ for x86-64 without FMA), so rays and pixels are bit-identical to the reference's.
// ---------------------------------------------------------------------------
struct FrameParams {
    float eye_x, eye_y, eye_z;
    float dir_x, dir_y, dir_z;
    float right_x, right_y, right_z;
    float up_x, up_y, up_z;
    float clip;
    int   width, height;
};

/// gen_rays for pixel (x, y): kx = 2x/w - 1, ky = 1 - 2y/h, dir = cam.dir + cam.right*kx + cam.up*ky
__device__ __forceinline__ void camera_ray(const FrameParams& F, int x, int y, float4& a, float4& b) {
    const float kx = __fsub_rn(__fdiv_rn(__int2float_rn(2 * x), __int2float_rn(F.width)), 1.0f);
    const float ky = __fsub_rn(1.0f, __fdiv_rn(__int2float_rn(2 * y), __int2float_rn(F.height)));
    a = make_float4(F.eye_x, F.eye_y, F.eye_z, 0.0f);
    b.x = __fadd_rn(__fadd_rn(F.dir_x, __fmul_rn(F.right_x, kx)), __fmul_rn(F.up_x, ky));
    b.y = __fadd_rn(__fadd_rn(F.dir_y, __fmul_rn(F.right_y, kx)), __fmul_rn(F.up_y, ky));
    b.z = __fadd_rn(__fadd_rn(F.dir_z, __fmul_rn(F.right_z, kx)), __fmul_rn(F.up_z, ky));
    b.w = F.clip;
}

/// float -> uint8_t the way the x86 front end does it: truncate to int32, keep the low byte
__device__ __forceinline__ unsigned to_byte(float v) { return unsigned(__float2int_rz(v)) & 0xFFu; }

/// update_surface (src/main.cpp:90-111): BGRA of one pixel; mode 0 = depth, 1 = steps as grey, 2 = heat map
template <int kMode>
__device__ __forceinline__ unsigned shade_pixel(const RayState& r, float clip) {
    unsigned b, g, red;
    if (kMode == 0) {
        b = g = red = to_byte(__fdiv_rn(__fmul_rn(255.0f, r.hit_t), clip));
    } else if (kMode == 1) {
        b = g = red = unsigned(min(255, r.steps));
    } else {
        // gradient(), src/main.cpp:68-88
        const float gx[5] = {0.0f, 0.0f, 0.0f, 255.0f, 255.0f};
        const float gy[5] = {0.0f, 255.0f, 128.0f, 255.0f, 0.0f};
        const float gz[5] = {255.0f, 255.0f, 0.0f, 0.0f, 0.0f};
        const float s = 1.0f / 5;
        const float k = __fdiv_rn(__int2float_rn(min(100, r.steps)), 100.0f);
        const int i = min(4, __float2int_rz(__fmul_rn(k, 5.0f)));
        const int j = min(4, i + 1);
        const float t = __fdiv_rn(__fsub_rn(k, __fmul_rn(__int2float_rn(i), s)), s);
        const float u = __fsub_rn(1.0f, t);
        red = to_byte(__fadd_rn(__fmul_rn(u, gx[i]), __fmul_rn(t, gx[j])));
        g   = to_byte(__fadd_rn(__fmul_rn(u, gy[i]), __fmul_rn(t, gy[j])));
        b   = to_byte(__fadd_rn(__fmul_rn(u, gz[i]), __fmul_rn(t, gz[j])));
    }
    return b | (g << 8) | (red << 16) | 0xFF000000u;
}

/// Pixel of tile `tile`, lane `lane`: 8x4 tiles when the image allows it, scan-line order otherwise
__device__ __forceinline__ int frame_pixel(const FrameParams& F, int tile, int lane) {
    const bool tiled = (F.width % kTileW == 0) && (F.height % kTileH == 0);
    return tiled ? tiled_ray_index_nodiv(tile, lane, F.width) : tile * 32 + lane;
}

/// Fused frame: generate, trace, shade. One launch, resident warps pulling tiles.
template <typename CellT, int kMode>
__global__ void __launch_bounds__(kTileBlock, kTileBlocksPerSm)
render_tiles(const __grid_constant__ TraversalParams P, const __grid_constant__ FrameParams F,
             const uint32_t* __restrict__ entries, const CellT* __restrict__ cells,
             const int* __restrict__ ref_ids, const Tri* __restrict__ tris,
             unsigned* __restrict__ pixels, unsigned* __restrict__ next_tile, unsigned ticket_base) {
    const int lane = threadIdx.x & 31;
    const int num_pixels = F.width * F.height;
    const int num_tiles = (num_pixels + 31) >> 5;
    const int first_dynamic = gridDim.x * 4;
    int tile = blockIdx.x * 4 + (threadIdx.x >> 5);
    while (tile < num_tiles) {
        const bool live = tile * 32 + lane < num_pixels;
        int id = 0;
        RayState r;
        bool ok = false;
        if (live) {
            id = frame_pixel(F, tile, lane);
            const int y = id / F.width, x = id - y * F.width;
            float4 a, b;
            camera_ray(F, x, y, a, b);
            ok = init_ray(r, P, a, b);
        }
        walk_warp(ok, r, P, entries, cells, ref_ids, tris);
        if (live) pixels[id] = shade_pixel<kMode>(r, F.clip);
        __syncwarp();
        tile = fetch_tile(next_tile, ticket_base, first_dynamic, lane);
    }
}

/// gen_rays alone: the ray buffer the reference's front end would upload
__global__ void __launch_bounds__(256) generate_camera_rays(const __grid_constant__ FrameParams F, Ray* __restrict__ rays) {
    const int id = blockIdx.x * 256 + threadIdx.x;
    if (id >= F.width * F.height) return;
    const int y = id / F.width, x = id - y * F.width;
    float4 a, b;
    camera_ray(F, x, y, a, b);
    reinterpret_cast<float4*>(rays + id)[0] = a;
    reinterpret_cast<float4*>(rays + id)[1] = b;
}

constexpr int kBlockThreads = 128;

// ---------------------------------------------------------------------------
// Kernel B: persistent warps with majority scheduling (incoherent buffers). Every iteration the warp
// counts the lanes that need a cell step and the lanes that own references and serves the larger
// group (offline SIMT model, DESIGN.md section 4.1, policy "count": 0.31 cell + 0.34 triangle iterations per ray against
// 0.86 + 0.29 when the cell phase runs until every lane owns references). With fewer instructions the
// kernel is bound by the latency of its dependent loads, the ray-counter atomic first among them (ncu:
// 13 % of the stall samples), so rays are reserved 32 at a time two blocks ahead: the atomic of a block
// is issued one block before its value is read. Software prefetches of the next voxel-map word and of
// the next triangle were measured and lose (the L1 tag stage is the second bottleneck), and so does a
// per-reference copy of the triangles (64-byte records: one latency less per test, but the duplicates
// of a triangle no longer share cache lines).
// ---------------------------------------------------------------------------
constexpr int kVoteBlock = 32;          // rays reserved per atomic
constexpr int kVoteRefillMinLanes = 4;
constexpr int kVoteTurn = 3;            // references a lane tests per triangle turn
#ifndef HGB_VOTE_BLOCKS
#define HGB_VOTE_BLOCKS 10
#endif
constexpr int kVoteBlocksPerSm = HGB_VOTE_BLOCKS;    // 10: <= 51 registers

template <typename CellT, bool kPrimId>
__global__ void __launch_bounds__(kBlockThreads, kVoteBlocksPerSm)
traverse_voting(const __grid_constant__ TraversalParams P,
                const uint32_t* __restrict__ entries, const CellT* __restrict__ cells,
                const int* __restrict__ ref_ids, const Tri* __restrict__ tris,
                const Ray* __restrict__ rays, Hit* __restrict__ hits, int num_rays,
                int* __restrict__ next_ray, const int* __restrict__ order) {
    // `order` (may be null): the k-th ray handed out is rays[order[k]] (ray binning, see bin_rays); hits go to the
    // ray's own slot, so the caller sees nothing of it
    constexpr bool kSentinel = sizeof(CellT) == sizeof(SmallCell);
    constexpr unsigned kAll = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;

    RayState r;
    int   ray_id = -1;          // -1: lane is idle
    int   ref = -1;             // reference being tested next (-1: none parked)
    int   cur = 0, end = 0;     // rest of the parked reference range
    float texit = 0.0f;

    // ray reservation: [pool, pool_end) is handed out now, `ahead` is the block after it, `requested`
    // (lane 0) is the atomic in flight for the block after that: its value is read one block later
    int pool, pool_end, ahead, requested = 0;
    {
        int first = 0;
        if (lane == 0) first = atomicAdd(next_ray, 2 * kVoteBlock);
        first = __shfl_sync(kAll, first, 0);
        pool = min(first, num_rays); pool_end = min(first + kVoteBlock, num_rays);
        ahead = first + kVoteBlock;
        if (lane == 0) requested = atomicAdd(next_ray, kVoteBlock);
    }

    while (true) {
        // ---- refill idle lanes from the reserved block
        const unsigned idle = __ballot_sync(kAll, ray_id < 0);
        const bool drained = pool >= num_rays;
        if (idle == kAll && drained) break;
        if (!drained && __popc(idle) >= kVoteRefillMinLanes) {
            if (ray_id < 0) {
                int id = pool + __popc(idle & ((1u << lane) - 1u));
                if (id < pool_end) {
                    if (order) id = __ldg(order + id);
                    if (start_ray(r, P, rays, id)) ray_id = id;
                    else finish_ray<kPrimId>(r, hits, id);
                }
            }
            pool = min(pool + __popc(idle), pool_end);
            if (pool == pool_end) {            // block used up: move on to the next one, ask for another
                pool = min(ahead, num_rays); pool_end = min(ahead + kVoteBlock, num_rays);
                ahead = __shfl_sync(kAll, requested, 0);
                if (lane == 0) requested = atomicAdd(next_ray, kVoteBlock);
            }
        }

        const unsigned walkers = __ballot_sync(kAll, ray_id >= 0 && ref < 0);
        const unsigned testers = __ballot_sync(kAll, ref >= 0);
        if (__popc(walkers) >= __popc(testers)) {        // weighted votes (2:1 ... 1:2) measured slower
            if (ray_id >= 0 && ref < 0) {
                dev::CellBox cell;
                texit = enter_cell(r, P, entries, cells, cell);
                cur = cell.begin;
                if (kSentinel) {
                    ref = cur >= 0 ? __ldg(ref_ids + cur++) : -1;
                    r.steps += 1 + (ref >= 0 ? 1 : 0);      // +1 per reference word read, sentinel included
                } else {
                    end = cell.end;
                    ref = cur < end ? __ldg(ref_ids + cur++) : -1;
                    r.steps += 1 + (cell.end - cell.begin);
                }
                if (ref < 0 && (r.hit_t <= texit || outside(r, P))) {
                    finish_ray<kPrimId>(r, hits, ray_id);
                    ray_id = -1;
                }
            }
        } else if (ref >= 0) {
            // up to kVoteTurn references per turn: their triangles' loads are in flight together
            int batch[kVoteTurn];
            batch[0] = ref;
#pragma unroll
            for (int k = 1; k <= kVoteTurn; k++) {
                int next = -1;
                if (batch[k - 1] >= 0) {
                    if (kSentinel) { next = __ldg(ref_ids + cur++); r.steps++; }
                    else           { next = cur < end ? __ldg(ref_ids + cur++) : -1; }
                }
                if (k < kVoteTurn) batch[k] = next; else ref = next;
            }
#pragma unroll
            for (int k = 0; k < kVoteTurn; k++)
                if (batch[k] >= 0) intersect_tri(r, tris, batch[k]);
            if (ref < 0 && (r.hit_t <= texit || outside(r, P))) {
                finish_ray<kPrimId>(r, hits, ray_id);
                ray_id = -1;
            }
        }
    }
}

/// A tile counter in device memory and the value the host knows it to have (fetch_tile)
struct Ticket {
    unsigned* word = nullptr;
    unsigned base = 0;
};

/// Per-device launch state. One host thread at a time per device (`lock`): the tickets and the buffer
/// classification are host-side bookkeeping of what has been enqueued.
struct DeviceState {
    std::mutex lock;
    int* words = nullptr;            // 256 bytes of device memory, one 32-byte sector per word in use (below)
    int* vote_counter = nullptr;     // ray counter of the persistent voting kernel (reset before every launch)
    Ticket tiles;                    // tile counter of traverse_tiles / render_tiles on the default stream
    // What is known about the ray buffers traced lately ((address, count) -> raster width or incoherent). A frame loop
    // alternates between a few buffers (primary rays, second wave, ...): each keeps its own answer, its own look-ahead
    // word and its own feedback counters, so that switching buffers costs no new look and no host synchronisation.
    struct SeenBuffer {
        const void* rays = nullptr;
        int count = -1;
        int cls = -1;                // -1: first look in flight, 0: incoherent, > 0: raster of that width
        int* layout = nullptr;       // device: [0] = raster width found by detect_raster (0 = none)
        int* feedback_dev = nullptr; // device: [0] mixed-octant warps, [1] launches of traverse_tiles (see there)
        int* host = nullptr;         // pinned, mapped: [0] copy of layout[0] (-1 = look in flight), [2..3] copy of feedback_dev
        int  feedback_tick = 0;
        int  feedback_mixed = 0, feedback_launches = 0;   // counter values when the buffer was armed
        bool feedback_armed = false;
        unsigned long long used = 0; // launch number of the last use (least recently used entry is replaced)
        // tile history of a raster buffer (TileHistory): two cost arrays written alternately by the launches, the order
        // made from the older one on the side stream, and the event that says the order is complete
        int* history = nullptr;      // one allocation: the arrays below
        int history_tiles = 0;       // tiles the allocation is good for
        int* tickets[2] = {};        // ticket lists, tiles + kMaxSplitExtra each
        int* sorted = nullptr;       // the tiles by cost class as the last sort left them (the list of a timed launch)
        int* sort_space = nullptr;   // keys and values of the sort, twice, and its scratch
        unsigned* stats = nullptr;
        unsigned short* cost = nullptr;
        int tile_launches = 0;       // launches of traverse_tiles on this buffer
        int history_epoch = 0;
        int pending_since = 0;       // the timed launch the pending order comes from
        int order_ready = -1;        // index of the order array a launch may use (-1: none yet)
        int order_pending = -1;      // index of the order array being made on the side stream (-1: none)
        int extra[2] = {};           // tickets beyond one per tile in each list
        int split_log[2] = {};
        cudaEvent_t timed = nullptr, ordered = nullptr;
    };
    static constexpr int kMaxSplit = 1024, kMaxSplitExtra = kMaxSplit * 31;
    cudaStream_t side = nullptr;     // stream the orders are made on
    static constexpr int kSeenBuffers = 4;
    SeenBuffer seen[kSeenBuffers];
    unsigned long long launches = 0;
    int* seen_host = nullptr;        // pinned block behind seen[].host
    int num_sms = 0;
    int* sort_scratch = nullptr;     // ray binning ("ray_sort"): keys, order and the sort's scratch
    size_t sort_capacity = 0;
    // host-buffer frames (traverse_grid_host): one upload stream, one download stream, two traversal streams
    static constexpr int kStreams = 4, kMaxChunks = 64;
    cudaStream_t streams[kStreams] = {};            // 0 = upload, 1 = download, 2 and 3 = traversal
    cudaEvent_t  uploaded[kMaxChunks] = {};
    cudaEvent_t  traced[kMaxChunks] = {};
    cudaEvent_t  stream_done[kStreams] = {};
    cudaEvent_t  frame_start = nullptr;
    int* stream_vote_counters[2] = {};              // per traversal stream
    Ticket stream_tiles[2];
};

DeviceState& device_state() {
    static DeviceState states[64];
    static std::mutex init_lock;
    int dev = 0;
    HGB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { std::fprintf(stderr, "hagrid_b200: device index out of range\n"); std::abort(); }
    DeviceState& st = states[dev];
    std::lock_guard<std::mutex> guard(init_lock);
    if (!st.words) {
        HGB_CUDA(cudaMalloc(&st.words, 256));
        HGB_CUDA(cudaMemset(st.words, 0, 256));
        st.vote_counter = st.words;
        st.tiles.word = reinterpret_cast<unsigned*>(st.words + 8);
        st.stream_vote_counters[0] = st.words + 16;
        st.stream_tiles[0].word = reinterpret_cast<unsigned*>(st.words + 24);
        st.stream_vote_counters[1] = st.words + 32;
        st.stream_tiles[1].word = reinterpret_cast<unsigned*>(st.words + 40);
        HGB_CUDA(cudaHostAlloc(&st.seen_host, DeviceState::kSeenBuffers * 4 * sizeof(int), cudaHostAllocMapped));
        for (int i = 0; i < DeviceState::kSeenBuffers; i++) {
            st.seen[i].layout = st.words + 48 + 4 * i;
            st.seen[i].feedback_dev = st.words + 48 + 4 * i + 2;
            st.seen[i].host = st.seen_host + 4 * i;
            for (int k = 0; k < 4; k++) st.seen[i].host[k] = 0;
        }
        HGB_CUDA(cudaDeviceGetAttribute(&st.num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    return st;
}

void prepare_streams(DeviceState& st) {
    if (st.frame_start) return;
    for (int i = 0; i < DeviceState::kStreams; i++) {
        HGB_CUDA(cudaStreamCreateWithFlags(&st.streams[i], cudaStreamNonBlocking));
        HGB_CUDA(cudaEventCreateWithFlags(&st.stream_done[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < DeviceState::kMaxChunks; i++) {
        HGB_CUDA(cudaEventCreateWithFlags(&st.uploaded[i], cudaEventDisableTiming));
        HGB_CUDA(cudaEventCreateWithFlags(&st.traced[i], cudaEventDisableTiming));
    }
    HGB_CUDA(cudaEventCreateWithFlags(&st.frame_start, cudaEventDisableTiming));
}

// 0: per thread, buffer order   1: persistent warps, majority-scheduled   2: per thread, re-tiled when a raster is detected
// 4: resident warps pulling tiles, re-tiled when a raster is detected
// 3 (default): 4 for buffers that are (or may be) rasters, 1 once a buffer is known not to be one
std::atomic<int> g_variant{-1};

// Host-buffer frames: rays per full-size chunk (tuned on B200/PCIe 5, profiles/r02_e2e.md: 256 K rays = 8.6 MB up, 4.3 MB down)
std::atomic<int> g_host_frame_chunk{256 * 1024};
// trace_two_waves: chunks a frame is cut into (their chains alternate between two streams)
std::atomic<int> g_two_wave_chunks{1};
// incoherent buffers smaller than this many rays are traced one thread per ray: the persistent voting warps pay off from
// about 600 K rays (tools/gpu_incoherent_threshold.py: 259 K rays of the C5 second wave 0.363 vs 0.432 ms, 1 M rays 0.785 vs 0.722)
std::atomic<int> g_vote_min_rays{768 << 10};
// rasters smaller than this many rays are traced one thread per ray (re-tiled): a small launch does not fill the resident
// warps. Two lines (profiles/r02_traverse_experiments.md): a launch that has a ticket list for its buffer, or is about to
// time its tiles for one, wins down to an eighth of a 1920x1080 frame (C2 1 M / 518 K / 259 K rays: 0.083 / 0.053 / 0.046 ms
// against 0.098 / 0.067 / 0.049 one thread per ray); without a list half a frame is on the line.
std::atomic<int> g_tile_min_rays{128 << 10};
std::atomic<int> g_tile_cold_min_rays{1280 << 10};

int traverse_variant() {
    int v = g_variant.load();
    if (v < 0) {
        const char* e = std::getenv("HGB_TRAVERSE_VARIANT");
        v = e ? std::atoi(e) : 3;
        g_variant.store(v);
    }
    return v;
}

/// Ray binning for incoherent buffers (the north star's "warp-sorted ray packets"): key = direction octant (3 bits) above
/// the top-level cell the ray enters the grid in; rays sorted by that key start their march in the same top-level
/// cell, heading the same way. The sort is this library's own (primitives.cuh) and runs inside the traversal call.
/// Measured on C3 and on the second wave of C5 (profiles/r02_traverse_experiments.md): off by default.
__global__ void __launch_bounds__(256) ray_bin_keys(const __grid_constant__ TraversalParams P, const Ray* __restrict__ rays, int num_rays,
                                                    int* __restrict__ keys, int* __restrict__ order) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= num_rays) return;
    RayState r;
    const bool ok = start_ray(r, P, rays, i);
    const int top_z = P.dims_z >> P.shift;
    int key = P.top_x * P.top_y * top_z * 8;               // rays that miss the grid: behind everything else
    if (ok) key = octant_of(r) * (P.top_x * P.top_y * top_z) + (r.vx >> P.shift) + P.top_x * ((r.vy >> P.shift) + P.top_y * (r.vz >> P.shift));
    keys[i] = key;
    order[i] = i;
}

std::atomic<int> g_ray_sort{0};

// "tile_order": cost classes (bits, 1 ... 8) the tiles of a raster are handed out by; 0 = always buffer order
std::atomic<int> g_tile_order_bits{8};
// "tile_split" / "tile_split_log" / "tile_split_share": how many of the most expensive tiles are handed out in 2^tile_split_log
// parts, if such a tile alone costs at least that many percent of one resident warp's share of the launch
std::atomic<int> g_tile_split{256}, g_tile_split_log{2}, g_tile_split_share{50};
std::atomic<int> g_tile_history_epoch{0};      // bumped when one of these options changes: what was learnt under other settings is dropped

/// What the next launch of traverse_tiles on `buf` gets (TileHistory): the newest order -- the one made from timed
/// launch k is used from launch k + 2 on, so that it is made while launch k + 1 runs and nobody waits for it -- and a
/// cost array when this launch is one of those that time their tiles: the first ones on a buffer, later every 16th,
/// and never while an order is still being made from the previous timing. Returns whether the launch is timed.
bool tile_history(DeviceState& st, DeviceState::SeenBuffer& buf, int num_rays, TileHistory& history, cudaStream_t stream) {
    const int tiles = (num_rays + 31) / 32;
    if (!st.side) HGB_CUDA(cudaStreamCreateWithFlags(&st.side, cudaStreamNonBlocking));
    if (buf.history_tiles < tiles) {
        if (buf.order_pending >= 0) HGB_CUDA(cudaEventSynchronize(buf.ordered));
        if (buf.history) { HGB_CUDA(cudaDeviceSynchronize()); HGB_CUDA(cudaFree(buf.history)); }
        const size_t list = size_t(tiles) + DeviceState::kMaxSplitExtra;
        const size_t ints = 2 * list + 4 * size_t(tiles) + prim::sort_scratch_ints(tiles) + 2;
        HGB_CUDA(cudaMalloc(&buf.history, ints * sizeof(int) + size_t(tiles) * sizeof(unsigned short)));
        buf.tickets[0] = buf.history; buf.tickets[1] = buf.history + list;
        buf.sort_space = buf.history + 2 * list;
        buf.stats = reinterpret_cast<unsigned*>(buf.history + ints - 2);
        buf.cost = reinterpret_cast<unsigned short*>(buf.history + ints);
        buf.sorted = nullptr;
        buf.history_tiles = tiles;
        buf.order_ready = buf.order_pending = -1; buf.tile_launches = 0;
        if (!buf.timed) {
            HGB_CUDA(cudaEventCreateWithFlags(&buf.timed, cudaEventDisableTiming));
            HGB_CUDA(cudaEventCreateWithFlags(&buf.ordered, cudaEventDisableTiming));
        }
    }
    if (buf.history_epoch != g_tile_history_epoch.load()) {
        if (buf.order_pending >= 0) HGB_CUDA(cudaEventSynchronize(buf.ordered));
        buf.history_epoch = g_tile_history_epoch.load();
        buf.order_ready = buf.order_pending = -1; buf.tile_launches = 0;
    }
    const int k = buf.tile_launches++;
    if (buf.order_pending >= 0 && k >= buf.pending_since + 2) {
        // made during the launch before this one; the wait is a formality unless the device was idle in between
        HGB_CUDA(cudaStreamWaitEvent(stream, buf.ordered, 0));
        buf.order_ready = buf.order_pending;
        buf.order_pending = -1;
    }
    const bool timed = buf.order_pending < 0 && (k < 4 || (k & 15) == 0);
    if (timed) buf.pending_since = k;
    history.cost = timed ? buf.cost : nullptr;
    if (buf.order_ready >= 0 && !timed) {
        history.tickets = buf.tickets[buf.order_ready];
        history.extra_tickets = buf.extra[buf.order_ready];
        history.split_log = buf.split_log[buf.order_ready];
    } else if (buf.order_ready >= 0) {
        history.tickets = buf.sorted;        // whole tiles: a part's time says nothing about its tile
    }
    return timed;
}

/// After a timed launch on `stream`: the ticket list for later launches, made on the side stream from the costs that
/// launch leaves behind, into the list no launch is reading. (Launches of one buffer are ordered among themselves: on
/// the legacy default stream, or on streams that start behind it and are joined back into it, launch_two_waves.)
void order_tiles(DeviceState& st, DeviceState::SeenBuffer& buf, int num_rays, int bits, cudaStream_t stream) {
    const int tiles = (num_rays + 31) / 32;
    const int target = buf.order_ready == 0 ? 1 : 0;
    int* keys = buf.sort_space; int* vals = keys + tiles; int* keys_alt = vals + tiles; int* vals_alt = keys_alt + tiles;
    int* scratch = vals_alt + tiles;
    const int split_log = std::max(0, std::min(5, g_tile_split_log.load()));
    const int split = split_log ? std::min(std::min(tiles, g_tile_split.load()), DeviceState::kMaxSplit) : 0;
    const int extra = split * ((1 << split_log) - 1);
    const int warps = std::min(st.num_sms * kTileBlocksPerSm, round_div(tiles * 32, kTileBlock)) * (kTileBlock / 32);
    HGB_CUDA(cudaEventRecord(buf.timed, stream));
    HGB_CUDA(cudaStreamWaitEvent(st.side, buf.timed, 0));
    tile_cost_stats<<<1, 1024, 0, st.side>>>(buf.cost, tiles, buf.stats); count_launch();
    tile_order_keys<<<round_div(tiles, 256), 256, 0, st.side>>>(buf.cost, tiles, bits, buf.stats, keys, vals); count_launch();
    buf.sorted = prim::sort_pairs(keys, vals, keys_alt, vals_alt, tiles, bits, scratch, st.side) ? vals_alt : vals;
    tile_tickets<<<round_div(tiles + extra, 256), 256, 0, st.side>>>(buf.sorted, buf.cost, buf.stats, tiles, split, split_log,
                                                                      g_tile_split_share.load(), warps, buf.tickets[target]); count_launch();
    HGB_CUDA(cudaEventRecord(buf.ordered, st.side));
    buf.order_pending = target;
    buf.extra[target] = extra; buf.split_log[target] = split_log;
}

/// Enqueues one traversal launch on `stream`: 1 = persistent voting warps (needs `vote_counter`), 4 = resident
/// warps pulling tiles (needs `ticket`), otherwise one thread per ray; 2 and 4 re-tile by the raster width in
/// `layout[0]` (device) or `host_width`.
template <typename CellT, bool kPrimId>
void enqueue(const Grid& grid, const CellT* cells, const Tri* tris, const Ray* rays, Hit* hits, int num_rays,
             int variant, const int* layout, int host_width, int* vote_counter, Ticket& ticket, int num_sms, cudaStream_t stream,
             int* feedback = nullptr, const int* order = nullptr, TileHistory history = TileHistory{nullptr, nullptr, 0, 0}) {
    auto entries = reinterpret_cast<const uint32_t*>(grid.entries);
    const TraversalParams P = params_of(grid);
    if (variant == 4) {
        const int blocks = min(num_sms * kTileBlocksPerSm, round_div(num_rays, kTileBlock));
        if (history.cost)
            traverse_tiles<CellT, kPrimId, true><<<blocks, kTileBlock, 0, stream>>>(
                P, entries, cells, grid.ref_ids, tris, rays, hits, num_rays, layout, host_width, ticket.word, ticket.base, feedback, history);
        else
            traverse_tiles<CellT, kPrimId, false><<<blocks, kTileBlock, 0, stream>>>(
                P, entries, cells, grid.ref_ids, tris, rays, hits, num_rays, layout, host_width, ticket.word, ticket.base, feedback, history);
        // one fetch per ticket (fetch_tile)
        ticket.base += unsigned((num_rays + 31) >> 5) + unsigned(history.extra_tickets);
        count_launch();
    } else if (variant == 1) {
        HGB_CUDA(cudaMemsetAsync(vote_counter, 0, sizeof(int), stream));
        // every warp reserves two blocks of rays up front: no more warps than there are blocks
        const int blocks = max(1, min(num_sms * kVoteBlocksPerSm, round_div(num_rays, 2 * kVoteBlock * (kBlockThreads / 32))));
        traverse_voting<CellT, kPrimId><<<blocks, kBlockThreads, 0, stream>>>(
            P, entries, cells, grid.ref_ids, tris, rays, hits, num_rays, vote_counter, order); count_launch();
    } else {
        traverse_per_thread<CellT, kPrimId><<<round_div(num_rays, 128), 128, 0, stream>>>(
            P, entries, cells, grid.ref_ids, tris, rays, hits, num_rays, layout, host_width); count_launch();
    }
}

/// What kind of buffer is this? A new one (other address or size) is looked at on the device before its first
/// launch (a wait of about 10 us, once per buffer, instead of tracing its first frame with a kernel that may be the
/// wrong one by a factor of 1.7); the answer is remembered on the host. A buffer known as a raster is not looked at
/// again — the tile kernel reports when its contents stop behaving like one (feedback words) — and a buffer known
/// as incoherent is looked at before every launch (a 10 us kernel in front of one that takes several hundred):
/// callers reuse ray buffers. Work is enqueued on the legacy default stream. Returns the buffer's cache entry.
DeviceState::SeenBuffer* classify_buffer(DeviceState& st, const Ray* rays, int num_rays) {
    DeviceState::SeenBuffer* buf = nullptr;
    st.launches++;
    for (auto& e : st.seen)
        if (e.rays == rays && e.count == num_rays) buf = &e;
    const bool same = buf != nullptr;
    if (!same) {
        buf = &st.seen[0];
        for (auto& e : st.seen)
            if (e.used < buf->used) buf = &e;
        buf->cls = -1; buf->feedback_armed = false;
        // what was learnt about another buffer's tiles says nothing about this one (a list still being made for it is
        // made for nobody: the side stream finishes it before it starts the next)
        buf->tile_launches = 0; buf->order_ready = buf->order_pending = -1;
    }
    buf->used = st.launches;
    volatile int* answer = buf->host;
    volatile int* fb = buf->host + 2;
    if (!same) ;                                                          // nothing known yet
    else if (buf->cls < 0 && *answer >= 0) buf->cls = *answer;            // the first look has finished
    else if (buf->cls == 0 && *answer > 0) buf->cls = *answer;            // the latest look found a raster again
    else if (buf->cls > 0 && buf->feedback_armed) {
        // the launches reported since the last look
        const int now_mixed = fb[0], now_launches = fb[1];
        const int mixed = now_mixed - buf->feedback_mixed, launches = now_launches - buf->feedback_launches;
        buf->feedback_mixed = now_mixed; buf->feedback_launches = now_launches;
        const int tiles = (num_rays + 31) / 32;
        if (launches > 0 && (long long)mixed * 4 > (long long)launches * tiles) {     // most warps mixed: not camera rays any more
            buf->cls = 0;
            buf->feedback_armed = false;
            *answer = 0;
        }
    }
    if (!same || buf->cls == 0) {
        *answer = -1;
        int* host_alias = nullptr;
        HGB_CUDA(cudaHostGetDevicePointer(&host_alias, buf->host, 0));
        detect_raster<<<1, 256>>>(rays, num_rays, buf->layout, host_alias); count_launch();
        if (!same) {
            HGB_CUDA(cudaStreamSynchronize(0));
            buf->cls = *answer;
        }
        buf->rays = rays;
        buf->count = num_rays;
    }
    return buf;
}

template <typename CellT, bool kPrimId>
void launch(const Grid& grid, const CellT* cells, const Tri* tris, const Ray* rays, Hit* hits, int num_rays) {
    if (num_rays <= 0) return;
    DeviceState& st = device_state();
    std::lock_guard<std::mutex> guard(st.lock);
    int variant = traverse_variant();
    int* feedback = nullptr;
    DeviceState::SeenBuffer* buf = nullptr;
    if (variant >= 2 && variant <= 4) {
        buf = classify_buffer(st, rays, num_rays);
        if (variant == 3) {
            // small buffers do not fill the resident-warp kernels (7 104 warps on 148 SMs): one thread per ray then
            if (buf->cls == 0) variant = num_rays < g_vote_min_rays.load() ? 0 : 1;
            else               variant = num_rays < (g_tile_order_bits.load() > 0 ? g_tile_min_rays : g_tile_cold_min_rays).load() ? 2 : 4;
        }
        if (variant == 4 && buf->cls > 0) {
            feedback = buf->feedback_dev;
            if (!buf->feedback_armed) {
                // fresh baseline: counters restart from zero for this buffer
                HGB_CUDA(cudaMemsetAsync(buf->feedback_dev, 0, 2 * sizeof(int), 0));
                buf->host[2] = buf->host[3] = 0;
                buf->feedback_mixed = 0; buf->feedback_launches = 0; buf->feedback_armed = true; buf->feedback_tick = 0;
            }
        }
    }
    const bool tiled = variant == 2 || variant == 4;
    const int* order = nullptr;
    if (variant == 1 && g_ray_sort.load()) {
        // experiment ("ray_sort"): bin the rays by (octant, entry cell) first; scratch is kept per device
        const TraversalParams P = params_of(grid);
        const size_t need = size_t(num_rays) * 4 + prim::sort_scratch_ints(num_rays);
        if (need > st.sort_capacity) {
            if (st.sort_scratch) HGB_CUDA(cudaFree(st.sort_scratch));
            HGB_CUDA(cudaMalloc(&st.sort_scratch, need * sizeof(int)));
            st.sort_capacity = need;
        }
        int* keys = st.sort_scratch; int* idx = keys + num_rays; int* keys_alt = idx + num_rays; int* idx_alt = keys_alt + num_rays;
        ray_bin_keys<<<round_div(num_rays, 256), 256>>>(P, rays, num_rays, keys, idx); count_launch();
        const int cells_top = P.top_x * P.top_y * (P.dims_z >> P.shift) * 8 + 1;
        int bits = 1;
        while ((1 << bits) < cells_top) bits++;
        order = prim::sort_pairs(keys, idx, keys_alt, idx_alt, num_rays, bits, idx_alt + num_rays) ? idx_alt : idx;
    }
    TileHistory history{nullptr, nullptr, 0, 0};
    bool timed = false;
    const int order_bits = g_tile_order_bits.load();
    if (variant == 4 && buf && buf->cls > 0 && order_bits > 0) {
        timed = tile_history(st, *buf, num_rays, history, 0);
        // the one launch of a small buffer that neither has a list nor makes one (the second on the buffer)
        if (traverse_variant() == 3 && !timed && !history.tickets && num_rays < g_tile_cold_min_rays.load()) { variant = 2; feedback = nullptr; }
    }
    enqueue<CellT, kPrimId>(grid, cells, tris, rays, hits, num_rays, variant, tiled && buf ? buf->layout : nullptr, 0,
                            st.vote_counter, st.tiles, st.num_sms, 0, feedback, order, history);
    if (timed) order_tiles(st, *buf, num_rays, order_bits, 0);
    if (feedback) {
        // copied back after launches 1, 2, 4 and then every 8th: one 8-byte copy in eight launches
        const int tick = ++buf->feedback_tick;
        if (tick <= 2 || (tick & 3) == 0 && (tick == 4 || (tick & 7) == 0))
            HGB_CUDA(cudaMemcpyAsync(buf->host + 2, buf->feedback_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, 0));
    }
    HGB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// Host-buffer frames (the interactive loop of src/main.cpp:599-613: upload the rays, trace, download
// the hits). The three steps of one frame are cut into chunks that travel round-robin over a few
// streams, so the upload of chunk i+1, the traversal of chunk i and the download of chunk i-1 overlap;
// PCIe is full duplex, so a frame costs about as long as its 32 B/ray upload alone.
// ---------------------------------------------------------------------------

/// Host mirror of detect_raster for a buffer in host memory: the row length W of a W x H raster of
/// smoothly varying rays (W % 8 == 0, H % 4 == 0), else 0. Any answer is safe (the re-tiling is a
/// bijection for every such W); the scan stops at the first row break, so it reads W rays, not n.
int host_raster_width(const Ray* rays, int n) {
    constexpr int kScan = 16384;
    if (n < 4 * 64) return 0;
    auto delta = [&](int i, float d[6]) {
        const Ray& a = rays[i]; const Ray& b = rays[i - 1];
        d[0] = a.org.x - b.org.x; d[1] = a.org.y - b.org.y; d[2] = a.org.z - b.org.z;
        d[3] = a.dir.x - b.dir.x; d[4] = a.dir.y - b.dir.y; d[5] = a.dir.z - b.dir.z;
    };
    float step[6];
    delta(1, step);
    float tol = 0.0f;
    for (int k = 0; k < 6; k++) tol = std::fmax(tol, std::fabs(step[k]));
    tol *= 0.25f;
    if (!(tol > 0.0f)) return 0;
    auto continues = [&](int i) {
        float d[6];
        delta(i, d);
        float err = 0.0f;
        for (int k = 0; k < 6; k++) err = std::fmax(err, std::fabs(d[k] - step[k]));
        return err <= tol;
    };
    const int limit = n < kScan ? n : kScan;
    int width = 2;
    while (width < limit && continues(width)) width++;
    if (width >= limit || width < 64 || width % kTileW != 0 || n % width != 0 || (n / width) % kTileH != 0) return 0;
    for (int k = 0; k < 62; k++)
        if (!continues(width + 1 + k) || !continues(n - width + 1 + k)) return 0;
    constexpr int kSpot = 1024;                      // same spot check as detect_raster
    int misses = 0;
    for (int k = 0; k < kSpot; k++) {
        int i = int((long long)n * k / kSpot) + 1 + (k * 37) % 61;
        if (i >= n) i = n - 1;
        if (i % width != 0 && !continues(i)) misses++;
    }
    return misses > kSpot / 16 ? 0 : width;
}

template <typename CellT, bool kPrimId>
void launch_host_frame(const Grid& grid, const CellT* cells, const Tri* tris, const Ray* host_rays, Hit* host_hits,
                       int num_rays, Ray* dev_rays, Hit* dev_hits) {
    if (num_rays <= 0) return;
    DeviceState& st = device_state();
    std::lock_guard<std::mutex> guard(st.lock);
    prepare_streams(st);

    int variant = traverse_variant();
    int width = 0;
    if (variant >= 2 && variant <= 4) {
        width = host_raster_width(host_rays, num_rays);
        variant = width > 0 ? (variant == 2 ? 2 : 4) : (variant == 3 ? 1 : 0);
    }
    HGB_CUDA(cudaEventRecord(st.frame_start, 0));           // frames are ordered after earlier default-stream work
    for (int i = 0; i < DeviceState::kStreams; i++) HGB_CUDA(cudaStreamWaitEvent(st.streams[i], st.frame_start, 0));
    // Chunks are whole 4-row tile bands of a raster, else whole blocks. Full-size chunks keep the copy
    // engines busy with few, large transfers; the last one is split 1/2, 1/4, 1/4 because nothing
    // overlaps the final traversal (+ download).
    const int granule = width > 0 ? width * kTileH : kBlockThreads;
    const int chunk = g_host_frame_chunk.load();
    std::vector<int> sizes;
    {
        long long full = std::max<long long>(granule, (long long)round_div(chunk, granule) * granule);
        full = std::max<long long>(full, (long long)round_div(round_div(num_rays, DeviceState::kMaxChunks - 4), granule) * granule);
        long long rest = num_rays;
        while (rest > full + full / 2) { sizes.push_back(int(full)); rest -= full; }
        const long long half = std::min<long long>(rest, (long long)round_div(int(rest / 2), granule) * granule);
        const long long quarter = std::min<long long>(rest - half, (long long)round_div(int(rest / 4), granule) * granule);
        if (half > 0) sizes.push_back(int(half));
        if (quarter > 0) sizes.push_back(int(quarter));
        if (rest - half - quarter > 0) sizes.push_back(int(rest - half - quarter));
    }
    cudaStream_t up = st.streams[0], down = st.streams[1];
    // HGB_FRAME_TRACE=1: time stamps of every chunk's upload, traversal and download (diagnosis only)
    static const bool trace = std::getenv("HGB_FRAME_TRACE") != nullptr;
    std::vector<cudaEvent_t> stamps;
    auto stamp = [&](cudaStream_t on) {
        if (!trace) return;
        cudaEvent_t e; HGB_CUDA(cudaEventCreate(&e)); HGB_CUDA(cudaEventRecord(e, on)); stamps.push_back(e);
    };
    stamp(up);
    long long begin = 0;
    for (size_t c = 0; c < sizes.size(); begin += sizes[c], c++) {
        const int count = sizes[c];
        cudaStream_t run = st.streams[2 + (c & 1)];      // consecutive traversals may overlap (tails of incoherent chunks)
        HGB_CUDA(cudaMemcpyAsync(dev_rays + begin, host_rays + begin, sizeof(Ray) * size_t(count), cudaMemcpyHostToDevice, up));
        HGB_CUDA(cudaEventRecord(st.uploaded[c], up));
        stamp(up);
        HGB_CUDA(cudaStreamWaitEvent(run, st.uploaded[c], 0));
        // an incoherent chunk below the voting kernel's break-even size is traced one thread per ray
        const int chunk_variant = traverse_variant() == 3 && variant == 1 && count < g_vote_min_rays.load() ? 0 : variant;
        enqueue<CellT, kPrimId>(grid, cells, tris, dev_rays + begin, dev_hits + begin, count, chunk_variant, nullptr, width,
                                st.stream_vote_counters[c & 1], st.stream_tiles[c & 1], st.num_sms, run);
        stamp(run);
        HGB_CUDA(cudaEventRecord(st.traced[c], run));
        HGB_CUDA(cudaStreamWaitEvent(down, st.traced[c], 0));
        HGB_CUDA(cudaMemcpyAsync(host_hits + begin, dev_hits + begin, sizeof(Hit) * size_t(count), cudaMemcpyDeviceToHost, down));
        stamp(down);
    }
    if (trace) {
        HGB_CUDA(cudaDeviceSynchronize());
        std::fprintf(stderr, "frame trace (ms since the first upload was enqueued): chunk rays uploaded traced downloaded\n");
        for (size_t c = 0; c < sizes.size(); c++) {
            float t[3];
            for (int k = 0; k < 3; k++) HGB_CUDA(cudaEventElapsedTime(&t[k], stamps[0], stamps[1 + 3 * c + k]));
            std::fprintf(stderr, "  %2zu %8d %7.3f %7.3f %7.3f\n", c, sizes[c], t[0], t[1], t[2]);
        }
        for (cudaEvent_t e : stamps) cudaEventDestroy(e);
    }
    HGB_CUDA(cudaGetLastError());
    for (int i = 0; i < DeviceState::kStreams; i++) {
        HGB_CUDA(cudaEventRecord(st.stream_done[i], st.streams[i]));
        HGB_CUDA(cudaStreamWaitEvent(0, st.stream_done[i], 0));   // a timer on the default stream brackets the frame
    }
    for (int i = 0; i < DeviceState::kStreams; i++) HGB_CUDA(cudaStreamSynchronize(st.streams[i]));   // hits are in host memory
}

/// Device-resident rays, hits wanted in host memory (the second wave of a two-wave frame: its rays were made on the
/// device): the buffer is traced in chunks on the default stream and every chunk's hits start their way to the host
/// as soon as it is traced. Returns when `host_hits` is complete.
template <typename CellT, bool kPrimId>
void launch_to_host(const Grid& grid, const CellT* cells, const Tri* tris, const Ray* dev_rays, Hit* dev_hits, Hit* host_hits, int num_rays) {
    if (num_rays <= 0) return;
    DeviceState& st = device_state();
    std::lock_guard<std::mutex> guard(st.lock);
    prepare_streams(st);
    cudaStream_t down = st.streams[1];
    HGB_CUDA(cudaEventRecord(st.frame_start, 0));           // ordered after earlier default-stream work (the rays are made there)
    for (int i = 1; i < DeviceState::kStreams; i++) HGB_CUDA(cudaStreamWaitEvent(st.streams[i], st.frame_start, 0));
    // Chunks alternate between the two traversal streams: a launch over incoherent rays ends with a long tail of a few
    // marching rays, and the next chunk's warps move in as the previous chunk's run dry
    constexpr int kChunks = 4;
    const long long step = ((long long)round_div(num_rays, kChunks) + kBlockThreads - 1) / kBlockThreads * kBlockThreads;
    int c = 0;
    for (long long begin = 0; begin < num_rays; begin += step, c++) {
        const int count = int(std::min<long long>(step, num_rays - begin));
        cudaStream_t run = st.streams[2 + (c & 1)];
        // second-wave rays are incoherent by construction: no look at their layout
        int variant = traverse_variant();
        if (variant == 3) variant = count < g_vote_min_rays.load() ? 0 : 1;
        else if (variant != 1) variant = 0;
        enqueue<CellT, kPrimId>(grid, cells, tris, dev_rays + begin, dev_hits + begin, count, variant, nullptr, 0,
                                st.stream_vote_counters[c & 1], st.stream_tiles[c & 1], st.num_sms, run);
        HGB_CUDA(cudaEventRecord(st.traced[c], run));
        HGB_CUDA(cudaStreamWaitEvent(down, st.traced[c], 0));
        HGB_CUDA(cudaMemcpyAsync(host_hits + begin, dev_hits + begin, sizeof(Hit) * size_t(count), cudaMemcpyDeviceToHost, down));
    }
    HGB_CUDA(cudaGetLastError());
    for (int i = 1; i < DeviceState::kStreams; i++) {
        HGB_CUDA(cudaEventRecord(st.stream_done[i], st.streams[i]));
        HGB_CUDA(cudaStreamWaitEvent(0, st.stream_done[i], 0));   // a timer on the default stream brackets the work
    }
    HGB_CUDA(cudaStreamSynchronize(down));
}

/// One two-wave frame with everything resident (BASELINE config C5): trace, count, bounce rays, trace, count. The
/// frame can be cut into chunks ("two_wave_chunks") whose chains alternate between two streams, the idea being that one
/// chunk's work fills the tails of the other's launches; measured on the 7.8 M-triangle frame it loses (1.42 / 1.59 /
/// 1.93 / 1.96 / 2.68 ms for 1 / 2 / 3 / 4 / 8 chunks: every chunk's launches bring their own tail and the smaller
/// launches fall off the tile kernel), so a frame is one chain by default. Joined back into the legacy default stream.
template <typename CellT>
void launch_two_waves(const Grid& grid, const CellT* cells, const Tri* tris, int num_tris, const Ray* rays, int num_rays,
                      const int* keys, float offset, float tmax, unsigned seed, Hit* hits_primary, Ray* bounce, Hit* hits_bounce,
                      unsigned long long* counters) {
    if (num_rays <= 0) return;
    DeviceState& st = device_state();
    std::lock_guard<std::mutex> guard(st.lock);
    prepare_streams(st);
    int forced = traverse_variant();
    int width = 0;
    DeviceState::SeenBuffer* buf = nullptr;
    if (forced >= 2 && forced <= 4) { buf = classify_buffer(st, rays, num_rays); width = std::max(0, buf->cls); }
    const int granule = width > 0 ? width * kTileH : kBlockThreads;
    const int chunks = std::max(1, std::min(g_two_wave_chunks.load(), num_rays / (256 << 10)));
    const long long step = ((long long)round_div(num_rays, chunks) + granule - 1) / granule * granule;
    HGB_CUDA(cudaEventRecord(st.frame_start, 0));
    for (int i = 2; i < DeviceState::kStreams; i++) HGB_CUDA(cudaStreamWaitEvent(st.streams[i], st.frame_start, 0));
    int c = 0;
    for (long long begin = 0; begin < num_rays; begin += step, c++) {
        const int count = int(std::min<long long>(step, num_rays - begin));
        cudaStream_t run = st.streams[2 + (c & 1)];
        int first = forced, second = forced;
        if (forced == 3) {
            const bool listed = count == num_rays && g_tile_order_bits.load() > 0;      // see below: a frame in one piece
            first = width > 0 ? (count < (listed ? g_tile_min_rays : g_tile_cold_min_rays).load() ? 2 : 4) : (count < g_vote_min_rays.load() ? 0 : 1);
            second = count < g_vote_min_rays.load() ? 0 : 1;
        } else if (forced == 2 || forced == 4) {
            if (width <= 0) first = 0;
            second = 0;
        }
        // a frame traced in one piece has its tiles handed out by what the frames before it cost (TileHistory)
        TileHistory history{nullptr, nullptr, 0, 0};
        const int order_bits = g_tile_order_bits.load();
        const bool timed = first == 4 && count == num_rays && buf && width > 0 && order_bits > 0 && tile_history(st, *buf, count, history, run);
        enqueue<CellT, true>(grid, cells, tris, rays + begin, hits_primary + begin, count, first, nullptr, width,
                             st.stream_vote_counters[c & 1], st.stream_tiles[c & 1], st.num_sms, run, nullptr, nullptr, history);
        if (timed) order_tiles(st, *buf, count, order_bits, run);
        if (counters) count_hits_on(run, hits_primary + begin, count, counters);
        generate_bounce_rays_on(run, tris, num_tris, rays + begin, hits_primary + begin, count, offset, tmax, seed, bounce + begin,
                                keys ? keys + begin : nullptr, int(begin));
        enqueue<CellT, true>(grid, cells, tris, bounce + begin, hits_bounce + begin, count, second, nullptr, 0,
                             st.stream_vote_counters[c & 1], st.stream_tiles[c & 1], st.num_sms, run);
        if (counters) count_hits_on(run, hits_bounce + begin, count, counters);
    }
    HGB_CUDA(cudaGetLastError());
    for (int i = 2; i < DeviceState::kStreams; i++) {
        HGB_CUDA(cudaEventRecord(st.stream_done[i], st.streams[i]));
        HGB_CUDA(cudaStreamWaitEvent(0, st.stream_done[i], 0));
    }
}

template <bool kPrimId>
void dispatch(const Grid& grid, const Tri* tris, const Ray* rays, Hit* hits, int num_rays) {
    if (grid.small_cells) launch<SmallCell, kPrimId>(grid, grid.small_cells, tris, rays, hits, num_rays);
    else                  launch<Cell, kPrimId>(grid, grid.cells, tris, rays, hits, num_rays);
}

} // namespace

void setup_traversal(const Grid& grid) {
    // Nothing to capture: the constants the reference stores in __constant__ memory here (src/traverse.cu:97-108)
    // are derived from the grid at every launch (params_of). The call creates the per-device launch state, so
    // that the first traverse_grid does not pay for it.
    (void)grid;
    device_state();
}

namespace {

FrameParams frame_params(const FrameCamera& cam, float clip, int width, int height) {
    FrameParams F;
    F.eye_x = cam.eye.x; F.eye_y = cam.eye.y; F.eye_z = cam.eye.z;
    F.dir_x = cam.dir.x; F.dir_y = cam.dir.y; F.dir_z = cam.dir.z;
    F.right_x = cam.right.x; F.right_y = cam.right.y; F.right_z = cam.right.z;
    F.up_x = cam.up.x; F.up_y = cam.up.y; F.up_z = cam.up.z;
    F.clip = clip; F.width = width; F.height = height;
    return F;
}

template <typename CellT>
void launch_frame(const Grid& grid, const CellT* cells, const Tri* tris, const FrameParams& F, int mode, unsigned* pixels) {
    DeviceState& st = device_state();
    std::lock_guard<std::mutex> guard(st.lock);
    auto entries = reinterpret_cast<const uint32_t*>(grid.entries);
    const TraversalParams P = params_of(grid);
    const int num_pixels = F.width * F.height;
    const int blocks = std::min(st.num_sms * kTileBlocksPerSm, round_div(num_pixels, kTileBlock));
    Ticket& t = st.tiles;
    if (mode == 0)      render_tiles<CellT, 0><<<blocks, kTileBlock>>>(P, F, entries, cells, grid.ref_ids, tris, pixels, t.word, t.base);
    else if (mode == 1) render_tiles<CellT, 1><<<blocks, kTileBlock>>>(P, F, entries, cells, grid.ref_ids, tris, pixels, t.word, t.base);
    else                render_tiles<CellT, 2><<<blocks, kTileBlock>>>(P, F, entries, cells, grid.ref_ids, tris, pixels, t.word, t.base);
    t.base += unsigned((num_pixels + 31) >> 5);
    count_launch();
    HGB_CUDA(cudaGetLastError());
}

} // namespace

FrameCamera make_camera(const vec3& eye, const vec3& center, const vec3& up, float fov, float ratio) {
    // gen_camera, src/main.cpp:42-50 (host IEEE arithmetic, libm tanf)
    FrameCamera cam;
    const float f = tanf(M_PI * fov / 360);
    cam.dir = normalize(center - eye);
    cam.right = normalize(cross(cam.dir, up)) * (f * ratio);
    cam.up = normalize(cross(cam.right, cam.dir)) * f;
    cam.eye = eye;
    return cam;
}

void generate_rays(const FrameCamera& cam, float clip, int width, int height, Ray* rays) {
    if (width <= 0 || height <= 0) return;
    const FrameParams F = frame_params(cam, clip, width, height);
    generate_camera_rays<<<round_div(width * height, 256), 256>>>(F, rays); count_launch();
    HGB_CUDA(cudaGetLastError());
}

void render_frame(const Grid& grid, const Tri* tris, const FrameCamera& cam, float clip, int width, int height,
                  int mode, unsigned* pixels) {
    if (width <= 0 || height <= 0) return;
    const FrameParams F = frame_params(cam, clip, width, height);
    if (grid.small_cells) launch_frame<SmallCell>(grid, grid.small_cells, tris, F, mode, pixels);
    else                  launch_frame<Cell>(grid, grid.cells, tris, F, mode, pixels);
}

bool set_traversal_option(const char* key, int value) {
    if (!std::strcmp(key, "traverse_variant")) { g_variant.store(value); return true; }
    if (!std::strcmp(key, "host_frame_chunk_rays")) { g_host_frame_chunk.store(value > 0 ? value : 256 * 1024); return true; }
    if (!std::strcmp(key, "ray_sort")) { g_ray_sort.store(value != 0); return true; }
    if (!std::strcmp(key, "tile_order")) { g_tile_order_bits.store(std::max(0, std::min(12, value))); g_tile_history_epoch++; return true; }
    if (!std::strcmp(key, "tile_split")) { g_tile_split.store(std::max(0, value)); g_tile_history_epoch++; return true; }
    if (!std::strcmp(key, "tile_split_log")) { g_tile_split_log.store(value); g_tile_history_epoch++; return true; }
    if (!std::strcmp(key, "tile_split_share")) { g_tile_split_share.store(std::max(0, value)); g_tile_history_epoch++; return true; }
    if (!std::strcmp(key, "two_wave_chunks")) { g_two_wave_chunks.store(value > 0 ? min(value, 16) : 1); return true; }
    if (!std::strcmp(key, "vote_min_rays")) { g_vote_min_rays.store(value >= 0 ? value : (768 << 10)); return true; }
    if (!std::strcmp(key, "tile_min_rays")) { g_tile_min_rays.store(value >= 0 ? value : (128 << 10)); return true; }
    if (!std::strcmp(key, "tile_cold_min_rays")) { g_tile_cold_min_rays.store(value >= 0 ? value : (1280 << 10)); return true; }
    return false;
}

void traverse_grid(const Grid& grid, const Tri* tris, const Ray* rays, Hit* hits, int num_rays) {
    dispatch<false>(grid, tris, rays, hits, num_rays);
}

void traverse_grid_prim_ids(const Grid& grid, const Tri* tris, const Ray* rays, Hit* hits, int num_rays) {
    dispatch<true>(grid, tris, rays, hits, num_rays);
}

void trace_two_waves(const Grid& grid, const Tri* tris, int num_tris, const Ray* rays, int num_rays, const int* keys,
                     float offset, float tmax, unsigned seed, Hit* hits_primary, Ray* bounce, Hit* hits_bounce,
                     unsigned long long* counters) {
    if (grid.small_cells) launch_two_waves<SmallCell>(grid, grid.small_cells, tris, num_tris, rays, num_rays, keys, offset, tmax, seed,
                                                      hits_primary, bounce, hits_bounce, counters);
    else                  launch_two_waves<Cell>(grid, grid.cells, tris, num_tris, rays, num_rays, keys, offset, tmax, seed,
                                                 hits_primary, bounce, hits_bounce, counters);
}

void traverse_grid_to_host(const Grid& grid, const Tri* tris, const Ray* dev_rays, Hit* dev_hits, Hit* host_hits, int num_rays,
                           bool prim_ids) {
    if (grid.small_cells) {
        if (prim_ids) launch_to_host<SmallCell, true>(grid, grid.small_cells, tris, dev_rays, dev_hits, host_hits, num_rays);
        else          launch_to_host<SmallCell, false>(grid, grid.small_cells, tris, dev_rays, dev_hits, host_hits, num_rays);
    } else {
        if (prim_ids) launch_to_host<Cell, true>(grid, grid.cells, tris, dev_rays, dev_hits, host_hits, num_rays);
        else          launch_to_host<Cell, false>(grid, grid.cells, tris, dev_rays, dev_hits, host_hits, num_rays);
    }
}

void traverse_grid_host(const Grid& grid, const Tri* tris, const Ray* host_rays, Hit* host_hits, int num_rays,
                        Ray* dev_rays, Hit* dev_hits, bool prim_ids) {
    if (grid.small_cells) {
        if (prim_ids) launch_host_frame<SmallCell, true>(grid, grid.small_cells, tris, host_rays, host_hits, num_rays, dev_rays, dev_hits);
        else          launch_host_frame<SmallCell, false>(grid, grid.small_cells, tris, host_rays, host_hits, num_rays, dev_rays, dev_hits);
    } else {
        if (prim_ids) launch_host_frame<Cell, true>(grid, grid.cells, tris, host_rays, host_hits, num_rays, dev_rays, dev_hits);
        else          launch_host_frame<Cell, false>(grid, grid.cells, tris, host_rays, host_hits, num_rays, dev_rays, dev_hits);
    }
}

/// Diagnosis (tools/gpu_tile_costs.py): the tile times last recorded for the ray buffer (`rays`, `num_rays`), clock
/// ticks / 64 per tile of 32 rays; returns the number of tiles copied (0: nothing recorded for that buffer).
int debug_tile_costs(const void* rays, int num_rays, unsigned short* out, int capacity) {
    DeviceState& st = device_state();
    std::lock_guard<std::mutex> guard(st.lock);
    for (auto& e : st.seen)
        if (e.rays == rays && e.count == num_rays && e.cost && e.history_tiles >= (num_rays + 31) / 32) {
            const int tiles = std::min(capacity, (num_rays + 31) / 32);
            HGB_CUDA(cudaDeviceSynchronize());
            HGB_CUDA(cudaMemcpy(out, e.cost, sizeof(unsigned short) * size_t(tiles), cudaMemcpyDeviceToHost));
            return tiles;
        }
    return 0;
}

} // namespace hagrid

#ifdef HGB_TILE_TRACE
extern "C" __attribute__((visibility("default"))) int hgb_debug_tile_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, hagrid::g_tile_trace, sizeof(long long) * 3 * 8192) == cudaSuccess ? 0 : -1;
}
#endif
