// build_grid for sm_100a: scene box -> top-level grid (Cleary density) ->
// reference emission + exact triangle/cell filtering -> per-top-cell octree
// refinement -> concatenation, stable sort by cell, cell ranges and the octree
// voxel map. Output (entries, cells, ref_ids, offsets, counts) is identical to
// the reference's build (src/build.cu:719-760), including the order of the
// references inside every cell; how it gets there is not:
//
//   * the per-primitive bounding boxes are never stored: three streaming passes
//     recompute them from the 48-byte triangle records;
//   * emission and the exact filter are one kernel (the triangle is loaded once
//     per primitive, large cell ranges are spread over the warp with
//     __ballot_sync/__shfl_sync work sharing);
//   * the reference's two flagged partitions per octree level (kept first,
//     split refs reversed behind them, CUB DevicePartition semantics) are not
//     materialised. One packed 64-bit scan over the references yields both the
//     stable rank of every kept reference and, because the reversed position of
//     split reference i is total - inclusive_prefix(i), the exact slot of each
//     child reference of the next level. Kept references go straight to the
//     level's final array, children straight to the next level;
//   * one host synchronisation per octree level (three totals in one copy)
//     instead of four, all buffers from the stream-ordered pool.
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <vector>

#include "build.h"
#include "device_math.cuh"
#include "primitives.cuh"
#include "runtime.h"
#include "tri_box.cuh"

namespace hagrid {

namespace {

constexpr int kBlock = 128;
constexpr unsigned kAll = 0xFFFFFFFFu;
constexpr int kWarpShareMin = 16;     // cell ranges at least this long are processed by the whole warp

struct BuildParams {
    int   dims_x, dims_y, dims_z;     // top-level dims
    int   shift;                      // log2(virtual / top-level resolution)
    float min_x, min_y, min_z;        // grid box (host-computed)
    float max_x, max_y, max_z;
    float cell_x, cell_y, cell_z;     // virtual cell size (host-computed once `shift` is known)
};

struct CellRange { int lx, ly, lz, hx, hy, hz; };

// ---- order-preserving float <-> uint mapping for the atomic scene-box reduction
__device__ __forceinline__ unsigned ordered(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float from_ordered(unsigned o) {
    const unsigned u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

/// Scene box = union of the triangle boxes (src/build.cu:725-727).
__global__ void __launch_bounds__(kBlock) scene_bounds(const Tri* __restrict__ tris, int n, unsigned* __restrict__ box) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        dev::Float3 a, b;
        dev::tri_bounds(dev::load_tri(tris, i), a, b);
        lo[0] = fminf(lo[0], a.x); lo[1] = fminf(lo[1], a.y); lo[2] = fminf(lo[2], a.z);
        hi[0] = fmaxf(hi[0], b.x); hi[1] = fmaxf(hi[1], b.y); hi[2] = fmaxf(hi[2], b.z);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(kAll, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(kAll, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(box + k, ordered(lo[k]));
            atomicMax(box + 3 + k, ordered(hi[k]));
        }
    }
}

/// Top-level cells overlapped by a box (src/grid.h:84-93): inv = dims / extents is
/// the fast-math quotient (I2F, MUFU.RCP, FMUL), coordinates truncate toward zero.
__device__ __forceinline__ CellRange cell_range(const BuildParams& P, const dev::Float3& lo, const dev::Float3& hi) {
    using namespace dev;
    const float ix = div_approx(int_to_float(P.dims_x), sub(P.max_x, P.min_x));
    const float iy = div_approx(int_to_float(P.dims_y), sub(P.max_y, P.min_y));
    const float iz = div_approx(int_to_float(P.dims_z), sub(P.max_z, P.min_z));
    CellRange r;
    r.lx = max(trunc_to_int(mul(sub(lo.x, P.min_x), ix)), 0);
    r.ly = max(trunc_to_int(mul(sub(lo.y, P.min_y), iy)), 0);
    r.lz = max(trunc_to_int(mul(sub(lo.z, P.min_z), iz)), 0);
    r.hx = min(trunc_to_int(mul(sub(hi.x, P.min_x), ix)), P.dims_x - 1);
    r.hy = min(trunc_to_int(mul(sub(hi.y, P.min_y), iy)), P.dims_y - 1);
    r.hz = min(trunc_to_int(mul(sub(hi.z, P.min_z), iz)), P.dims_z - 1);
    return r;
}

__device__ __forceinline__ int range_size(const CellRange& r) {
    return (r.hx - r.lx + 1) * (r.hy - r.ly + 1) * (r.hz - r.lz + 1);
}

/// k-th cell of a range in emission order (x fastest, src/build.cu:92-103,127-134)
__device__ __forceinline__ int range_cell(const BuildParams& P, int lx, int ly, int lz, int sx, int sy, int k) {
    const int x = lx + k % sx, y = ly + (k / sx) % sy, z = lz + k / (sx * sy);
    return x + P.dims_x * (y + P.dims_y * z);
}

/// Pass 1: references each primitive will emit, and bbox-overlap references per
/// top-level cell (the reference counts those after emission, src/build.cu:57-67,246-253).
__global__ void __launch_bounds__(kBlock) count_refs(const __grid_constant__ BuildParams P, const Tri* __restrict__ tris,
                                                     int n, int* __restrict__ counts, int* __restrict__ refs_per_cell) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    CellRange r = {0, 0, 0, -1, -1, -1};
    int count = 0;
    if (id < n) {
        dev::Float3 lo, hi;
        dev::tri_bounds(dev::load_tri(tris, id), lo, hi);
        r = cell_range(P, lo, hi);
        count = max(0, range_size(r));
        counts[id] = count;
    }
    const int sx = r.hx - r.lx + 1, sy = r.hy - r.ly + 1;
    const bool shared = count >= kWarpShareMin;
    if (!shared)
        for (int k = 0; k < count; k++) atomicAdd(refs_per_cell + range_cell(P, r.lx, r.ly, r.lz, sx, sy, k), 1);
    unsigned todo = __ballot_sync(kAll, shared);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int c = __shfl_sync(kAll, count, src);
        const int lx = __shfl_sync(kAll, r.lx, src), ly = __shfl_sync(kAll, r.ly, src), lz = __shfl_sync(kAll, r.lz, src);
        const int wx = __shfl_sync(kAll, sx, src), wy = __shfl_sync(kAll, sy, src);
        for (int k = lane; k < c; k += 32) atomicAdd(refs_per_cell + range_cell(P, lx, ly, lz, wx, wy, k), 1);
    }
}

/// Octree depth wanted by each top-level cell: ceil(log2(max dim)) of Cleary's
/// resolution for its reference count, in the reference's device arithmetic
/// (src/build.cu:256-270, src/grid.h:96-101): approximate divisions, fast cbrtf.
__global__ void __launch_bounds__(kBlock) top_cell_depths(const __grid_constant__ BuildParams P, const int* __restrict__ refs_per_cell,
                                                          float snd_density, int num_top, int* __restrict__ log_dims,
                                                          int* __restrict__ max_depth) {
    using namespace dev;
    const int id = blockIdx.x * kBlock + threadIdx.x;
    int depth = 0;
    if (id < num_top) {
        const float ex = div_approx(sub(P.max_x, P.min_x), int_to_float(P.dims_x));
        const float ey = div_approx(sub(P.max_y, P.min_y), int_to_float(P.dims_y));
        const float ez = div_approx(sub(P.max_z, P.min_z), int_to_float(P.dims_z));
        const float volume = mul(mul(ex, ey), ez);
        const float ratio = cbrtf(div_approx(mul(snd_density, int_to_float(refs_per_cell[id])), volume));
        const int dx = max(1, trunc_to_int(mul(ex, ratio)));
        const int dy = max(1, trunc_to_int(mul(ey, ratio)));
        const int dz = max(1, trunc_to_int(mul(ez, ratio)));
        const int m = max(dx, max(dy, dz));
        depth = 31 - __clz(m);
        if ((1 << depth) < m) depth++;
        log_dims[id] = depth;
    }
    depth = __reduce_max_sync(kAll, depth);
    if ((threadIdx.x & 31) == 0 && depth > 0) atomicMax(max_depth, depth);
}

/// Top-level cell boxes in virtual-grid units (src/build.cu:332-351)
__global__ void __launch_bounds__(kBlock) emit_top_cells(const __grid_constant__ BuildParams P, Cell* __restrict__ cells, int num_top) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= num_top) return;
    const int x = (id % P.dims_x) << P.shift;
    const int y = ((id / P.dims_x) % P.dims_y) << P.shift;
    const int z = (id / (P.dims_x * P.dims_y)) << P.shift;
    const int inc = 1 << P.shift;
    dev::store_cell(cells, id, x, y, z, 0, x + inc, y + inc, z + inc, 0);
}

__device__ __forceinline__ void cell_world_box(const BuildParams& P, int min_x, int min_y, int min_z, int max_x, int max_y, int max_z,
                                               dev::Float3& lo, dev::Float3& hi) {
    using namespace dev;
    // grid_min + vec3(cell.min) * cell_size:  I2F, FFMA
    lo = {fma(int_to_float(min_x), P.cell_x, P.min_x), fma(int_to_float(min_y), P.cell_y, P.min_y), fma(int_to_float(min_z), P.cell_z, P.min_z)};
    hi = {fma(int_to_float(max_x), P.cell_x, P.min_x), fma(int_to_float(max_y), P.cell_y, P.min_y), fma(int_to_float(max_z), P.cell_z, P.min_z)};
}

/// One (primitive, top cell) reference: exact overlap test against the cell box
/// (filter_refs, src/build.cu:140-158), write it or (-1, -1), and flag the cell
/// for splitting when it keeps a reference and still has depth budget
/// (compute_dims, src/build.cu:286-302).
__device__ __forceinline__ void emit_one(const BuildParams& P, const dev::TriData& tri, int prim, int cell, int slot,
                                         int* __restrict__ ref_ids, int* __restrict__ cell_ids,
                                         const int* __restrict__ log_dims, int* __restrict__ split) {
    const int x = cell % P.dims_x, y = (cell / P.dims_x) % P.dims_y, z = cell / (P.dims_x * P.dims_y);
    dev::Float3 lo, hi;
    cell_world_box(P, x << P.shift, y << P.shift, z << P.shift, (x + 1) << P.shift, (y + 1) << P.shift, (z + 1) << P.shift, lo, hi);
    const bool ok = dev::tri_overlaps_box(tri, lo, hi);
    ref_ids[slot] = ok ? prim : -1;
    cell_ids[slot] = ok ? cell : -1;
    if (ok && log_dims[cell] > 0) split[cell] = 1;
}

/// Pass 2: emission in primitive-major, x-fastest order fused with the exact filter.
__global__ void __launch_bounds__(kBlock) emit_filter_refs(const __grid_constant__ BuildParams P, const Tri* __restrict__ tris, int n,
                                                           const int* __restrict__ start_emit, const int* __restrict__ log_dims,
                                                           int* __restrict__ ref_ids, int* __restrict__ cell_ids, int* __restrict__ split) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    CellRange r = {0, 0, 0, -1, -1, -1};
    dev::TriData tri = {};
    int start = 0, count = 0;
    if (id < n) {
        start = start_emit[id];
        count = start_emit[id + 1] - start;
        if (count > 0) {
            tri = dev::load_tri(tris, id);
            dev::Float3 lo, hi;
            dev::tri_bounds(tri, lo, hi);
            r = cell_range(P, lo, hi);
        }
    }
    const int sx = r.hx - r.lx + 1, sy = r.hy - r.ly + 1;
    const bool shared = count >= kWarpShareMin;
    if (!shared)
        for (int k = 0; k < count; k++)
            emit_one(P, tri, id, range_cell(P, r.lx, r.ly, r.lz, sx, sy, k), start + k, ref_ids, cell_ids, log_dims, split);
    unsigned todo = __ballot_sync(kAll, shared);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        dev::TriData t;
        t.v0 = {__shfl_sync(kAll, tri.v0.x, src), __shfl_sync(kAll, tri.v0.y, src), __shfl_sync(kAll, tri.v0.z, src)};
        t.e1 = {__shfl_sync(kAll, tri.e1.x, src), __shfl_sync(kAll, tri.e1.y, src), __shfl_sync(kAll, tri.e1.z, src)};
        t.e2 = {__shfl_sync(kAll, tri.e2.x, src), __shfl_sync(kAll, tri.e2.y, src), __shfl_sync(kAll, tri.e2.z, src)};
        t.n  = {__shfl_sync(kAll, tri.n.x, src), __shfl_sync(kAll, tri.n.y, src), __shfl_sync(kAll, tri.n.z, src)};
        const int c = __shfl_sync(kAll, count, src), s = __shfl_sync(kAll, start, src), prim = __shfl_sync(kAll, id, src);
        const int lx = __shfl_sync(kAll, r.lx, src), ly = __shfl_sync(kAll, r.ly, src), lz = __shfl_sync(kAll, r.lz, src);
        const int wx = __shfl_sync(kAll, sx, src), wy = __shfl_sync(kAll, sy, src);
        for (int k = lane; k < c; k += 32)
            emit_one(P, t, prim, range_cell(P, lx, ly, lz, wx, wy, k), s + k, ref_ids, cell_ids, log_dims, split);
    }
}

// ------------------------------------------------------------------ octree levels
struct SplitToEight {
    const int* split;
    __device__ __forceinline__ int operator()(int i) const { return split[i] ? 8 : 0; }
};

/// 8-bit child mask of a reference in a cell that splits (compute_split_masks,
/// src/build.cu:160-216): bbox half-space pruning, then the exact test per octant.
__device__ __forceinline__ int child_mask(const BuildParams& P, const dev::TriData& tri, const dev::CellBox& c) {
    using namespace dev;
    Float3 cmin, cmax;
    cell_world_box(P, c.min_x, c.min_y, c.min_z, c.max_x, c.max_y, c.max_z, cmin, cmax);
    const Float3 mid = {mul(add(cmin.x, cmax.x), 0.5f), mul(add(cmin.y, cmax.y), 0.5f), mul(add(cmin.z, cmax.z), 0.5f)};
    Float3 lo, hi;
    tri_bounds(tri, lo, hi);
    int mask = 0xFF;
    if (lo.x > cmax.x || hi.x < cmin.x) mask = 0;
    if (lo.x > mid.x) mask &= 0xAA;
    if (hi.x < mid.x) mask &= 0x55;
    if (lo.y > cmax.y || hi.y < cmin.y) mask = 0;
    if (lo.y > mid.y) mask &= 0xCC;
    if (hi.y < mid.y) mask &= 0x33;
    if (lo.z > cmax.z || hi.z < cmin.z) mask = 0;
    if (lo.z > mid.z) mask &= 0xF0;
    if (hi.z < mid.z) mask &= 0x0F;
    for (int rest = mask; rest; rest &= rest - 1) {
        const int i = __ffs(rest) - 1;
        const Float3 blo = {i & 1 ? mid.x : cmin.x, i & 2 ? mid.y : cmin.y, i & 4 ? mid.z : cmin.z};
        const Float3 bhi = {i & 1 ? cmax.x : mid.x, i & 2 ? cmax.y : mid.y, i & 4 ? cmax.z : mid.z};
        if (!tri_overlaps_box(tri, blo, bhi)) mask &= ~(1 << i);
    }
    return mask;
}

constexpr int kKeptCode = 0x100;

/// Per reference: kept (its cell is a leaf), dropped (filtered out) or the child
/// mask of a splitting cell. code = kKeptCode | 0 | mask.
///
/// HGB_CLASSIFY_STAGED (build-time alternative, tools/gpu_build_variants.py): the north star's "cp.async staging of
/// triangle and cell arrays into shared memory" for this kernel -- every thread requests its triangle (3 x 16 B) and
/// its cell (2 x 16 B) with cp.async.ca into the block's shared memory as soon as it knows the two indices, all
/// threads wait once (cp.async.wait_all), and the separating-axis tests read shared memory. Measured on the 2 M-triangle
/// build (profiles/r02_build_staging.md): no faster than loading into registers -- the gathers are the same dependent
/// 16-byte requests either way, what the kernel waits for is their latency, and the 10 KB of shared memory per block
/// cost resident warps -- so the register form below is what ships.
#ifdef HGB_CLASSIFY_STAGED
__device__ __forceinline__ void stage16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(unsigned(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
#endif

__global__ void __launch_bounds__(kBlock) classify_refs(const __grid_constant__ BuildParams P, const Tri* __restrict__ tris,
                                                        const int* __restrict__ ref_ids, const int* __restrict__ cell_ids,
                                                        const Cell* __restrict__ cells, const int* __restrict__ split,
                                                        int num_refs, int* __restrict__ codes) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
#ifdef HGB_CLASSIFY_STAGED
    __shared__ float4 staged_tri[kBlock][3];
    __shared__ int4 staged_cell[kBlock][2];
    const int cell = id < num_refs ? cell_ids[id] : -1;
    const bool splits = cell >= 0 && split[cell];
    if (splits) {
        const float4* t = reinterpret_cast<const float4*>(tris + ref_ids[id]);
        const int4* c = reinterpret_cast<const int4*>(cells + cell);
        stage16(&staged_tri[threadIdx.x][0], t); stage16(&staged_tri[threadIdx.x][1], t + 1); stage16(&staged_tri[threadIdx.x][2], t + 2);
        stage16(&staged_cell[threadIdx.x][0], c); stage16(&staged_cell[threadIdx.x][1], c + 1);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (id >= num_refs) return;
    int code = 0;
    if (cell >= 0) {
        if (!splits) code = kKeptCode;
        else {
            const float4 a = staged_tri[threadIdx.x][0], b = staged_tri[threadIdx.x][1], c = staged_tri[threadIdx.x][2];
            dev::TriData t;
            t.v0 = {a.x, a.y, a.z}; t.e1 = {b.x, b.y, b.z}; t.e2 = {c.x, c.y, c.z}; t.n = {a.w, b.w, c.w};
            const int4 lo = staged_cell[threadIdx.x][0], hi = staged_cell[threadIdx.x][1];
            dev::CellBox box;
            box.min_x = lo.x; box.min_y = lo.y; box.min_z = lo.z; box.begin = lo.w;
            box.max_x = hi.x; box.max_y = hi.y; box.max_z = hi.z; box.end = hi.w;
            code = child_mask(P, t, box);
        }
    }
    codes[id] = code;
#else
    if (id >= num_refs) return;
    const int cell = cell_ids[id];
    int code = 0;
    if (cell >= 0) {
        if (!split[cell]) code = kKeptCode;
        else code = child_mask(P, dev::load_tri(tris, ref_ids[id]), dev::load_cell_box(cells, cell));
    }
    codes[id] = code;
#endif
}

/// Packed counters of one reference: high word = 1 if kept, low word = number of children
struct CodeCounts {
    const int* codes;
    __device__ __forceinline__ unsigned long long operator()(int i) const {
        const int c = codes[i];
        return c == kKeptCode ? (1ull << 32) : (unsigned long long)__popc(c);
    }
};

/// Moves every reference to its final place: kept ones, in order, to the level's
/// kept arrays; the children of split ones to the next level at the slot the
/// reference's reversed partition would have produced (see file header), child
/// bits ascending (split_refs, src/build.cu:219-243). Flags the child cells that
/// will split again.
__global__ void __launch_bounds__(kBlock) distribute_refs(const __grid_constant__ BuildParams P,
                                                          const int* __restrict__ ref_ids, const int* __restrict__ cell_ids,
                                                          const int* __restrict__ codes, const unsigned long long* __restrict__ pos,
                                                          const int* __restrict__ child_start, const Cell* __restrict__ cells,
                                                          const int* __restrict__ log_dims, int level, int num_refs, int num_children,
                                                          int* __restrict__ kept_refs, int* __restrict__ kept_cells,
                                                          int* __restrict__ next_refs, int* __restrict__ next_cells,
                                                          int* __restrict__ next_split) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= num_refs) return;
    const int code = codes[id];
    if (code == 0) return;
    const unsigned long long p = pos[id];
    const int ref = ref_ids[id], cell = cell_ids[id];
    if (code == kKeptCode) {
        const int slot = int(p >> 32);
        kept_refs[slot] = ref;
        kept_cells[slot] = cell;
        return;
    }
    int slot = num_children - (int(p & 0xFFFFFFFFu) + __popc(code));
    const int first_child = child_start[cell];
    const int4 cmin = dev::ldg4i(cells + cell);
    const int top = (cmin.x >> P.shift) + P.dims_x * ((cmin.y >> P.shift) + P.dims_y * (cmin.z >> P.shift));
    const bool deeper = log_dims[top] > level + 1;
    for (int rest = code; rest; rest &= rest - 1) {
        const int child = first_child + __ffs(rest) - 1;
        next_refs[slot] = ref;
        next_cells[slot] = child;
        if (deeper) next_split[child] = 1;
        slot++;
    }
}

/// Voxel-map word of every cell of the level and the eight children of the cells
/// that split (update_entries + emit_new_cells, src/build.cu:317-383).
__global__ void __launch_bounds__(kBlock) emit_children(const Cell* __restrict__ cells, const int* __restrict__ split,
                                                        const int* __restrict__ child_start, int num_cells,
                                                        uint32_t* __restrict__ entries, Cell* __restrict__ next_cells) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= num_cells) return;
    if (!split[id]) {
        entries[id] = uint32_t(id) << 2;
        return;
    }
    const int first = child_start[id];
    entries[id] = (uint32_t(first) << 2) | 1u;
    const dev::CellBox c = dev::load_cell_box(cells, id);
    const int inc = (c.max_x - c.min_x) >> 1;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int x = c.min_x + (i & 1) * inc, y = c.min_y + ((i >> 1) & 1) * inc, z = c.min_z + (i >> 2) * inc;
        dev::store_cell(next_cells, first + i, x, y, z, 0, x + inc, y + inc, z + inc, 0);
    }
}

// ------------------------------------------------------------------ concatenation
struct IsLeaf {
    const uint32_t* entries;
    __device__ __forceinline__ int operator()(int i) const { return (entries[i] & 3u) == 0; }
};

/// Compacts the leaves of one level into the final cell array and rewrites the
/// level's voxel-map words (copy_cells + copy_entries, src/build.cu:407-440).
__global__ void __launch_bounds__(kBlock) finish_level(const Cell* __restrict__ cells, const int* __restrict__ leaf_index,
                                                       uint32_t* __restrict__ entries, Cell* __restrict__ out_cells,
                                                       int cell_off, int next_level_off, int num_cells) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= num_cells) return;
    const uint32_t e = entries[cell_off + id];
    if ((e & 3u) == 0) {
        const int dst = leaf_index[cell_off + id];
        const dev::CellBox c = dev::load_cell_box(cells, id);
        dev::store_cell(out_cells, dst, c.min_x, c.min_y, c.min_z, c.begin, c.max_x, c.max_y, c.max_z, c.end);
        entries[cell_off + id] = uint32_t(dst) << 2;
    } else {
        entries[cell_off + id] = (((e >> 2) + uint32_t(next_level_off)) << 2) | (e & 3u);
    }
}

/// Final cell index of every kept reference of one level (copy_refs + remap_refs)
__global__ void __launch_bounds__(kBlock) gather_refs(const int* __restrict__ kept_refs, const int* __restrict__ kept_cells,
                                                      const int* __restrict__ leaf_index, int cell_off, int count,
                                                      int* __restrict__ out_refs, int* __restrict__ out_keys) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= count) return;
    out_refs[id] = kept_refs[id];
    out_keys[id] = leaf_index[cell_off + kept_cells[id]];
}

/// begin/end of every non-empty cell from the sorted keys (src/build.cu:453-468)
__global__ void __launch_bounds__(kBlock) cell_ranges(const int* __restrict__ keys, Cell* __restrict__ cells, int num_refs) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= num_refs) return;
    const int cell = keys[id];
    if (id == num_refs - 1) { cells[cell].end = id + 1; return; }
    const int next = keys[id + 1];
    if (cell != next) {
        cells[cell].end = id + 1;
        cells[next].begin = id + 1;
    }
}

inline int blocks_for(int n) { return (n + kBlock - 1) / kBlock; }

struct Level {
    int* kept_refs = nullptr;     // references whose cell is a leaf of this level, in order
    int* kept_cells = nullptr;    // their level-local cell indices (second half of kept_refs' buffer)
    int  num_kept = 0;
    Cell* cells = nullptr;
    int  num_cells = 0;
    uint32_t* entries = nullptr;  // level-local voxel-map words
};

} // namespace

void build_grid(MemManager& mem, const Tri* tris, int num_tris, Grid& grid, float top_density, float snd_density) {
    // ---- scene box (one 24-byte copy back: the top-level resolution is host arithmetic)
    int* totals = mem.alloc<int>(16);
    unsigned* box_bits = reinterpret_cast<unsigned*>(totals) + 8;
    {
        unsigned init[6];
        unsigned lo = 0xFFFFFFFFu, hi = 0u;
        for (int k = 0; k < 3; k++) { init[k] = lo; init[3 + k] = hi; }
        HGB_CUDA(cudaMemcpyAsync(box_bits, init, sizeof(init), cudaMemcpyHostToDevice, 0));
        scene_bounds<<<std::max(1, std::min(blocks_for(num_tris), sm_count() * 8)), kBlock>>>(tris, num_tris, box_bits); count_launch();
    }
    unsigned box_host[6];
    HGB_CUDA(cudaMemcpy(box_host, box_bits, sizeof(box_host), cudaMemcpyDeviceToHost));
    BBox grid_bb(vec3(from_ordered(box_host[0]), from_ordered(box_host[1]), from_ordered(box_host[2])),
                 vec3(from_ordered(box_host[3]), from_ordered(box_host[4]), from_ordered(box_host[5])));

    // Domain extension (the only place this build leaves the reference's arithmetic): a scene box
    // with a zero extent (all triangles in one axis-aligned plane) makes the reference divide by a
    // zero volume (src/grid.h:96-101) and ask for a 31-level octree, i.e. it runs out of memory.
    // Such a box is given a thickness of 0.1 % of its largest extent; boxes with positive extents
    // are untouched, so every input the reference can build is built identically.
    {
        const vec3 e = grid_bb.extents();
        const float widest = std::max(e.x, std::max(e.y, e.z));
        const float pad = (widest > 0.0f ? widest : 1.0f) * 0.0005f;
        if (!(e.x > 0.0f)) { grid_bb.min.x -= pad; grid_bb.max.x += pad; }
        if (!(e.y > 0.0f)) { grid_bb.min.y -= pad; grid_bb.max.y += pad; }
        if (!(e.z > 0.0f)) { grid_bb.min.z -= pad; grid_bb.max.z += pad; }
    }

    // Cleary resolution, dims rounded up to even, box grown by 0.1 % per side (src/build.cu:728-737)
    ivec3 dims = compute_grid_dims(grid_bb, num_tris, top_density);
    dims.x += dims.x & 1; dims.y += dims.y & 1; dims.z += dims.z & 1;
    const vec3 extents = grid_bb.extents();
    grid_bb.min = grid_bb.min - extents * 0.001f;
    grid_bb.max = grid_bb.max + extents * 0.001f;

    BuildParams P;
    P.dims_x = dims.x; P.dims_y = dims.y; P.dims_z = dims.z;
    P.shift = 0;
    P.min_x = grid_bb.min.x; P.min_y = grid_bb.min.y; P.min_z = grid_bb.min.z;
    P.max_x = grid_bb.max.x; P.max_y = grid_bb.max.y; P.max_z = grid_bb.max.z;
    P.cell_x = P.cell_y = P.cell_z = 0.0f;

    const int num_top = dims.x * dims.y * dims.z;

    // ---- pass 1: reference counts per primitive and per top cell, octree depths, emission offsets
    int* start_emit = mem.alloc<int>(size_t(num_tris) + 1);
    int* counts = mem.alloc<int>(size_t(num_tris) + 1);
    int* refs_per_cell = mem.alloc<int>(num_top);
    int* log_dims = mem.alloc<int>(num_top);
    int* scan_tmp = mem.alloc<int>(prim::scan_scratch_elems<int>(num_tris));
    mem.zero(refs_per_cell, num_top);
    mem.zero(totals, 8);
    count_refs<<<blocks_for(num_tris), kBlock>>>(P, tris, num_tris, counts, refs_per_cell); count_launch();
    top_cell_depths<<<blocks_for(num_top), kBlock>>>(P, refs_per_cell, snd_density, num_top, log_dims, totals + 1); count_launch();
    prim::exclusive_scan<int>(prim::LoadInt{counts}, num_tris, start_emit, scan_tmp, totals + 0);
    int host_totals[4];
    HGB_CUDA(cudaMemcpy(host_totals, totals, 2 * sizeof(int), cudaMemcpyDeviceToHost));
    mem.free(counts);
    mem.free(refs_per_cell);
    mem.free(scan_tmp);
    int num_refs = host_totals[0];
    P.shift = host_totals[1];
    {
        const vec3 cell_size = grid_bb.extents() / vec3(dims << P.shift);   // host IEEE (src/build.cu:509)
        P.cell_x = cell_size.x; P.cell_y = cell_size.y; P.cell_z = cell_size.z;
    }

    // ---- level 0: top cells, emission fused with the exact filter
    std::vector<Level> levels;
    int* ref_ids = mem.alloc<int>(2 * size_t(std::max(num_refs, 1)));
    int* cell_ids = ref_ids + num_refs;
    Cell* cells = mem.alloc<Cell>(num_top);
    int* split = mem.alloc<int>(num_top);
    int num_cells = num_top;
    mem.zero(split, num_top);
    emit_top_cells<<<blocks_for(num_top), kBlock>>>(P, cells, num_top); count_launch();
    emit_filter_refs<<<blocks_for(num_tris), kBlock>>>(P, tris, num_tris, start_emit, log_dims, ref_ids, cell_ids, split); count_launch();
    mem.free(start_emit);

    // ---- octree refinement, one level per iteration (build_iter, src/build.cu:528-619)
    for (int level = 0;; level++) {
        Level L;
        L.cells = cells;
        L.num_cells = num_cells;
        L.entries = mem.alloc<uint32_t>(num_cells);

        int* child_start = mem.alloc<int>(size_t(num_cells) + 1);
        int* codes = mem.alloc<int>(std::max(num_refs, 1));
        auto pos = mem.alloc<unsigned long long>(size_t(num_refs) + 1);
        auto scan_tmp64 = mem.alloc<unsigned long long>(prim::scan_scratch_elems<unsigned long long>(std::max(num_refs, num_cells)));
        auto totals64 = reinterpret_cast<unsigned long long*>(totals + 4);

        prim::exclusive_scan<int>(SplitToEight{split}, num_cells, child_start, reinterpret_cast<int*>(scan_tmp64), totals + 2);
        if (num_refs > 0)
            classify_refs<<<blocks_for(num_refs), kBlock>>>(P, tris, ref_ids, cell_ids, cells, split, num_refs, codes); count_launch();
        prim::exclusive_scan<unsigned long long>(CodeCounts{codes}, num_refs, pos, scan_tmp64, totals64);

        struct { int pad[2]; int num_children_cells; int pad2; unsigned long long ref_totals; } host;
        HGB_CUDA(cudaMemcpy(&host, totals, sizeof(host), cudaMemcpyDeviceToHost));
        const int num_new_cells = host.num_children_cells;
        const int num_kept = int(host.ref_totals >> 32);
        const int num_new_refs = int(host.ref_totals & 0xFFFFFFFFu);

        L.num_kept = num_kept;
        L.kept_refs = mem.alloc<int>(2 * size_t(std::max(num_kept, 1)));
        L.kept_cells = L.kept_refs + num_kept;

        int* next_refs = nullptr;
        int* next_cell_ids = nullptr;
        Cell* next_cells = nullptr;
        int* next_split = nullptr;
        if (num_new_cells > 0) {
            next_refs = mem.alloc<int>(2 * size_t(std::max(num_new_refs, 1)));
            next_cell_ids = next_refs + num_new_refs;
            next_cells = mem.alloc<Cell>(num_new_cells);
            next_split = mem.alloc<int>(num_new_cells);
            mem.zero(next_split, num_new_cells);
        }
        if (num_refs > 0)
            distribute_refs<<<blocks_for(num_refs), kBlock>>>(P, ref_ids, cell_ids, codes, pos, child_start, cells, log_dims, level,
                                                              num_refs, num_new_refs, L.kept_refs, L.kept_cells,
                                                              next_refs, next_cell_ids, next_split); count_launch();
        emit_children<<<blocks_for(num_cells), kBlock>>>(cells, split, child_start, num_cells, L.entries, next_cells); count_launch();
        HGB_CUDA(cudaGetLastError());

        mem.free(child_start);
        mem.free(codes);
        mem.free(pos);
        mem.free(scan_tmp64);
        mem.free(ref_ids);
        mem.free(split);
        levels.push_back(L);

        if (num_new_cells == 0) break;
        ref_ids = next_refs;
        cell_ids = next_cell_ids;
        cells = next_cells;
        split = next_split;
        num_refs = num_new_refs;
        num_cells = num_new_cells;
    }
    mem.free(log_dims);

    // ---- concatenate the levels (concat_levels, src/build.cu:621-716)
    const int num_levels = int(levels.size());
    int total_refs = 0, total_cells = 0;
    for (const Level& L : levels) { total_refs += L.num_kept; total_cells += L.num_cells; }

    uint32_t* entries = mem.alloc<uint32_t>(total_cells);
    for (int i = 0, off = 0; i < num_levels; off += levels[i].num_cells, i++) {
        HGB_CUDA(cudaMemcpyAsync(entries + off, levels[i].entries, sizeof(uint32_t) * levels[i].num_cells, cudaMemcpyDeviceToDevice, 0));
        mem.free(levels[i].entries);
    }
    int* leaf_index = mem.alloc<int>(size_t(total_cells) + 1);
    int* leaf_tmp = mem.alloc<int>(prim::scan_scratch_elems<int>(total_cells));
    prim::exclusive_scan<int>(IsLeaf{entries}, total_cells, leaf_index, leaf_tmp, totals + 0);
    int num_leaves = 0;
    HGB_CUDA(cudaMemcpy(&num_leaves, totals, sizeof(int), cudaMemcpyDeviceToHost));
    mem.free(leaf_tmp);

    Cell* out_cells = mem.alloc<Cell>(std::max(num_leaves, 1));
    int* out_refs = mem.alloc<int>(std::max(total_refs, 1));
    int* out_keys = mem.alloc<int>(std::max(total_refs, 1));
    for (int i = 0, cell_off = 0, ref_off = 0; i < num_levels; i++) {
        const Level& L = levels[i];
        finish_level<<<blocks_for(L.num_cells), kBlock>>>(L.cells, leaf_index, entries, out_cells, cell_off, cell_off + L.num_cells, L.num_cells); count_launch();
        if (L.num_kept > 0)
            gather_refs<<<blocks_for(L.num_kept), kBlock>>>(L.kept_refs, L.kept_cells, leaf_index, cell_off, L.num_kept,
                                                            out_refs + ref_off, out_keys + ref_off); count_launch();
        cell_off += L.num_cells;
        ref_off += L.num_kept;
    }
    HGB_CUDA(cudaGetLastError());
    for (Level& L : levels) { mem.free(L.cells); mem.free(L.kept_refs); }
    mem.free(leaf_index);

    // ---- stable sort of the references by final cell, then the cell ranges
    if (total_refs > 0) {
        int* alt_refs = mem.alloc<int>(total_refs);
        int* alt_keys = mem.alloc<int>(total_refs);
        int* sort_tmp = mem.alloc<int>(prim::sort_scratch_ints(total_refs));
        const bool in_alt = prim::sort_pairs(out_keys, out_refs, alt_keys, alt_refs, total_refs, ilog2(num_leaves), sort_tmp);
        if (in_alt) { std::swap(out_refs, alt_refs); std::swap(out_keys, alt_keys); }
        cell_ranges<<<blocks_for(total_refs), kBlock>>>(out_keys, out_cells, total_refs); count_launch();
        HGB_CUDA(cudaGetLastError());
        mem.free(alt_refs);
        mem.free(alt_keys);
        mem.free(sort_tmp);
    }
    mem.free(out_keys);
    mem.free(totals);

    grid.entries = reinterpret_cast<Entry*>(entries);
    grid.ref_ids = out_refs;
    grid.cells = out_cells;
    grid.small_cells = nullptr;
    grid.shift = num_levels - 1;
    grid.num_cells = num_leaves;
    grid.num_entries = total_cells;
    grid.num_refs = total_refs;
    grid.dims = dims;
    grid.bbox = grid_bb;
    grid.offsets.resize(num_levels);
    for (int i = 0, off = 0; i < num_levels; i++) {
        off += levels[i].num_cells;
        grid.offsets[i] = off;
    }
}

} // namespace hagrid
