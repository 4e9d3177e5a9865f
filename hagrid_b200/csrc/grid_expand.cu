// expand_grid for sm_100a (semantics of src/expand.cu:11-225, subset-only mode —
// the only one the reference compiles in, src/expand.cu:159): a cell's box may grow
// across a face by the smallest extent of the cells touching that face, provided
// every one of them references a subset of the cell's own primitives. Cells keep
// their voxel-map ownership; only the boxes overlap, which lets rays skip ahead.
//
// The pass structure is the reference's, on purpose: `iters` x (x, y, z) steps that
// ping-pong between two cell arrays, with cells whose flag bit is clear NOT copied
// to the output array (so they fall back to their state of two steps earlier,
// src/expand.cu:154-155,186-195). Every intermediate state is a valid expansion;
// reproducing the quirk keeps the cell boxes — and the traversal step counts —
// identical to the reference's. Pure integer work, no host synchronisation.
#include <algorithm>

#include "build.h"
#include "device_math.cuh"
#include "runtime.h"

namespace hagrid {

namespace {

constexpr int kBlock = 128;

struct ExpandParams {
    int dims_x, dims_y, dims_z;     // virtual dims
    int top_x, top_y;
    int shift;
};

template <int axis> __device__ __forceinline__ int pick(int x, int y, int z) { return axis == 0 ? x : (axis == 1 ? y : z); }

/// Is list `sub` (length m) contained in list `own` (length n)? One forward walk
/// over `own`, valid for ascending lists and reproduced literally for the others
/// (src/expand.cu:21-36).
__device__ __forceinline__ bool contains_all(const int* __restrict__ own, int n, const int* __restrict__ sub, int m) {
    if (m > n) return false;
    if (m == 0) return true;
    int i = 0, j = 0;
    do {
        const int a = own[i], b = sub[j];
        if (b < a) return false;
        j += a == b;
        i++;
    } while (i < n && j < m);
    return j == m;
}

/// Growth of `cell` across its low (dir = false) or high (dir = true) face on `axis`.
/// The face is scanned in (axis1, axis2) order, stepping by the extents of the
/// neighbours found through the voxel map (src/expand.cu:59-143).
template <int axis, bool dir>
__device__ __forceinline__ int face_growth(const ExpandParams& P, const uint32_t* __restrict__ entries, const int* __restrict__ refs,
                                           const Cell* __restrict__ cells, const dev::CellBox& cell, bool& keep_going) {
    constexpr int axis1 = (axis + 1) % 3, axis2 = (axis + 2) % 3;
    const int lo = pick<axis>(cell.min_x, cell.min_y, cell.min_z), hi = pick<axis>(cell.max_x, cell.max_y, cell.max_z);
    const int size = pick<axis>(P.dims_x, P.dims_y, P.dims_z);
    if (dir ? hi >= size : lo <= 0) return 0;

    const int lo1 = pick<axis1>(cell.min_x, cell.min_y, cell.min_z), hi1 = pick<axis1>(cell.max_x, cell.max_y, cell.max_z);
    const int lo2 = pick<axis2>(cell.min_x, cell.min_y, cell.min_z), hi2 = pick<axis2>(cell.max_x, cell.max_y, cell.max_z);
    const int size2 = pick<axis2>(P.dims_x, P.dims_y, P.dims_z);
    const int face = dir ? hi : lo - 1;
    int d = dir ? size : -size;
    int limit = d;
    int step2 = size2;
    int i = lo1, j = lo2;
    while (true) {
        int vx, vy, vz;
        if (axis == 0) { vx = face; vy = i; vz = j; }
        if (axis == 1) { vx = j; vy = face; vz = i; }
        if (axis == 2) { vx = i; vy = j; vz = face; }
        const dev::CellBox next = dev::load_cell_box(cells, dev::lookup_cell(entries, P.shift, P.top_x, P.top_y, vx, vy, vz));
        if (dir) {
            limit = min(limit, pick<axis>(next.max_x, next.max_y, next.max_z) - hi);
            d = min(d, limit);
        } else {
            limit = max(limit, pick<axis>(next.min_x, next.min_y, next.min_z) - lo);
            d = max(d, limit);
        }
        if (!contains_all(refs + cell.begin, cell.end - cell.begin, refs + next.begin, next.end - next.begin)) {
            d = 0;
            break;
        }
        const int step1 = pick<axis1>(next.max_x, next.max_y, next.max_z) - i;
        step2 = min(step2, pick<axis2>(next.max_x, next.max_y, next.max_z) - j);
        i += step1;
        if (i >= hi1) {
            i = lo1;
            j += step2;
            step2 = size2;
            if (j >= hi2) break;
        }
    }
    keep_going |= d == limit;
    return d;
}

template <int axis>
__global__ void __launch_bounds__(kBlock) grow_cells(const __grid_constant__ ExpandParams P, const uint32_t* __restrict__ entries,
                                                     const int* __restrict__ refs, const Cell* __restrict__ cells,
                                                     Cell* __restrict__ out_cells, int* __restrict__ flags, int num_cells) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= num_cells) return;
    const int flag = flags[id];
    if ((flag & (1 << axis)) == 0) return;        // not copied either: see the file header
    dev::CellBox c = dev::load_cell_box(cells, id);
    bool keep_going = false;
    const int low = face_growth<axis, false>(P, entries, refs, cells, c, keep_going);
    const int high = face_growth<axis, true>(P, entries, refs, cells, c, keep_going);
    if (axis == 0) { c.min_x += low; c.max_x += high; }
    if (axis == 1) { c.min_y += low; c.max_y += high; }
    if (axis == 2) { c.min_z += low; c.max_z += high; }
    flags[id] = (keep_going ? 1 << axis : 0) | (flag & ~(1 << axis));
    dev::store_cell(out_cells, id, c.min_x, c.min_y, c.min_z, c.begin, c.max_x, c.max_y, c.max_z, c.end);
}

/// The same step with two lanes per cell, one per face of the axis. The two face scans are independent (both
/// start from the same box) and the scan of a large face is a long chain of dependent voxel-map look-ups, so
/// splitting them halves the critical path of the heavy cells that decide the kernel's duration. Pays off while
/// the kernel is latency-bound (C2, 234 K cells: -9 %); with millions of cells it is bandwidth-bound and the
/// doubled cell loads cost more than the shorter chains save (C4, 7.1 M cells: +49 %), so expand_grid() picks.
template <int axis>
__global__ void __launch_bounds__(kBlock) grow_cells_paired(const __grid_constant__ ExpandParams P, const uint32_t* __restrict__ entries,
                                                            const int* __restrict__ refs, const Cell* __restrict__ cells,
                                                            Cell* __restrict__ out_cells, int* __restrict__ flags, int num_cells) {
    constexpr unsigned kAll = 0xFFFFFFFFu;
    const int thread = blockIdx.x * kBlock + threadIdx.x;
    const int id = thread >> 1;
    const bool high_face = thread & 1;
    int flag = 0;
    bool active = false;
    if (id < num_cells) {
        flag = flags[id];
        active = (flag & (1 << axis)) != 0;
    }
    dev::CellBox c = {};
    bool keep_going = false;
    int growth = 0;
    if (active) {
        c = dev::load_cell_box(cells, id);
        growth = high_face ? face_growth<axis, true>(P, entries, refs, cells, c, keep_going)
                           : face_growth<axis, false>(P, entries, refs, cells, c, keep_going);
    }
    // the low-face lane (even) collects the high-face lane's result and writes the cell
    const int other_growth = __shfl_xor_sync(kAll, growth, 1);
    const bool other_keep = __shfl_xor_sync(kAll, int(keep_going), 1) != 0;
    if (!active || high_face) return;
    const int low = growth, high = other_growth;
    keep_going |= other_keep;
    if (axis == 0) { c.min_x += low; c.max_x += high; }
    if (axis == 1) { c.min_y += low; c.max_y += high; }
    if (axis == 2) { c.min_z += low; c.max_z += high; }
    flags[id] = (keep_going ? 1 << axis : 0) | (flag & ~(1 << axis));
    dev::store_cell(out_cells, id, c.min_x, c.min_y, c.min_z, c.begin, c.max_x, c.max_y, c.max_z, c.end);
}

template <int axis>
void grow_step(const ExpandParams& P, const uint32_t* entries, const Grid& grid, Cell* out, int* flags) {
    constexpr int kPairedBelow = 1 << 20;
    if (grid.num_cells < kPairedBelow)
        grow_cells_paired<axis><<<(2 * grid.num_cells + kBlock - 1) / kBlock, kBlock>>>(P, entries, grid.ref_ids, grid.cells, out, flags, grid.num_cells);
    else
        grow_cells<axis><<<(grid.num_cells + kBlock - 1) / kBlock, kBlock>>>(P, entries, grid.ref_ids, grid.cells, out, flags, grid.num_cells);
    count_launch();
}

} // namespace

void expand_grid(MemManager& mem, Grid& grid, const Tri*, int iters) {
    if (iters == 0) return;
    Cell* other = mem.alloc<Cell>(std::max(grid.num_cells, 1));
    int* flags = mem.alloc<int>(std::max(grid.num_cells, 1));
    mem.one(flags, grid.num_cells);

    const ivec3 dims = grid.dims << grid.shift;
    ExpandParams P;
    P.dims_x = dims.x; P.dims_y = dims.y; P.dims_z = dims.z;
    P.top_x = grid.dims.x; P.top_y = grid.dims.y;
    P.shift = grid.shift;

    auto entries = reinterpret_cast<const uint32_t*>(grid.entries);
    for (int i = 0; i < iters && grid.num_cells > 0; i++) {
        grow_step<0>(P, entries, grid, other, flags);
        std::swap(other, grid.cells);
        grow_step<1>(P, entries, grid, other, flags);
        std::swap(other, grid.cells);
        grow_step<2>(P, entries, grid, other, flags);
        std::swap(other, grid.cells);
    }
    HGB_CUDA(cudaGetLastError());
    mem.free(flags);
    mem.free(other);
}

} // namespace hagrid
