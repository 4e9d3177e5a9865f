// merge_grid for sm_100a: SAH-driven merging of face-aligned neighbour cells
// along x, y, z, repeated while the cell count keeps shrinking by `alpha`
// (semantics of src/merge.cu:292-377, output identical cell by cell).
//
// Per axis pass:
//   1. pair_up<axis>     every cell looks up the cell behind its far face; if the
//                        two boxes are face-aligned, merging is allowed at this
//                        position and the surface-area heuristic says so, it
//                        records the size of the merged reference list and links
//                        itself to that neighbour (next/prev).
//   2. resolve_chains    chain heads walk their chain: members at even positions
//                        survive (and absorb their successor), odd ones vanish.
//                        The survivor's new reference count is written in the
//                        same walk (the reference uses a second kernel).
//   3. one packed 64-bit scan gives each survivor its new cell index and its
//      reference offset; both totals come back in one copy.
//   4. merge_cells<axis> writes the merged boxes and reference lists. A warp
//      handles 32 consecutive cells; list copies and the two-way merges are
//      done by the owning lane for short lists and spread over the warp for
//      long ones.
//   5. remap_entries     leaf words of the voxel map follow their cells.
#include <algorithm>

#include "build.h"
#include "device_math.cuh"
#include "primitives.cuh"
#include "runtime.h"

namespace hagrid {

namespace {

constexpr int kBlock = 128;
constexpr unsigned kAll = 0xFFFFFFFFu;

struct MergeParams {
    int   dims_x, dims_y, dims_z;     // virtual dims
    int   top_x, top_y;               // top-level dims
    int   shift;
    float cell_x, cell_y, cell_z;     // host-computed virtual cell size
};

template <int axis> __device__ __forceinline__ int pick(int x, int y, int z) { return axis == 0 ? x : (axis == 1 ? y : z); }

/// Cell count of the grid a pass works on. The counts of a pass (cells in the high word, references in the
/// low word: the totals of its scan) stay on the device and feed the next pass directly; the host only reads
/// them once per round of three passes, for the termination test, and launches every pass of the round over
/// the cell count it knew at the round's start (cells can only get fewer).
__device__ __forceinline__ int live_cells(const unsigned long long* __restrict__ live) { return int(__ldg(live) >> 32); }

/// |A u B| counted by the reference's two-pointer walk (src/merge.cu:58-69). The
/// lists are only sorted for cells of even octree depth; the walk is reproduced
/// literally because its result on unsorted input decides merges too.
__device__ __forceinline__ int union_size(const int* __restrict__ p0, int c0, const int* __restrict__ p1, int c1) {
    int i = 0, j = 0, c = 0;
    while (i < c0 && j < c1) {
        const int a = p0[i], b = p1[j];
        i += a <= b;
        j += a >= b;
        c++;
    }
    return c + (c1 - j) + (c0 - i);
}

/// Two-pointer merge matching union_size (src/merge.cu:72-88)
__device__ __forceinline__ void merge_lists(const int* __restrict__ p0, int c0, const int* __restrict__ p1, int c1,
                                            int* __restrict__ q) {
    int i = 0, j = 0;
    while (i < c0 && j < c1) {
        const int a = p0[i], b = p1[j];
        *q++ = a < b ? a : b;
        i += a <= b;
        j += a >= b;
    }
    for (; i < c0; i++) *q++ = p0[i];
    for (; j < c1; j++) *q++ = p1[j];
}

/// Step 1. merge_counts[id] >= 0: size of the merged list, the cell wants to absorb
/// nexts[id]; otherwise -(own count + 1). (compute_merge_counts, src/merge.cu:92-143)
///
/// SAH with unit traversal cost, half-areas in world units. The rounding of every
/// product follows the reference's SASS, which differs per axis because the two
/// aligned boxes share two extents and the compiler reuses those products:
///   axis 0:  A_i = fma(ex_i, ey + ez, rn(ey * ez))      A = (A1 + A2) - rn(ey * ez)
///   axis 1:  A_i = fma(ey_i, ez, rn(ex * (ey_i + ez)))  A = fma(-ez, ex, A1 + A2)
///   axis 2:  A_i = fma(ey, ez_i, rn(ex * (ey + ez_i)))  A = fma(-ex, ey, A1 + A2)
///   cost of not merging = fma(A1, n1 + 1, rn(A2 * (n2 + 1)))
template <int axis>
__global__ void __launch_bounds__(kBlock) pair_up(const __grid_constant__ MergeParams P, const uint32_t* __restrict__ entries,
                                                  const Cell* __restrict__ cells, const int* __restrict__ refs,
                                                  int* __restrict__ merge_counts, int* __restrict__ nexts, int* __restrict__ prevs,
                                                  int empty_mask, const unsigned long long* __restrict__ live) {
    using namespace dev;
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= live_cells(live)) return;
    const CellBox c1 = load_cell_box(cells, id);
    const int n1 = c1.end - c1.begin;
    int count = -(n1 + 1);
    int next_id = -1;

    // merging is restricted at top-level cell boundaries during the first rounds (src/merge.cu:34-39)
    const int pos = pick<axis>(c1.min_x, c1.min_y, c1.min_z);
    const bool shifted = ((pos >> P.shift) & empty_mask) != 0;
    const bool on_top_boundary = (pos & ((1 << P.shift) - 1)) == 0;
    const int far = pick<axis>(c1.max_x, c1.max_y, c1.max_z);
    if ((!shifted || !on_top_boundary) && far < pick<axis>(P.dims_x, P.dims_y, P.dims_z)) {
        next_id = lookup_cell(entries, P.shift, P.top_x, P.top_y,
                              axis == 0 ? c1.max_x : c1.min_x, axis == 1 ? c1.max_y : c1.min_y, axis == 2 ? c1.max_z : c1.min_z);
        const CellBox c2 = load_cell_box(cells, next_id);
        bool aligned;
        if (axis == 0) aligned = c1.max_x == c2.min_x && c1.min_y == c2.min_y && c1.min_z == c2.min_z && c1.max_y == c2.max_y && c1.max_z == c2.max_z;
        if (axis == 1) aligned = c1.max_y == c2.min_y && c1.min_z == c2.min_z && c1.min_x == c2.min_x && c1.max_z == c2.max_z && c1.max_x == c2.max_x;
        if (axis == 2) aligned = c1.max_z == c2.min_z && c1.min_x == c2.min_x && c1.min_y == c2.min_y && c1.max_x == c2.max_x && c1.max_y == c2.max_y;
        if (aligned) {
            const int n2 = c2.end - c2.begin;
            float a1, a2, a;
            if (axis == 0) {
                const float ey = mul(int_to_float(c1.max_y - c1.min_y), P.cell_y), ez = mul(int_to_float(c1.max_z - c1.min_z), P.cell_z);
                const float ex1 = mul(int_to_float(c1.max_x - c1.min_x), P.cell_x), ex2 = mul(int_to_float(c2.max_x - c2.min_x), P.cell_x);
                const float s = add(ey, ez), p = mul(ey, ez);
                a1 = fma(ex1, s, p); a2 = fma(ex2, s, p);
                a = sub(add(a1, a2), p);
            } else if (axis == 1) {
                const float ex = mul(int_to_float(c1.max_x - c1.min_x), P.cell_x), ez = mul(int_to_float(c1.max_z - c1.min_z), P.cell_z);
                const float ey1 = mul(int_to_float(c1.max_y - c1.min_y), P.cell_y), ey2 = mul(int_to_float(c2.max_y - c2.min_y), P.cell_y);
                a1 = fma(ey1, ez, mul(ex, add(ey1, ez))); a2 = fma(ey2, ez, mul(ex, add(ey2, ez)));
                a = fma(-ez, ex, add(a1, a2));
            } else {
                const float ex = mul(int_to_float(c1.max_x - c1.min_x), P.cell_x), ey = mul(int_to_float(c1.max_y - c1.min_y), P.cell_y);
                const float ez1 = mul(int_to_float(c1.max_z - c1.min_z), P.cell_z), ez2 = mul(int_to_float(c2.max_z - c2.min_z), P.cell_z);
                a1 = fma(ey, ez1, mul(ex, add(ey, ez1))); a2 = fma(ey, ez2, mul(ex, add(ey, ez2)));
                a = fma(-ex, ey, add(a1, a2));
            }
            const float apart = fma(a1, add(int_to_float(n1), 1.0f), mul(a2, add(int_to_float(n2), 1.0f)));
            // the union holds at least max(n1, n2) references: cheap rejection first
            if (mul(a, add(int_to_float(max(n1, n2)), 1.0f)) <= apart) {
                const int n = union_size(refs + c1.begin, n1, refs + c2.begin, n2);
                if (mul(a, add(int_to_float(n), 1.0f)) <= apart) count = n;
            }
        }
    }
    merge_counts[id] = count;
    next_id = count >= 0 ? next_id : -1;
    nexts[id] = next_id;
    if (next_id >= 0) prevs[next_id] = id;
}

/// Step 2. kept[id] = 1 for survivors; new_counts[id] = their reference count after
/// the merge, 0 for absorbed cells (compute_cell_flags + compute_ref_counts,
/// src/merge.cu:146-187).
__global__ void __launch_bounds__(kBlock) resolve_chains(const int* __restrict__ nexts, const int* __restrict__ prevs,
                                                         const int* __restrict__ merge_counts, int* __restrict__ kept,
                                                         int* __restrict__ new_counts, const unsigned long long* __restrict__ live) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= live_cells(live) || prevs[id] >= 0) return;
    int cur = id;
    bool keep = true;
    while (cur >= 0) {
        const int m = merge_counts[cur];
        kept[cur] = keep;
        new_counts[cur] = keep ? (m >= 0 ? m : -(m + 1)) : 0;
        cur = nexts[cur];
        keep = !keep;
    }
}

/// Scanned over the host's upper bound of the cell count: slots past the live count contribute nothing
struct KeptAndCount {
    const int* kept;
    const int* new_counts;
    const unsigned long long* live;
    __device__ __forceinline__ unsigned long long operator()(int i) const {
        if (i >= live_cells(live)) return 0ull;
        return ((unsigned long long)kept[i] << 32) | (unsigned)new_counts[i];
    }
};

/// Step 4 (merge, src/merge.cu:190-278).
template <int axis>
__global__ void __launch_bounds__(kBlock) merge_cells(const __grid_constant__ MergeParams P, const uint32_t* __restrict__ entries,
                                                      const Cell* __restrict__ cells, const int* __restrict__ refs,
                                                      const unsigned long long* __restrict__ scan, const int* __restrict__ merge_counts,
                                                      int* __restrict__ new_cell_ids, Cell* __restrict__ new_cells,
                                                      int* __restrict__ new_refs, const unsigned long long* __restrict__ live) {
    using namespace dev;
    const int id = blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int src0 = 0, n0 = 0, src1 = 0, n1 = 0, dst = 0;     // list(s) this lane has to write
    bool two_way = false;
    if (id < live_cells(live)) {
        const unsigned long long here = scan[id], after = scan[id + 1];
        const int new_id = int(here >> 32);
        if (int(after >> 32) > new_id) {                  // survivor
            const CellBox c = load_cell_box(cells, id);
            const int merged = merge_counts[id];
            dst = int(here & 0xFFFFFFFFu);
            src0 = c.begin; n0 = c.end - c.begin;
            new_cell_ids[id] = new_id;
            if (merged >= 0) {
                const int next_id = lookup_cell(entries, P.shift, P.top_x, P.top_y,
                                                axis == 0 ? c.max_x : c.min_x, axis == 1 ? c.max_y : c.min_y, axis == 2 ? c.max_z : c.min_z);
                const CellBox d = load_cell_box(cells, next_id);
                new_cell_ids[next_id] = new_id;
                src1 = d.begin; n1 = d.end - d.begin;
                two_way = n1 > 0;                         // an empty partner degenerates to a copy
                store_cell(new_cells, new_id, min(c.min_x, d.min_x), min(c.min_y, d.min_y), min(c.min_z, d.min_z), dst,
                           max(c.max_x, d.max_x), max(c.max_y, d.max_y), max(c.max_z, d.max_z), dst + merged);
            } else {
                store_cell(new_cells, new_id, c.min_x, c.min_y, c.min_z, dst, c.max_x, c.max_y, c.max_z, dst + n0);
            }
        }
    }
    if (two_way) {
        merge_lists(refs + src0, n0, refs + src1, n1, new_refs + dst);
        n0 = 0;
    }
    // plain copies: short ones by the owner, long ones by the whole warp
    constexpr int kShare = 24;
    if (n0 > 0 && n0 < kShare)
        for (int k = 0; k < n0; k++) new_refs[dst + k] = refs[src0 + k];
    unsigned todo = __ballot_sync(kAll, n0 >= kShare);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int s = __shfl_sync(kAll, src0, src), d = __shfl_sync(kAll, dst, src), n = __shfl_sync(kAll, n0, src);
        for (int k = lane; k < n; k += 32) new_refs[d + k] = refs[s + k];
    }
}

/// Step 5 (remap_entries, src/merge.cu:281-290)
__global__ void __launch_bounds__(kBlock) remap_entries(uint32_t* __restrict__ entries, const int* __restrict__ new_cell_ids, int num_entries) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= num_entries) return;
    const uint32_t e = entries[id];
    if ((e & 3u) == 0) entries[id] = uint32_t(new_cell_ids[e >> 2]) << 2;
}

inline int blocks_for(int n) { return (n + kBlock - 1) / kBlock; }

struct MergeBuffers {
    int* merge_counts;
    int* nexts;
    int* prevs;
    int* kept;
    int* new_counts;
    int* new_cell_ids;
    unsigned long long* scan;
    unsigned long long* scan_tmp;
    unsigned long long* totals;
};

template <int axis>
void merge_pass(const MergeParams& P, Grid& grid, Cell*& spare_cells, int*& spare_refs, int empty_mask, MergeBuffers& b,
                int bound, int pass) {
    auto entries = reinterpret_cast<uint32_t*>(grid.entries);
    const unsigned long long* live = b.totals + (pass & 1);          // counts the previous pass left
    unsigned long long* next = b.totals + ((pass + 1) & 1);
    HGB_CUDA(cudaMemsetAsync(b.prevs, 0xFF, sizeof(int) * size_t(bound), 0));
    pair_up<axis><<<blocks_for(bound), kBlock>>>(P, entries, grid.cells, grid.ref_ids, b.merge_counts, b.nexts, b.prevs, empty_mask, live); count_launch();
    resolve_chains<<<blocks_for(bound), kBlock>>>(b.nexts, b.prevs, b.merge_counts, b.kept, b.new_counts, live); count_launch();
    prim::exclusive_scan<unsigned long long>(KeptAndCount{b.kept, b.new_counts, live}, bound, b.scan, b.scan_tmp, next);
    merge_cells<axis><<<blocks_for(bound), kBlock>>>(P, entries, grid.cells, grid.ref_ids, b.scan, b.merge_counts, b.new_cell_ids,
                                                     spare_cells, spare_refs, live); count_launch();
    remap_entries<<<blocks_for(grid.num_entries), kBlock>>>(entries, b.new_cell_ids, grid.num_entries); count_launch();
    HGB_CUDA(cudaGetLastError());
    std::swap(spare_cells, grid.cells);
    std::swap(spare_refs, grid.ref_ids);
}

} // namespace

void merge_grid(MemManager& mem, Grid& grid, float alpha) {
    // The ping-pong buffers exist even when alpha <= 0, like in the reference, so
    // that the allocator traffic (and therefore the caller's peak) is comparable.
    Cell* spare_cells = mem.alloc<Cell>(std::max(grid.num_cells, 1));
    int* spare_refs = mem.alloc<int>(std::max(grid.num_refs, 1));

    const size_t n = size_t(grid.num_cells) + 1;
    MergeBuffers b;
    b.merge_counts = mem.alloc<int>(n);
    b.nexts = mem.alloc<int>(n);
    b.prevs = mem.alloc<int>(n);
    b.kept = mem.alloc<int>(n);
    b.new_counts = mem.alloc<int>(n);
    b.new_cell_ids = mem.alloc<int>(n);
    b.scan = mem.alloc<unsigned long long>(n);
    b.scan_tmp = mem.alloc<unsigned long long>(prim::scan_scratch_elems<unsigned long long>(grid.num_cells) + 2);
    b.totals = b.scan_tmp + prim::scan_scratch_elems<unsigned long long>(grid.num_cells);       // two slots, used alternately

    const vec3 extents = grid.bbox.extents();
    const ivec3 dims = grid.dims << grid.shift;
    const vec3 cell_size = extents / vec3(dims);          // host IEEE (src/merge.cu:349-351)
    MergeParams P;
    P.dims_x = dims.x; P.dims_y = dims.y; P.dims_z = dims.z;
    P.top_x = dims.x >> grid.shift; P.top_y = dims.y >> grid.shift;
    P.shift = grid.shift;
    P.cell_x = cell_size.x; P.cell_y = cell_size.y; P.cell_z = cell_size.z;

    if (alpha > 0 && grid.num_cells > 0) {
        unsigned long long totals = ((unsigned long long)grid.num_cells << 32) | (unsigned)grid.num_refs;
        HGB_CUDA(cudaMemcpy(b.totals, &totals, sizeof(totals), cudaMemcpyHostToDevice));
        int before, round = 0, pass = 0;
        do {
            before = grid.num_cells;
            const int mask = round > 3 ? 0 : (1 << (round + 1)) - 1;
            merge_pass<0>(P, grid, spare_cells, spare_refs, mask, b, before, pass++);
            merge_pass<1>(P, grid, spare_cells, spare_refs, mask, b, before, pass++);
            merge_pass<2>(P, grid, spare_cells, spare_refs, mask, b, before, pass++);
            HGB_CUDA(cudaMemcpy(&totals, b.totals + (pass & 1), sizeof(totals), cudaMemcpyDeviceToHost));     // the round's only sync
            grid.num_cells = int(totals >> 32);
            grid.num_refs = int(totals & 0xFFFFFFFFu);
            round++;
        } while (grid.num_cells < alpha * before);
    }

    mem.free(b.merge_counts);
    mem.free(b.nexts);
    mem.free(b.prevs);
    mem.free(b.kept);
    mem.free(b.new_counts);
    mem.free(b.new_cell_ids);
    mem.free(b.scan);
    mem.free(b.scan_tmp);
    mem.free(spare_cells);
    mem.free(spare_refs);
}

} // namespace hagrid
