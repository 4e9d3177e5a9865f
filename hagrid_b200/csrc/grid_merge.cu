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
#include <atomic>
#include <cstring>
#include <cooperative_groups.h>

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

// Loads for the two ways the passes run. As separate launches every input of a kernel is read-only for that kernel and
// travels the non-coherent path (__ldg / const __restrict__). Inside the one-launch round (merge_round below) the same
// arrays are written by one phase and read by the next, with a grid-wide barrier in between: that needs ordinary,
// coherent loads (kCoherent), which the barrier's fence orders behind the other blocks' stores.
template <bool kCoherent> __device__ __forceinline__ int4 load_i4(const void* p) {
    if (!kCoherent) return dev::ldg4i(p);
    int4 v;
    asm volatile("ld.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
template <bool kCoherent, typename T> __device__ __forceinline__ T load_w(const T* p) {
    static_assert(sizeof(T) == 4, "32-bit words");
    if (!kCoherent) return *p;
    unsigned v;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return T(v);
}
template <bool kCoherent> __device__ __forceinline__ unsigned long long load_ll(const unsigned long long* p) {
    if (!kCoherent) return *p;
    unsigned long long v;
    asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
template <bool kCoherent> __device__ __forceinline__ dev::CellBox load_box(const Cell* cells, int id) {
    if (!kCoherent) return dev::load_cell_box(cells, id);
    const int4 a = load_i4<true>(cells + id);
    const int4 b = load_i4<true>(reinterpret_cast<const int4*>(cells + id) + 1);
    dev::CellBox c;
    c.min_x = a.x; c.min_y = a.y; c.min_z = a.z; c.begin = a.w;
    c.max_x = b.x; c.max_y = b.y; c.max_z = b.z; c.end = b.w;
    return c;
}
/// dev::lookup_cell with the loads above
template <bool kCoherent> __device__ __forceinline__ int find_cell(const uint32_t* entries, int shift, int top_x, int top_y, int vx, int vy, int vz) {
    if (!kCoherent) return dev::lookup_cell(entries, shift, top_x, top_y, vx, vy, vz);
    uint32_t e = load_w<true>(entries + ((vx >> shift) + top_x * ((vy >> shift) + top_y * (vz >> shift))));
    uint32_t log_dim = e & 3u;
    int depth = int(log_dim);
    while (log_dim) {
        const int s = shift - depth;
        const uint32_t mask = (1u << log_dim) - 1u;
        const uint32_t kx = (uint32_t(vx) >> s) & mask, ky = (uint32_t(vy) >> s) & mask, kz = (uint32_t(vz) >> s) & mask;
        e = load_w<true>(entries + ((e >> 2) + kx + ((ky + (kz << log_dim)) << log_dim)));
        log_dim = e & 3u;
        depth += int(log_dim);
    }
    return int(e >> 2);
}

/// Cell count of the grid a pass works on. The counts of a pass (cells in the high word, references in the
/// low word: the totals of its scan) stay on the device and feed the next pass directly; the host only reads
/// them once per round of three passes, for the termination test, and launches every pass of the round over
/// the cell count it knew at the round's start (cells can only get fewer).
__device__ __forceinline__ int live_cells(const unsigned long long* __restrict__ live) { return int(__ldg(live) >> 32); }

/// |A u B| counted by the reference's two-pointer walk (src/merge.cu:58-69). The
/// lists are only sorted for cells of even octree depth; the walk is reproduced
/// literally because its result on unsorted input decides merges too.
template <bool kCoherent = false>
__device__ __forceinline__ int union_size(const int* __restrict__ p0, int c0, const int* __restrict__ p1, int c1) {
    int i = 0, j = 0, c = 0;
    while (i < c0 && j < c1) {
        const int a = load_w<kCoherent>(p0 + i), b = load_w<kCoherent>(p1 + j);
        i += a <= b;
        j += a >= b;
        c++;
    }
    return c + (c1 - j) + (c0 - i);
}

/// Two-pointer merge matching union_size (src/merge.cu:72-88)
template <bool kCoherent = false>
__device__ __forceinline__ void merge_lists(const int* __restrict__ p0, int c0, const int* __restrict__ p1, int c1,
                                            int* __restrict__ q) {
    int i = 0, j = 0;
    while (i < c0 && j < c1) {
        const int a = load_w<kCoherent>(p0 + i), b = load_w<kCoherent>(p1 + j);
        *q++ = a < b ? a : b;
        i += a <= b;
        j += a >= b;
    }
    for (; i < c0; i++) *q++ = load_w<kCoherent>(p0 + i);
    for (; j < c1; j++) *q++ = load_w<kCoherent>(p1 + j);
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
template <int axis, bool kCoherent = false>
__device__ __forceinline__ void pair_up_cell(int id, const MergeParams& P, const uint32_t* __restrict__ entries,
                                             const Cell* __restrict__ cells, const int* __restrict__ refs,
                                             int* __restrict__ merge_counts, int* __restrict__ nexts, int* __restrict__ prevs,
                                             int empty_mask) {
    using namespace dev;
    const CellBox c1 = load_box<kCoherent>(cells, id);
    const int n1 = c1.end - c1.begin;
    int count = -(n1 + 1);
    int next_id = -1;

    // merging is restricted at top-level cell boundaries during the first rounds (src/merge.cu:34-39)
    const int pos = pick<axis>(c1.min_x, c1.min_y, c1.min_z);
    const bool shifted = ((pos >> P.shift) & empty_mask) != 0;
    const bool on_top_boundary = (pos & ((1 << P.shift) - 1)) == 0;
    const int far = pick<axis>(c1.max_x, c1.max_y, c1.max_z);
    if ((!shifted || !on_top_boundary) && far < pick<axis>(P.dims_x, P.dims_y, P.dims_z)) {
        next_id = find_cell<kCoherent>(entries, P.shift, P.top_x, P.top_y,
                                       axis == 0 ? c1.max_x : c1.min_x, axis == 1 ? c1.max_y : c1.min_y, axis == 2 ? c1.max_z : c1.min_z);
        const CellBox c2 = load_box<kCoherent>(cells, next_id);
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
                const int n = union_size<kCoherent>(refs + c1.begin, n1, refs + c2.begin, n2);
                if (mul(a, add(int_to_float(n), 1.0f)) <= apart) count = n;
            }
        }
    }
    merge_counts[id] = count;
    next_id = count >= 0 ? next_id : -1;
    nexts[id] = next_id;
    if (next_id >= 0) prevs[next_id] = id;
}

template <int axis>
__global__ void __launch_bounds__(kBlock) pair_up(const __grid_constant__ MergeParams P, const uint32_t* __restrict__ entries,
                                                  const Cell* __restrict__ cells, const int* __restrict__ refs,
                                                  int* __restrict__ merge_counts, int* __restrict__ nexts, int* __restrict__ prevs,
                                                  int empty_mask, const unsigned long long* __restrict__ live) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= live_cells(live)) return;
    pair_up_cell<axis>(id, P, entries, cells, refs, merge_counts, nexts, prevs, empty_mask);
}

/// Step 2. kept[id] = 1 for survivors; new_counts[id] = their reference count after
/// the merge, 0 for absorbed cells (compute_cell_flags + compute_ref_counts,
/// src/merge.cu:146-187).
template <bool kCoherent = false>
__device__ __forceinline__ void resolve_chain(int id, const int* __restrict__ nexts, const int* __restrict__ prevs,
                                              const int* __restrict__ merge_counts, int* __restrict__ kept, int* __restrict__ new_counts) {
    if (load_w<kCoherent>(prevs + id) >= 0) return;
    int cur = id;
    bool keep = true;
    while (cur >= 0) {
        const int m = load_w<kCoherent>(merge_counts + cur);
        kept[cur] = keep;
        new_counts[cur] = keep ? (m >= 0 ? m : -(m + 1)) : 0;
        cur = load_w<kCoherent>(nexts + cur);
        keep = !keep;
    }
}

__global__ void __launch_bounds__(kBlock) resolve_chains(const int* __restrict__ nexts, const int* __restrict__ prevs,
                                                         const int* __restrict__ merge_counts, int* __restrict__ kept,
                                                         int* __restrict__ new_counts, const unsigned long long* __restrict__ live) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= live_cells(live)) return;
    resolve_chain(id, nexts, prevs, merge_counts, kept, new_counts);
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
/// (called by whole warps: `id` may lie past `num_live`)
template <int axis, bool kCoherent = false>
__device__ __forceinline__ void merge_cell(int id, int num_live, const MergeParams& P, const uint32_t* __restrict__ entries,
                                           const Cell* __restrict__ cells, const int* __restrict__ refs,
                                           const unsigned long long* __restrict__ scan, const int* __restrict__ merge_counts,
                                           int* __restrict__ new_cell_ids, Cell* __restrict__ new_cells, int* __restrict__ new_refs) {
    using namespace dev;
    const int lane = threadIdx.x & 31;
    int src0 = 0, n0 = 0, src1 = 0, n1 = 0, dst = 0;     // list(s) this lane has to write
    bool two_way = false;
    if (id < num_live) {
        const unsigned long long here = load_ll<kCoherent>(scan + id), after = load_ll<kCoherent>(scan + id + 1);
        const int new_id = int(here >> 32);
        if (int(after >> 32) > new_id) {                  // survivor
            const CellBox c = load_box<kCoherent>(cells, id);
            const int merged = load_w<kCoherent>(merge_counts + id);
            dst = int(here & 0xFFFFFFFFu);
            src0 = c.begin; n0 = c.end - c.begin;
            new_cell_ids[id] = new_id;
            if (merged >= 0) {
                const int next_id = find_cell<kCoherent>(entries, P.shift, P.top_x, P.top_y,
                                                         axis == 0 ? c.max_x : c.min_x, axis == 1 ? c.max_y : c.min_y, axis == 2 ? c.max_z : c.min_z);
                const CellBox d = load_box<kCoherent>(cells, next_id);
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
        merge_lists<kCoherent>(refs + src0, n0, refs + src1, n1, new_refs + dst);
        n0 = 0;
    }
    // plain copies: short ones by the owner, long ones by the whole warp
    constexpr int kShare = 24;
    if (n0 > 0 && n0 < kShare)
        for (int k = 0; k < n0; k++) new_refs[dst + k] = load_w<kCoherent>(refs + src0 + k);
    unsigned todo = __ballot_sync(kAll, n0 >= kShare);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int s = __shfl_sync(kAll, src0, src), d = __shfl_sync(kAll, dst, src), n = __shfl_sync(kAll, n0, src);
        for (int k = lane; k < n; k += 32) new_refs[d + k] = load_w<kCoherent>(refs + s + k);
    }
}

template <int axis>
__global__ void __launch_bounds__(kBlock) merge_cells(const __grid_constant__ MergeParams P, const uint32_t* __restrict__ entries,
                                                      const Cell* __restrict__ cells, const int* __restrict__ refs,
                                                      const unsigned long long* __restrict__ scan, const int* __restrict__ merge_counts,
                                                      int* __restrict__ new_cell_ids, Cell* __restrict__ new_cells,
                                                      int* __restrict__ new_refs, const unsigned long long* __restrict__ live) {
    merge_cell<axis>(blockIdx.x * kBlock + threadIdx.x, live_cells(live), P, entries, cells, refs, scan, merge_counts, new_cell_ids, new_cells, new_refs);
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

// ---------------------------------------------------------------------------
// The whole of merge_grid in ONE launch, for grids whose passes are too short to be worth a launch each. On the C2
// scene a pass works on 234 000 cells and its five kernels take 6-10 us apiece, most of it the launch and the wait
// for the slowest thread's chain of dependent loads: 18 passes = 108 stream operations and six host round trips for
// 1.1 ms. Here resident blocks (one of 1 024 threads per SM: the barrier costs per block; 128 x 8 per SM is 0.3 ms
// slower on C2) run the same per-cell functions phase by phase with a grid-wide barrier in between
// (cooperative launch), pass after pass and round after round; the termination test of src/merge.cu:371-376 is
// evaluated by every block from the totals the scan leaves in device memory, so the host reads back once, at the end.
// The scan is the plain two-barrier kind (every block sums its contiguous segment, then scans it behind the sums of
// the blocks before it): with at most a few thousand resident blocks the look-back machinery of primitives.cuh buys
// nothing. Same functions, same order of stores per array: the grid is byte-identical to the pass-per-launch path.
// ---------------------------------------------------------------------------
namespace cg = cooperative_groups;

#ifndef HGB_ROUND_BLOCK
#define HGB_ROUND_BLOCK 1024
#endif
#ifndef HGB_ROUND_BLOCKS_PER_SM
#define HGB_ROUND_BLOCKS_PER_SM 1
#endif
constexpr int kRoundBlock = HGB_ROUND_BLOCK;          // few, large blocks: a grid-wide barrier costs per block

struct RoundArgs {
    MergeParams P;
    uint32_t* entries;
    int num_entries;
    Cell* cells[2];                     // [pass & 1] is read, the other written
    int* refs[2];
    int *merge_counts, *nexts, *prevs, *kept, *new_counts, *new_cell_ids;
    unsigned long long *scan, *block_sums, *totals;      // totals[pass & 1]: (cells << 32 | refs) the pass starts from
    int* passes_done;
    float alpha;
};

__device__ __forceinline__ unsigned long long block_total(unsigned long long v, unsigned long long* shared) {
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kAll, v, d);
    if ((threadIdx.x & 31) == 0) shared[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long sum = 0;
    for (int w = 0; w < kRoundBlock / 32; w++) sum += shared[w];
    __syncthreads();
    return sum;
}

/// Exclusive prefix of v over the block's threads; `total` = the block's sum
__device__ __forceinline__ unsigned long long block_exclusive(unsigned long long v, unsigned long long* shared, unsigned long long& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = v;
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long up = __shfl_up_sync(kAll, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) shared[warp] = incl;
    __syncthreads();
    unsigned long long before = 0, sum = 0;
    for (int w = 0; w < kRoundBlock / 32; w++) {
        const unsigned long long t = shared[w];
        if (w < warp) before += t;
        sum += t;
    }
    __syncthreads();
    total = sum;
    return before + incl - v;
}

template <int axis>
__device__ __forceinline__ void round_pass(cg::grid_group& grid, const RoundArgs& A, int pass, int empty_mask, unsigned long long* shared) {
    const Cell* cells = A.cells[pass & 1];         Cell* new_cells = A.cells[(pass + 1) & 1];
    const int* refs = A.refs[pass & 1];            int* new_refs = A.refs[(pass + 1) & 1];
    const int n = int(load_ll<true>(A.totals + (pass & 1)) >> 32);
    const int stride = gridDim.x * kRoundBlock, first = blockIdx.x * kRoundBlock + threadIdx.x;

    for (int id = first; id < n; id += stride)
        pair_up_cell<axis, true>(id, A.P, A.entries, cells, refs, A.merge_counts, A.nexts, A.prevs, empty_mask);
    grid.sync();
    for (int id = first; id < n; id += stride)
        resolve_chain<true>(id, A.nexts, A.prevs, A.merge_counts, A.kept, A.new_counts);
    grid.sync();

    // packed scan of (kept, new_counts) over [0, n): scan[i] = sum before i, scan[n] = totals of the next pass
    const int segment = ((n + int(gridDim.x) - 1) / int(gridDim.x) + kRoundBlock - 1) / kRoundBlock * kRoundBlock;
    const int seg_begin = min(n, int(blockIdx.x) * segment), seg_end = min(n, seg_begin + segment);
    {
        unsigned long long sum = 0;
        for (int i = seg_begin + threadIdx.x; i < seg_end; i += kRoundBlock)
            sum += ((unsigned long long)load_w<true>(A.kept + i) << 32) | unsigned(load_w<true>(A.new_counts + i));
        sum = block_total(sum, shared);
        if (threadIdx.x == 0) A.block_sums[blockIdx.x] = sum;
    }
    grid.sync();
    {
        unsigned long long carry = 0;
        for (int b = threadIdx.x; b < int(blockIdx.x); b += kRoundBlock) carry += load_ll<true>(A.block_sums + b);
        carry = block_total(carry, shared);
        for (int base = seg_begin; base < seg_end; base += kRoundBlock) {
            const int i = base + threadIdx.x;
            const unsigned long long v = i < seg_end ? ((unsigned long long)load_w<true>(A.kept + i) << 32) | unsigned(load_w<true>(A.new_counts + i)) : 0ull;
            unsigned long long chunk;
            const unsigned long long before = block_exclusive(v, shared, chunk);
            if (i < seg_end) A.scan[i] = carry + before;
            carry += chunk;
        }
        // the block whose segment ends the array (block 0 if there is nothing) leaves the totals
        const bool last = n == 0 ? blockIdx.x == 0 : (seg_begin < n && seg_end == n);
        if (last && threadIdx.x == 0) { A.scan[n] = carry; A.totals[(pass + 1) & 1] = carry; }
    }
    grid.sync();
    for (int base = blockIdx.x * kRoundBlock; base < n; base += stride)
        merge_cell<axis, true>(base + threadIdx.x, n, A.P, A.entries, cells, refs, A.scan, A.merge_counts, A.new_cell_ids, new_cells, new_refs);
    grid.sync();
    // voxel map follows the cells; the links of the next pass start out empty
    for (int id = first; id < A.num_entries; id += stride) {
        const uint32_t e = load_w<true>(A.entries + id);
        if ((e & 3u) == 0) A.entries[id] = uint32_t(load_w<true>(A.new_cell_ids + (e >> 2))) << 2;
    }
    const int n_next = int(load_ll<true>(A.totals + ((pass + 1) & 1)) >> 32);
    for (int id = first; id < n_next; id += stride) A.prevs[id] = -1;
    grid.sync();
}

constexpr int kRoundBlocksPerSm = HGB_ROUND_BLOCKS_PER_SM;

__global__ void __launch_bounds__(kRoundBlock, kRoundBlocksPerSm) merge_rounds(const __grid_constant__ RoundArgs A) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned long long shared[kRoundBlock / 32];
    {
        const int n = int(load_ll<true>(A.totals) >> 32);
        for (int id = blockIdx.x * kRoundBlock + threadIdx.x; id < n; id += gridDim.x * kRoundBlock) A.prevs[id] = -1;
        grid.sync();
    }
    int pass = 0, round = 0;
    while (true) {
        const int before = int(load_ll<true>(A.totals + (pass & 1)) >> 32);
        const int mask = round > 3 ? 0 : (1 << (round + 1)) - 1;
        round_pass<0>(grid, A, pass++, mask, shared);
        round_pass<1>(grid, A, pass++, mask, shared);
        round_pass<2>(grid, A, pass++, mask, shared);
        const int now = int(load_ll<true>(A.totals + (pass & 1)) >> 32);
        round++;
        // `grid.num_cells < alpha * before` of the host loop: int -> float, one rounded product, float compare
        if (!(__int2float_rn(now) < __fmul_rn(A.alpha, __int2float_rn(before)))) break;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *A.passes_done = pass;
}

// grids of up to this many cells (before merging) are merged in one launch (0: never). tools/gpu_merge_threshold.py, whole
// builds: 46 K cells 1.14 -> 1.01 ms, 167 K 1.24 -> 1.10, 521 K 1.49 -> 1.41, C2 (551 K) 2.70 -> 2.59; 1.29 M 2.10 -> 2.19,
// 3.1 M 3.44 -> 4.16, C4 (7.1 M after merging) 12.7 -> 19.0, C5 34 -> 55 -- big grids want every SM full of threads in each
// phase, not 1 024 per SM and a barrier.
std::atomic<int> g_one_launch_max_cells{768 << 10};

} // namespace

bool set_merge_option(const char* key, int value) {
    if (std::strcmp(key, "merge_one_launch_max_cells") != 0) return false;
    g_one_launch_max_cells.store(value >= 0 ? value : (768 << 10));
    return true;
}

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
    // (+ the block sums of the one-launch path in front, + three words behind: two totals used alternately, the pass count)
    const size_t scan_words = std::max<size_t>(prim::scan_scratch_elems<unsigned long long>(grid.num_cells), size_t(sm_count()) * kRoundBlocksPerSm);
    b.scan_tmp = mem.alloc<unsigned long long>(scan_words + 3);
    b.totals = b.scan_tmp + scan_words;

    const vec3 extents = grid.bbox.extents();
    const ivec3 dims = grid.dims << grid.shift;
    const vec3 cell_size = extents / vec3(dims);          // host IEEE (src/merge.cu:349-351)
    MergeParams P;
    P.dims_x = dims.x; P.dims_y = dims.y; P.dims_z = dims.z;
    P.top_x = dims.x >> grid.shift; P.top_y = dims.y >> grid.shift;
    P.shift = grid.shift;
    P.cell_x = cell_size.x; P.cell_y = cell_size.y; P.cell_z = cell_size.z;

    // blocks of merge_rounds that fit an SM (the same on every device of one kind); 0: no cooperative launches here
    static std::atomic<int> resident{-1};
    int blocks_per_sm = resident.load();
    if (blocks_per_sm < 0) {
        int dev = 0, cooperative = 0;
        HGB_CUDA(cudaGetDevice(&dev));
        HGB_CUDA(cudaDeviceGetAttribute(&cooperative, cudaDevAttrCooperativeLaunch, dev));
        HGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, merge_rounds, kRoundBlock, 0));
        blocks_per_sm = cooperative ? std::max(0, std::min(blocks_per_sm, kRoundBlocksPerSm)) : 0;
        resident.store(blocks_per_sm);
    }
    if (alpha > 0 && grid.num_cells > 0 && grid.num_cells <= g_one_launch_max_cells.load() && blocks_per_sm > 0) {
        const int blocks = std::min(sm_count() * blocks_per_sm, (grid.num_cells + kRoundBlock - 1) / kRoundBlock);
        // block sums and the pass count share the scan scratch (its look-back words are not used on this path)
        RoundArgs A;
        A.P = P; A.entries = reinterpret_cast<uint32_t*>(grid.entries); A.num_entries = grid.num_entries;
        A.cells[0] = grid.cells; A.cells[1] = spare_cells; A.refs[0] = grid.ref_ids; A.refs[1] = spare_refs;
        A.merge_counts = b.merge_counts; A.nexts = b.nexts; A.prevs = b.prevs; A.kept = b.kept; A.new_counts = b.new_counts;
        A.new_cell_ids = b.new_cell_ids; A.scan = b.scan; A.block_sums = b.scan_tmp; A.totals = b.totals;
        A.passes_done = reinterpret_cast<int*>(b.totals + 2);
        A.alpha = alpha;
        unsigned long long init[3] = {((unsigned long long)grid.num_cells << 32) | (unsigned)grid.num_refs, 0, 0};
        HGB_CUDA(cudaMemcpyAsync(b.totals, init, sizeof(init), cudaMemcpyHostToDevice, 0));
        void* args[] = {&A};
        HGB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(merge_rounds), dim3(blocks), dim3(kRoundBlock), args, 0, 0)); count_launch();
        unsigned long long out[3];
        HGB_CUDA(cudaMemcpy(out, b.totals, sizeof(out), cudaMemcpyDeviceToHost));                             // the only sync
        const int passes = int(out[2] & 0xFFFFFFFFull);
        grid.num_cells = int(out[passes & 1] >> 32);
        grid.num_refs = int(out[passes & 1] & 0xFFFFFFFFu);
        if (passes & 1) { std::swap(spare_cells, grid.cells); std::swap(spare_refs, grid.ref_ids); }
    } else if (alpha > 0 && grid.num_cells > 0) {
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
