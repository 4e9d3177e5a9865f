// compress_grid for sm_100a (semantics of src/compress.cu:6-63): 32-byte cells
// become 16-byte cells with 16-bit coordinates, and every non-empty reference list
// is re-emitted with a -1 terminator so that the cell only stores its first index.
// Lists are copied by the owning lane when short and by the whole warp when long.
#include <algorithm>

#include "build.h"
#include "device_math.cuh"
#include "primitives.cuh"
#include "runtime.h"

namespace hagrid {

namespace {

constexpr int kBlock = 128;
constexpr unsigned kAll = 0xFFFFFFFFu;

/// Words a cell's list occupies once terminated: n + 1, or nothing when empty
struct SentinelCount {
    const Cell* cells;
    __device__ __forceinline__ int operator()(int i) const {
        const int4 a = dev::ldg4i(cells + i);
        const int4 b = dev::ldg4i(reinterpret_cast<const int4*>(cells + i) + 1);
        const int n = b.w - a.w;
        return n > 0 ? n + 1 : 0;
    }
};

__global__ void __launch_bounds__(kBlock) emit_small_cells(const Cell* __restrict__ cells, const int* __restrict__ refs,
                                                           const int* __restrict__ list_start, SmallCell* __restrict__ small_cells,
                                                           int* __restrict__ out_refs, int num_cells) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int src = 0, dst = 0, n = 0;
    if (id < num_cells) {
        const dev::CellBox c = dev::load_cell_box(cells, id);
        n = c.end - c.begin;
        src = c.begin;
        dst = list_start[id];
        // {min.x | min.y << 16, min.z | max.x << 16, max.y | max.z << 16, begin} (src/grid.h:170-176)
        const uint4 packed = make_uint4((uint32_t(c.min_x) & 0xFFFFu) | (uint32_t(c.min_y) << 16),
                                        (uint32_t(c.min_z) & 0xFFFFu) | (uint32_t(c.max_x) << 16),
                                        (uint32_t(c.max_y) & 0xFFFFu) | (uint32_t(c.max_z) << 16),
                                        uint32_t(n > 0 ? dst : -1));
        *reinterpret_cast<uint4*>(small_cells + id) = packed;
        if (n > 0) out_refs[dst + n] = -1;
    }
    constexpr int kShare = 24;
    if (n > 0 && n < kShare)
        for (int k = 0; k < n; k++) out_refs[dst + k] = refs[src + k];
    unsigned todo = __ballot_sync(kAll, n >= kShare);
    while (todo) {
        const int from = __ffs(todo) - 1;
        todo &= todo - 1;
        const int s = __shfl_sync(kAll, src, from), d = __shfl_sync(kAll, dst, from), m = __shfl_sync(kAll, n, from);
        for (int k = lane; k < m; k += 32) out_refs[d + k] = refs[s + k];
    }
}

} // namespace

bool compress_grid(MemManager& mem, Grid& grid) {
    const ivec3 dims = grid.dims << grid.shift;
    if (dims.x >= (1 << 16) || dims.y >= (1 << 16) || dims.z >= (1 << 16)) return false;

    const int num_cells = grid.num_cells;
    int* list_start = mem.alloc<int>(size_t(num_cells) + 1);
    int* scan_tmp = mem.alloc<int>(prim::scan_scratch_elems<int>(num_cells) + 1);
    int* total_dev = scan_tmp + prim::scan_scratch_elems<int>(num_cells);
    SmallCell* small_cells = mem.alloc<SmallCell>(std::max(num_cells, 1));
    prim::exclusive_scan<int>(SentinelCount{grid.cells}, num_cells, list_start, scan_tmp, total_dev);
    int num_words = 0;
    HGB_CUDA(cudaMemcpy(&num_words, total_dev, sizeof(int), cudaMemcpyDeviceToHost));
    int* out_refs = mem.alloc<int>(std::max(num_words, 1));
    if (num_cells > 0)
        emit_small_cells<<<(num_cells + kBlock - 1) / kBlock, kBlock>>>(grid.cells, grid.ref_ids, list_start, small_cells, out_refs, num_cells); count_launch();
    HGB_CUDA(cudaGetLastError());

    grid.small_cells = small_cells;
    mem.free(grid.cells);
    mem.free(grid.ref_ids);
    mem.free(list_start);
    mem.free(scan_tmp);
    grid.cells = nullptr;
    grid.ref_ids = out_refs;
    grid.num_refs = num_words;
    return true;
}

} // namespace hagrid
