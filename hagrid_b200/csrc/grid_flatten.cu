// flatten_grid for sm_100a (semantics of src/flatten.cu:9-175): octree nodes whose
// eight children are identical collapse into their child, then up to three octree
// levels are fused into one dense (2^d)^3 voxel-map node, d = min(subtree depth, 3).
//
// Differences in execution, not in result: collapse and subtree depth are one
// kernel per level (bottom-up); the dense nodes are written one thread per output
// word, each finding its source entry by binary search in the scanned node sizes
// (the reference launches one 64-thread block per source entry, most of which
// exit immediately); one host synchronisation per group of three levels.
#include <algorithm>
#include <vector>

#include "build.h"
#include "device_math.cuh"
#include "primitives.cuh"
#include "runtime.h"

namespace hagrid {

namespace {

constexpr int kBlock = 128;
constexpr unsigned kAll = 0xFFFFFFFFu;
constexpr int kFlatLevels = (1 << Entry::LOG_DIM_BITS) - 1;   // 3

/// Bottom-up, one level per launch: collapse a node whose eight child words are
/// equal, then depth = 1 + max(child depths). Child blocks are 32-byte aligned
/// because the top-level dims are even (src/build.cu:730-733).
__global__ void __launch_bounds__(kBlock) collapse_and_depth(uint32_t* __restrict__ entries, int* __restrict__ depths, int first, int count) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= count) return;
    uint32_t e = entries[first + id];
    int depth = 0;
    if (e & 3u) {
        const uint4* kids = reinterpret_cast<const uint4*>(entries + (e >> 2));
        const uint4 a = kids[0], b = kids[1];
        if (a.x == a.y && a.x == a.z && a.x == a.w && a.x == b.x && a.x == b.y && a.x == b.z && a.x == b.w) {
            e = a.x;
            entries[first + id] = e;
        }
        if (e & 3u) {
            const int4* kd = reinterpret_cast<const int4*>(depths + (e >> 2));
            const int4 p = kd[0], q = kd[1];
            depth = 1 + max(max(max(p.x, q.x), max(p.y, q.y)), max(max(p.z, q.z), max(p.w, q.w)));
        }
    }
    depths[first + id] = depth;
}

/// Number of voxel-map words the flattened node of an entry occupies
struct FlatSize {
    const int* depths;
    __device__ __forceinline__ int operator()(int i) const {
        const int d = depths[i];
        return d > 0 ? 1 << (min(d, kFlatLevels) * 3) : 0;
    }
};

/// Top-level words: inner ones point at their flattened node (src/flatten.cu:49-62)
__global__ void __launch_bounds__(kBlock) copy_top(const uint32_t* __restrict__ entries, const int* __restrict__ node_start,
                                                   const int* __restrict__ depths, uint32_t* __restrict__ out, int count) {
    const int id = blockIdx.x * kBlock + threadIdx.x;
    if (id >= count) return;
    uint32_t e = entries[id];
    if (e & 3u) e = (uint32_t(count + node_start[id]) << 2) | uint32_t(min(depths[id], kFlatLevels));
    out[id] = e;
}

/// Dense nodes of one group of levels (flatten_level, src/flatten.cu:65-107): sub-entry i of a node is read as d
/// octal digits, most significant first; each digit picks a child while the walk is still on an inner word, and
/// sets one bit of x, y, z.
/// One thread per OUTPUT word: the thread finds the source entry whose node contains its word by
/// a binary search in the exclusive scan of the node sizes (the last entry that starts at or before the word;
/// leaves have size 0 and share their successor's start), then walks the d digits. Every thread has the same
/// amount of work, whereas a block (the reference) or a warp per group of source entries writes anything
/// between nothing and 16 384 words.
__global__ void __launch_bounds__(kBlock) write_nodes_by_word(const uint32_t* __restrict__ entries, const int* __restrict__ node_start,
                                                              const int* __restrict__ depths, uint32_t* __restrict__ out,
                                                              int first, int offset, int next_offset, int count, int total_words) {
    const int word = blockIdx.x * kBlock + threadIdx.x;
    if (word >= total_words) return;
    int lo = 0, hi = count;                                   // first index whose start exceeds `word`
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(node_start + first + mid) > word) hi = mid; else lo = mid + 1;
    }
    const int id = first + lo - 1;
    const int i = word - __ldg(node_start + id);
    const int d = min(__ldg(depths + id), kFlatLevels);
    uint32_t e = __ldg(entries + id);
    int x = 0, y = 0, z = 0, at = id;
    for (int level = d - 1; level >= 0; level--) {
        const int digit = (i >> (3 * level)) & 7;
        x |= (digit & 1) << level;
        y |= ((digit >> 1) & 1) << level;
        z |= (digit >> 2) << level;
        if (e & 3u) {
            at = int(e >> 2) + digit;
            e = __ldg(entries + at);
        }
    }
    if (e & 3u) e = (uint32_t(next_offset + __ldg(node_start + at)) << 2) | uint32_t(min(__ldg(depths + at), kFlatLevels));
    out[offset + (word - i) + x + ((y + (z << d)) << d)] = e;
}

inline int blocks_for(int n) { return (n + kBlock - 1) / kBlock; }

} // namespace

void flatten_grid(MemManager& mem, Grid& grid) {
    auto entries = reinterpret_cast<uint32_t*>(grid.entries);
    int* depths = mem.alloc<int>(size_t(grid.num_entries) + 1);

    for (int level = grid.shift; level >= 0; level--) {
        const int first = level > 0 ? grid.offsets[level - 1] : 0;
        const int count = grid.offsets[level] - first;
        if (count > 0) collapse_and_depth<<<blocks_for(count), kBlock>>>(entries, depths, first, count); count_launch();
    }

    // where each flattened node starts inside its group, and the size of every group
    int* node_start = mem.alloc<int>(size_t(grid.num_entries) + 1);
    int* scan_tmp = mem.alloc<int>(prim::scan_scratch_elems<int>(grid.num_entries) + 1);
    int* total_dev = scan_tmp + prim::scan_scratch_elems<int>(grid.num_entries);
    std::vector<int> group_offset(std::max(grid.shift, 1), 0), group_words(std::max(grid.shift, 1), 0);
    int total_entries = grid.offsets[0];
    for (int level = 0; level < grid.shift; level += kFlatLevels) {
        const int first = level > 0 ? grid.offsets[level - 1] : 0;
        const int count = grid.offsets[level] - first;
        prim::exclusive_scan<int>(FlatSize{depths + first}, count, node_start + first, scan_tmp, total_dev);
        int group_entries = 0;
        HGB_CUDA(cudaMemcpy(&group_entries, total_dev, sizeof(int), cudaMemcpyDeviceToHost));
        group_offset[level] = total_entries;
        group_words[level] = group_entries;
        total_entries += group_entries;
    }

    uint32_t* out = mem.alloc<uint32_t>(total_entries);
    std::vector<int> new_offsets;
    copy_top<<<blocks_for(grid.offsets[0]), kBlock>>>(entries, node_start, depths, out, grid.offsets[0]); count_launch();
    for (int level = 0; level < grid.shift; level += kFlatLevels) {
        const int first = level > 0 ? grid.offsets[level - 1] : 0;
        const int count = grid.offsets[level] - first;
        const int next_offset = level + kFlatLevels < grid.shift ? group_offset[level + kFlatLevels] : 0;
        if (count > 0 && group_words[level] > 0)
            write_nodes_by_word<<<blocks_for(group_words[level]), kBlock>>>(entries, node_start, depths, out, first, group_offset[level],
                                                                           next_offset, count, group_words[level]); count_launch();
        new_offsets.push_back(group_offset[level]);
    }
    new_offsets.push_back(total_entries);
    HGB_CUDA(cudaGetLastError());

    mem.free(grid.entries);
    grid.entries = reinterpret_cast<Entry*>(out);
    grid.offsets = new_offsets;
    grid.num_entries = total_entries;
    mem.free(depths);
    mem.free(node_start);
    mem.free(scan_tmp);
}

} // namespace hagrid
