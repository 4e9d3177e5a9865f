// Device-wide primitives of the construction pipeline, hand-written for
// sm_100a (the reference wraps CUB's DeviceScan / DeviceReduce /
// DevicePartition / DeviceRadixSort, src/parallel.cuh:12-89).
//
//   * exclusive_scan<T>(): reduce-then-scan over 2048-element tiles with a
//     fused input functor, two launches (the block that finishes the reduce
//     last also scans the tile sums); T = int or a packed 64-bit pair of counters. The
//     total stays on the device (no host round trip per call, unlike
//     parallel.cuh:40); callers fetch several totals with one copy.
//   * sort_pairs(): stable LSD radix sort, 8 bits per pass, warp-match
//     ranking (__match_any_sync) so equal keys keep their input order — the
//     property the build relies on for the reference's reference order.
//   * the reference's flagged partition (kept first, rejected reversed at the
//     rear, CUB semantics) is not materialised at all: the build derives both
//     target positions from one scan (see grid_build.cu).
//
// All kernels are plain grid launches on the legacy default stream; none of
// them spins on another block, so a bug cannot hang the device.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "runtime.h"

namespace hagrid {
namespace prim {

constexpr int kThreads = 256;
constexpr int kItems   = 8;                  // consecutive items per thread
constexpr int kTile    = kThreads * kItems;  // 2048

constexpr unsigned kFullMask = 0xFFFFFFFFu;

inline int num_tiles(int n) { return (n + kTile - 1) / kTile; }

template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(kFullMask, v, o);
        if (lane >= o) v += up;
    }
    return v;
}

/// Exclusive prefix of `v` over the block (kThreads threads); `total` = block sum.
/// `smem` needs kThreads / 32 + 1 elements; safe to call repeatedly.
template <typename T>
__device__ __forceinline__ T block_exclusive_sum(T v, T* smem, T& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T incl = warp_inclusive_sum(v);
    __syncthreads();                                   // protect smem from the previous call
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = lane < kThreads / 32 ? smem[lane] : T(0);
        const T wi = warp_inclusive_sum(w);
        if (lane < kThreads / 32) smem[lane] = wi - w;
        if (lane == kThreads / 32 - 1) smem[kThreads / 32] = wi;
    }
    __syncthreads();
    total = smem[kThreads / 32];
    return smem[warp] + incl - v;
}

// ------------------------------------------------------------------ scan
/// Blocks of scan_reduce_tiles take a ticket when their tile sum is visible; the block that draws the last
/// one scans the tile sums (nobody waits for anybody: the last block simply arrives last) and puts the
/// ticket counter back to zero for the next scan on the stream.
static __device__ unsigned int g_scan_tickets;

/// In-place exclusive scan of `sums[0, count)` by the calling block; writes the grand total.
template <typename T>
__device__ __forceinline__ void block_scan_tile_sums(T* sums, int count, T* total_out, T* smem) {
    T carry = 0;
    for (int start = 0; start < count; start += kTile) {
        const int base = start + threadIdx.x * kItems;
        T v[kItems];
        T sum = 0;
#pragma unroll
        for (int k = 0; k < kItems; k++) {
            v[k] = base + k < count ? __ldcg(sums + base + k) : T(0);      // written by other blocks of this launch
            sum += v[k];
        }
        T total;
        T run = carry + block_exclusive_sum(sum, smem, total);
#pragma unroll
        for (int k = 0; k < kItems; k++) {
            if (base + k < count) sums[base + k] = run;
            run += v[k];
        }
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename T, typename F>
__global__ void __launch_bounds__(kThreads) scan_reduce_tiles(F f, int n, T* __restrict__ tile_sums, T* __restrict__ total_out) {
    __shared__ T smem[kThreads / 32 + 1];
    __shared__ bool last_block;
    const int base = blockIdx.x * kTile + threadIdx.x * kItems;
    T sum = 0;
#pragma unroll
    for (int k = 0; k < kItems; k++)
        if (base + k < n) sum += f(base + k);
    T total;
    block_exclusive_sum(sum, smem, total);
    if (threadIdx.x == 0) {
        tile_sums[blockIdx.x] = total;
        __threadfence();                                        // the sum is visible before the ticket is drawn
        last_block = atomicAdd(&g_scan_tickets, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last_block) return;
    __threadfence();
    block_scan_tile_sums(tile_sums, int(gridDim.x), total_out, smem);
    if (threadIdx.x == 0) g_scan_tickets = 0;
}

template <typename T, typename F>
__global__ void __launch_bounds__(kThreads) scan_apply_tiles(F f, int n, const T* __restrict__ tile_offsets,
                                                             T* __restrict__ out) {
    __shared__ T smem[kThreads / 32 + 1];
    const int base = blockIdx.x * kTile + threadIdx.x * kItems;
    T v[kItems];
    T sum = 0;
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        v[k] = base + k < n ? f(base + k) : T(0);
        sum += v[k];
    }
    T total;
    T run = tile_offsets[blockIdx.x] + block_exclusive_sum(sum, smem, total);
#pragma unroll
    for (int k = 0; k < kItems; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    // the element one past the end receives the grand total (the reference scans n + 1 items)
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == kThreads - 1) out[n] = run;
}

/// out[i] = sum_{j<i} f(j) for i in [0, n]; out has n + 1 elements; *total_out = out[n].
/// `tile_scratch` needs num_tiles(n) elements. n == 0 writes out[0] = 0.
template <typename T, typename F>
void exclusive_scan(F f, int n, T* out, T* tile_scratch, T* total_out) {
    if (n <= 0) {
        HGB_CUDA(cudaMemsetAsync(out, 0, sizeof(T), 0));
        if (total_out) HGB_CUDA(cudaMemsetAsync(total_out, 0, sizeof(T), 0));
        return;
    }
    const int tiles = num_tiles(n);
    scan_reduce_tiles<T, F><<<tiles, kThreads>>>(f, n, tile_scratch, total_out); count_launch();
    scan_apply_tiles<T, F><<<tiles, kThreads>>>(f, n, tile_scratch, out); count_launch();
    HGB_CUDA(cudaGetLastError());
}

struct LoadInt {
    const int* data;
    __device__ __forceinline__ int operator()(int i) const { return data[i]; }
};

// ------------------------------------------------------------------ radix sort
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortItems = 8;                       // keys per thread
constexpr int kSortTile = kThreads * kSortItems;    // 2048 keys per block

inline int sort_tiles(int n) { return (n + kSortTile - 1) / kSortTile; }

/// hist[digit * tiles + tile] = number of keys of `tile` whose digit is `digit`
static __global__ void __launch_bounds__(kThreads) radix_histogram(const int* __restrict__ keys, int n, int shift, int tiles,
                                                            int* __restrict__ hist) {
    __shared__ int counts[kRadix];
    counts[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * kSortTile;
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const int i = base + k * kThreads + threadIdx.x;
        if (i < n) atomicAdd(&counts[(keys[i] >> shift) & (kRadix - 1)], 1);
    }
    __syncthreads();
    hist[threadIdx.x * tiles + blockIdx.x] = counts[threadIdx.x];
}

/// Stable scatter of one tile. Each warp owns a contiguous chunk of 32 * kSortItems
/// keys, read warp-striped so that (item, lane) order is input order.
static __global__ void __launch_bounds__(kThreads) radix_scatter(const int* __restrict__ keys_in, const int* __restrict__ vals_in,
                                                          int* __restrict__ keys_out, int* __restrict__ vals_out,
                                                          int n, int shift, int tiles, const int* __restrict__ offsets) {
    constexpr int kWarps = kThreads / 32;
    __shared__ int warp_counts[kWarps][kRadix];
    __shared__ int digit_base[kRadix];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w = 0; w < kWarps; w++) warp_counts[w][threadIdx.x] = 0;
    digit_base[threadIdx.x] = offsets[threadIdx.x * tiles + blockIdx.x];
    __syncthreads();

    const int chunk = blockIdx.x * kSortTile + warp * 32 * kSortItems;
    int key[kSortItems], rank[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const int i = chunk + k * 32 + lane;
        const bool live = i < n;
        key[k] = live ? keys_in[i] : 0;
        const int digit = live ? (key[k] >> shift) & (kRadix - 1) : kRadix;   // dead lanes match only each other
        const unsigned peers = __match_any_sync(kFullMask, digit);
        const int leader = __ffs(peers) - 1;
        int before = 0;
        if (live && lane == leader) {
            before = warp_counts[warp][digit];
            warp_counts[warp][digit] = before + __popc(peers);
        }
        before = __shfl_sync(kFullMask, before, leader);
        rank[k] = before + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive prefix over the warps of this block, per digit
        int run = 0;
        for (int w = 0; w < kWarps; w++) {
            const int c = warp_counts[w][threadIdx.x];
            warp_counts[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const int i = chunk + k * 32 + lane;
        if (i < n) {
            const int digit = (key[k] >> shift) & (kRadix - 1);
            const int pos = digit_base[digit] + warp_counts[warp][digit] + rank[k];
            keys_out[pos] = key[k];
            vals_out[pos] = vals_in[i];
        }
    }
}

/// Bytes of scratch sort_pairs needs (histogram + its scan tiles + total).
inline size_t sort_scratch_ints(int n) {
    const int hist = kRadix * sort_tiles(n);
    return size_t(hist) + 1 + num_tiles(hist) + 8;
}

/// Stable sort of (key, value) pairs on the low `bits` bits of the non-negative
/// keys. Ping-pongs between (keys, vals) and (keys_alt, vals_alt); returns true
/// when the result ended in the *_alt buffers.
inline bool sort_pairs(int* keys, int* vals, int* keys_alt, int* vals_alt, int n, int bits, int* scratch) {
    if (n <= 0 || bits <= 0) return false;
    const int tiles = sort_tiles(n);
    const int hist_n = kRadix * tiles;
    int* hist = scratch;                       // hist_n + 1 (scan output has n + 1 entries)
    int* scan_tiles = scratch + hist_n + 1;
    bool in_alt = false;
    for (int shift = 0; shift < bits; shift += kRadixBits) {
        int* kin = in_alt ? keys_alt : keys;   int* vin = in_alt ? vals_alt : vals;
        int* kout = in_alt ? keys : keys_alt;  int* vout = in_alt ? vals : vals_alt;
        radix_histogram<<<tiles, kThreads>>>(kin, n, shift, tiles, hist); count_launch();
        exclusive_scan<int>(LoadInt{hist}, hist_n, hist, scan_tiles, (int*)nullptr);
        radix_scatter<<<tiles, kThreads>>>(kin, vin, kout, vout, n, shift, tiles, hist); count_launch();
        in_alt = !in_alt;
    }
    HGB_CUDA(cudaGetLastError());
    return in_alt;
}

} // namespace prim
} // namespace hagrid
