// Device-wide primitives of the construction pipeline, hand-written for
// sm_100a (the reference wraps CUB's DeviceScan / DeviceReduce /
// DevicePartition / DeviceRadixSort, src/parallel.cuh:12-89).
//
//   * exclusive_scan<T>(): single pass with decoupled look-back over 2048-element
//     tiles and a fused input functor, one launch; T = int or a packed 64-bit
//     pair of counters. The total stays on the device (no host round trip per
//     call, unlike parallel.cuh:40); callers fetch several totals with one copy.
//   * sort_pairs(): stable LSD radix sort, 8 bits per pass, onesweep style (one
//     counting kernel for all passes, one kernel per pass with a look-back over
//     the tiles' digit counts), warp-match ranking (__match_any_sync) so equal
//     keys keep their input order — the property the build relies on for the
//     reference's reference order.
//   * the reference's flagged partition (kept first, rejected reversed at the
//     rear, CUB semantics) is not materialised at all: the build derives both
//     target positions from one scan (see grid_build.cu).
//
// All kernels are plain grid launches on the legacy default stream. In the scan and
// in the sort's passes a block may wait for blocks that were handed an earlier
// tile, i.e. blocks that are already running.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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
// Single pass with decoupled look-back (Merrill & Garland): every tile reads its input once, publishes its sum, looks
// back over the sums (or finished prefixes) of the tiles before it, and writes its outputs -- 2n memory traffic where
// reduce-then-scan moves 3n. Tiles are handed out by a counter in launch order, so the tile a block waits for is
// always held by a block that is already running (the wait cannot deadlock). A status word carries value and state
// together (one 8-byte store publishes both); a 64-bit T (two packed 32-bit counters) uses one word per half.
constexpr unsigned long long kStatusAggregate = 1ull << 32, kStatusPrefix = 2ull << 32;

template <typename T> struct ScanLanes { static constexpr int value = int(sizeof(T) / 4); };

/// Elements of T the scratch buffer of exclusive_scan needs for n items
template <typename T>
inline size_t scan_scratch_elems(int n) {
    // per tile ScanLanes<T>::value status words of 8 bytes, + one word for the tile counter
    return (size_t(num_tiles(n)) * ScanLanes<T>::value * 8 + 8 + sizeof(T) - 1) / sizeof(T) + 1;
}

template <typename T>
__device__ __forceinline__ void publish(unsigned long long* status, int tile, T value, unsigned long long state) {
    constexpr int kLanes = ScanLanes<T>::value;
    const unsigned long long v = (unsigned long long)value;
#pragma unroll
    for (int l = 0; l < kLanes; l++)
        *reinterpret_cast<volatile unsigned long long*>(status + size_t(tile) * kLanes + l) = ((v >> (32 * l)) & 0xFFFFFFFFull) | state;
}

/// Waits until tile `tile` has published something; returns its value and whether it is a finished prefix. The halves
/// of a 64-bit value are two words with their own state: equal states mean both belong to the same publication
/// (aggregate or prefix), anything else is re-read.
template <typename T>
__device__ __forceinline__ T peek(const unsigned long long* status, int tile, bool& is_prefix) {
    constexpr int kLanes = ScanLanes<T>::value;
    const volatile unsigned long long* words = status + size_t(tile) * kLanes;
    unsigned long long w[kLanes];
    bool settled;
    do {
        settled = true;
#pragma unroll
        for (int l = 0; l < kLanes; l++) {
            w[l] = words[l];
            settled = settled && (w[l] >> 32) != 0 && (w[l] >> 32) == (w[0] >> 32);
        }
    } while (!settled);
    unsigned long long v = 0;
#pragma unroll
    for (int l = 0; l < kLanes; l++) v |= (w[l] & 0xFFFFFFFFull) << (32 * l);
    is_prefix = (w[0] >> 32) == 2;
    return T(v);
}

constexpr int kScanThreads = 512;
template <typename T> struct ScanItems { static constexpr int value = 64 / int(sizeof(T)); };       // 16 ints or 8 packed pairs per thread
template <typename T> constexpr int scan_tile() { return kScanThreads * ScanItems<T>::value; }     // 8192 / 4096 items

inline int num_scan_tiles(int n, int tile) { return (n + tile - 1) / tile; }

/// Exclusive prefix of `v` over a block of kScanThreads threads; `total` = block sum. `smem` needs kScanThreads / 32 + 1 elements.
template <typename T>
__device__ __forceinline__ T scan_block_exclusive_sum(T v, T* smem, T& total) {
    constexpr int kWarps = kScanThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T incl = warp_inclusive_sum(v);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = lane < kWarps ? smem[lane] : T(0);
        const T wi = warp_inclusive_sum(w);
        if (lane < kWarps) smem[lane] = wi - w;
        if (lane == kWarps - 1) smem[kWarps] = wi;
    }
    __syncthreads();
    total = smem[kWarps];
    return smem[warp] + incl - v;
}

/// One tile: items are read striped (item k * threads + t by thread t: every load instruction of a warp covers
/// consecutive items, whatever the functor reads), turned into a blocked arrangement through padded shared memory
/// (thread t owns items t*K .. t*K+K-1, conflict-free with one pad word per 32), scanned, and written back the same way.
template <typename T, typename F>
__global__ void __launch_bounds__(kScanThreads) scan_single_pass(F f, int n, T* out, unsigned long long* status,
                                                                 unsigned* tile_counter, T* total_out) {
    constexpr int K = ScanItems<T>::value;
    constexpr int kTileItems = kScanThreads * K;
    __shared__ T buf[kTileItems + kTileItems / 32];
    __shared__ T sums[kScanThreads / 32 + 1];
    __shared__ int my_tile;
    __shared__ T tile_prefix;
    if (threadIdx.x == 0) my_tile = int(atomicAdd(tile_counter, 1u));
    __syncthreads();
    const int tile = my_tile;
    const int base = tile * kTileItems;
    auto padded = [](int i) { return i + (i >> 5); };
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int j = k * kScanThreads + int(threadIdx.x);
        buf[padded(j)] = base + j < n ? f(base + j) : T(0);
    }
    __syncthreads();
    T v[K];
    T sum = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        v[k] = buf[padded(int(threadIdx.x) * K + k)];
        sum += v[k];
    }
    T total;
    const T within = scan_block_exclusive_sum(sum, sums, total);
    if (threadIdx.x < 32) {
        // warp 0 publishes the tile's sum, then walks back over its predecessors 32 at a time
        T prefix = 0;
        if (tile == 0) {
            if (threadIdx.x == 0) publish<T>(status, 0, total, kStatusPrefix);
        } else {
            if (threadIdx.x == 0) publish<T>(status, tile, total, kStatusAggregate);
            for (int first = tile - 1; first >= 0; first -= 32) {
                const int look = first - int(threadIdx.x);
                bool is_prefix = false;
                T value = 0;
                if (look >= 0) value = peek<T>(status, look, is_prefix);
                const unsigned done = __ballot_sync(kFullMask, is_prefix);
                // lanes up to and including the nearest finished prefix contribute
                const int stop = done ? __ffs(done) - 1 : 31;
                T part = int(threadIdx.x) <= stop && look >= 0 ? value : T(0);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(kFullMask, part, d);
                prefix += part;
                if (done) break;
            }
            if (threadIdx.x == 0) publish<T>(status, tile, prefix + total, kStatusPrefix);
        }
        if (threadIdx.x == 0) tile_prefix = prefix;
    }
    __syncthreads();
    T run = tile_prefix + within;
#pragma unroll
    for (int k = 0; k < K; k++) {
        buf[padded(int(threadIdx.x) * K + k)] = run;
        run += v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int j = k * kScanThreads + int(threadIdx.x);
        if (base + j < n) out[base + j] = buf[padded(j)];
    }
    // the element one past the end receives the grand total (the reference scans n + 1 items)
    if (tile == int(gridDim.x) - 1 && threadIdx.x == kScanThreads - 1) {
        out[n] = run;
        if (total_out) *total_out = run;
    }
}

/// out[i] = sum_{j<i} f(j) for i in [0, n]; out has n + 1 elements (it may alias what f reads: every tile loads its
/// items before it stores them); *total_out = out[n]. `scratch` needs scan_scratch_elems<T>(n) elements.
/// n == 0 writes out[0] = 0. The halves of a 64-bit T must each stay below 2^32 in every prefix.
template <typename T, typename F>
void exclusive_scan(F f, int n, T* out, T* scratch, T* total_out) {
    if (n <= 0) {
        HGB_CUDA(cudaMemsetAsync(out, 0, sizeof(T), 0));
        if (total_out) HGB_CUDA(cudaMemsetAsync(total_out, 0, sizeof(T), 0));
        return;
    }
    const int tiles = num_scan_tiles(n, scan_tile<T>());
    auto words = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(scratch) + 7) & ~uintptr_t(7));
    const size_t status_words = size_t(tiles) * ScanLanes<T>::value;
    HGB_CUDA(cudaMemsetAsync(words, 0, (status_words + 1) * 8, 0));
    scan_single_pass<T, F><<<tiles, kScanThreads>>>(f, n, out, words, reinterpret_cast<unsigned*>(words + status_words), total_out); count_launch();
    HGB_CUDA(cudaGetLastError());
}

struct LoadInt {
    const int* data;
    __device__ __forceinline__ int operator()(int i) const { return data[i]; }
};

// ------------------------------------------------------------------ radix sort
// Stable LSD radix sort of (key, value) pairs, 8 bits per pass, "onesweep" style: one kernel counts the digits of
// ALL passes at once (each key is read once for that), then every pass is a single kernel that reads a tile of keys
// and values once and writes them once -- where a tile's keys of one digit go is found by a decoupled look-back over
// the digit counts of the tiles before it (one status word per tile and digit, the pass number in its flag bits so
// that the words are cleared once per sort, not once per pass). Ranks inside a tile come from __match_any_sync on
// warp-striped keys (equal keys keep their input order), and keys and values pass through shared memory in sorted
// order, so a warp's stores are runs of consecutive addresses. Tiles are handed out by a counter, as in the scan.
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                          // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;    // 4096 keys per tile
constexpr int kSortMaxPasses = 4;
constexpr unsigned kSortValueBits = 28, kSortValueMask = (1u << kSortValueBits) - 1u;   // n < 2^28

inline int sort_tiles(int n) { return (n + kSortTile - 1) / kSortTile; }

/// hist[pass * 256 + digit] += keys whose digit of pass `pass` is `digit`, for every pass of the sort
static __global__ void __launch_bounds__(kSortThreads) radix_histograms(const int* __restrict__ keys, int n, int bits, unsigned* __restrict__ hist) {
    __shared__ unsigned counts[kSortMaxPasses][kRadix];
    const int passes = (bits + kRadixBits - 1) / kRadixBits;
    for (int p = 0; p < kSortMaxPasses; p++) counts[p][threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * kSortThreads + threadIdx.x; i < n; i += gridDim.x * kSortThreads) {
        const unsigned key = unsigned(keys[i]);
        for (int p = 0; p < passes; p++) {
            const int left = min(kRadixBits, bits - p * kRadixBits);
            atomicAdd(&counts[p][(key >> (p * kRadixBits)) & ((1u << left) - 1u)], 1u);
        }
    }
    __syncthreads();
    for (int p = 0; p < passes; p++)
        if (counts[p][threadIdx.x]) atomicAdd(hist + p * kRadix + threadIdx.x, counts[p][threadIdx.x]);
}

/// hist[pass][digit] -> number of keys with a smaller digit in that pass (one block, one warp-scan per pass)
static __global__ void __launch_bounds__(kRadix) radix_digit_offsets(unsigned* hist, int passes) {
    __shared__ unsigned warp_sums[kRadix / 32];
    for (int p = 0; p < passes; p++) {
        const unsigned c = hist[p * kRadix + threadIdx.x];
        const unsigned incl = warp_inclusive_sum(c);
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned before = 0;
        for (int w = 0; w < int(threadIdx.x >> 5); w++) before += warp_sums[w];
        hist[p * kRadix + threadIdx.x] = before + incl - c;
        __syncthreads();
    }
}

static __global__ void __launch_bounds__(kSortThreads)
radix_onesweep(const int* __restrict__ keys_in, const int* __restrict__ vals_in, int* __restrict__ keys_out, int* __restrict__ vals_out,
               int n, int shift, int digit_bits, const unsigned* __restrict__ digit_offsets, unsigned* status, unsigned* tile_counter, unsigned pass) {
    constexpr int kWarps = kSortThreads / 32;
    __shared__ int warp_counts[kWarps][kRadix];     // per warp and digit: count, then exclusive prefix over the warps
    __shared__ int local_start[kRadix];             // where the digit's run starts in the sorted tile
    __shared__ int global_start[kRadix];            // where it starts in the output
    __shared__ int sorted_keys[kSortTile], sorted_vals[kSortTile];
    __shared__ unsigned scan_tmp[kRadix / 32];
    __shared__ int my_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned digit_mask = (1u << digit_bits) - 1u;
    if (threadIdx.x == 0) my_tile = int(atomicAdd(tile_counter, 1u));
    for (int w = 0; w < kWarps; w++) warp_counts[w][threadIdx.x] = 0;
    __syncthreads();
    const int tile = my_tile;
    const int chunk = tile * kSortTile + warp * 32 * kSortItems;     // each warp owns 512 consecutive keys, read striped

    int key[kSortItems], rank[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const int i = chunk + k * 32 + lane;
        const bool live = i < n;
        key[k] = live ? keys_in[i] : 0;
        const int digit = live ? int((unsigned(key[k]) >> shift) & digit_mask) : kRadix;   // dead lanes match only each other
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
    {   // thread d owns digit d: its count in this tile, the warps' prefixes, and the look-back over the tiles before
        const int d = threadIdx.x;
        int count = 0;
        for (int w = 0; w < kWarps; w++) {
            const int c = warp_counts[w][d];
            warp_counts[w][d] = count;
            count += c;
        }
        const unsigned aggregate = (2 * pass + 1) << kSortValueBits, prefix_flag = (2 * pass + 2) << kSortValueBits;
        volatile unsigned* words = status;
        unsigned before = 0;
        if (tile > 0) {
            words[size_t(tile) * kRadix + d] = aggregate | unsigned(count);
            for (int prev = tile - 1; prev >= 0; ) {
                const unsigned w = words[size_t(prev) * kRadix + d];
                const unsigned flag = w & ~kSortValueMask;
                if (flag == prefix_flag) { before += w & kSortValueMask; break; }
                if (flag == aggregate) { before += w & kSortValueMask; prev--; }
            }
        }
        words[size_t(tile) * kRadix + d] = prefix_flag | (before + unsigned(count));
        global_start[d] = int(digit_offsets[d] + before);
        // where the digit's run starts inside the tile: exclusive scan of the counts over the digits
        const unsigned incl = warp_inclusive_sum(unsigned(count));
        if (lane == 31) scan_tmp[warp] = incl;
        __syncthreads();
        unsigned lower = 0;
        for (int w = 0; w < warp; w++) lower += scan_tmp[w];
        local_start[d] = int(lower + incl) - count;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; k++) {
        const int i = chunk + k * 32 + lane;
        if (i < n) {
            const int digit = int((unsigned(key[k]) >> shift) & digit_mask);
            const int pos = local_start[digit] + warp_counts[warp][digit] + rank[k];
            sorted_keys[pos] = key[k];
            sorted_vals[pos] = vals_in[i];
        }
    }
    __syncthreads();
    const int live_keys = min(kSortTile, n - tile * kSortTile);
    for (int i = threadIdx.x; i < live_keys; i += kSortThreads) {
        const int k = sorted_keys[i];
        const int digit = int((unsigned(k) >> shift) & digit_mask);
        const int pos = global_start[digit] + (i - local_start[digit]);
        keys_out[pos] = k;
        vals_out[pos] = sorted_vals[i];
    }
}

/// Ints of scratch sort_pairs needs (digit counts, tile counters, one status word per tile and digit).
inline size_t sort_scratch_ints(int n) { return size_t(kSortMaxPasses) * kRadix + 8 + size_t(sort_tiles(n)) * kRadix + 16; }

/// Stable sort of (key, value) pairs on the low `bits` bits (at most 32 - 4 passes of 8) of the keys; n < 2^28.
/// Ping-pongs between (keys, vals) and (keys_alt, vals_alt); returns true when the result ended in the *_alt buffers.
/// Enqueued on `stream` (the construction uses the legacy default stream).
inline bool sort_pairs(int* keys, int* vals, int* keys_alt, int* vals_alt, int n, int bits, int* scratch, cudaStream_t stream = 0) {
    if (n <= 0 || bits <= 0) return false;
    if (n >= (1 << kSortValueBits) || bits > kSortMaxPasses * kRadixBits) {
        std::fprintf(stderr, "hagrid_b200: sort_pairs handles fewer than 2^28 pairs and at most 32 key bits\n");
        std::abort();
    }
    const int tiles = sort_tiles(n);
    const int passes = (bits + kRadixBits - 1) / kRadixBits;
    unsigned* hist = reinterpret_cast<unsigned*>(scratch);               // kSortMaxPasses * 256
    unsigned* counters = hist + kSortMaxPasses * kRadix;                 // one tile counter per pass
    unsigned* status = counters + 8;                                     // tiles * 256
    HGB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int) * (size_t(kSortMaxPasses) * kRadix + 8 + size_t(tiles) * kRadix), stream));
    radix_histograms<<<std::min(tiles, sm_count() * 8), kSortThreads, 0, stream>>>(keys, n, bits, hist); count_launch();
    radix_digit_offsets<<<1, kRadix, 0, stream>>>(hist, passes); count_launch();
    bool in_alt = false;
    for (int p = 0; p < passes; p++) {
        int* kin = in_alt ? keys_alt : keys;   int* vin = in_alt ? vals_alt : vals;
        int* kout = in_alt ? keys : keys_alt;  int* vout = in_alt ? vals : vals_alt;
        radix_onesweep<<<tiles, kSortThreads, 0, stream>>>(kin, vin, kout, vout, n, p * kRadixBits, std::min(kRadixBits, bits - p * kRadixBits),
                                                hist + p * kRadix, status, counters + p, unsigned(p)); count_launch();
        in_alt = !in_alt;
    }
    HGB_CUDA(cudaGetLastError());
    return in_alt;
}

} // namespace prim
} // namespace hagrid
