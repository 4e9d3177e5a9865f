// Callable forms of the device-wide primitives: what the reference's `Parallel` wrapper offers on top of CUB
// (src/parallel.cuh:12-89: scan, reduce, partition, sort_pairs), on this library's own kernels (primitives.cuh). The
// construction pipeline uses the templates directly (fused input functors, several totals fetched with one copy); these
// entry points exist so that the primitives can be tested and timed on their own (tests/test_primitives.py,
// tools/gpu_primitives_bench.py against the toolkit's CUB).
#include <algorithm>

#include "hgb_api.h"
#include "primitives.cuh"
#include "runtime.h"

namespace hagrid {

namespace {

struct LoadU64 {
    const unsigned long long* data;
    __device__ __forceinline__ unsigned long long operator()(int i) const { return data[i]; }
};

/// out[i] for a kept item = its rank among the kept ones; for a rejected one n - 1 - its rank among the rejected:
/// cub::DevicePartition::Flagged's order (selected first in input order, the rest reversed at the rear), the order
/// the reference's build relies on (src/build.cu:568-569)
__global__ void __launch_bounds__(256) partition_scatter(const int* __restrict__ in, const int* __restrict__ flags,
                                                         const int* __restrict__ kept_before, int n, int* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int k = kept_before[i];
    out[flags[i] ? k : n - 1 - (i - k)] = in[i];
}

struct FlagSet {
    const int* flags;
    __device__ __forceinline__ int operator()(int i) const { return flags[i] != 0; }
};

/// float <-> unsigned whose integer order is the float order (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned ordered(float f) {
    const unsigned u = __float_as_uint(f);
    return u & 0x80000000u ? ~u : u | 0x80000000u;
}
__device__ __forceinline__ float unordered(unsigned u) { return __uint_as_float(u & 0x80000000u ? u & 0x7FFFFFFFu : ~u); }

template <int kOp>
__global__ void __launch_bounds__(256) reduce_kernel(const void* __restrict__ in, int n, unsigned* __restrict__ acc) {
    // kOp 0: sum of int32, 1: max of int32, 2: min of float, 3: max of float
    const int* ints = static_cast<const int*>(in);
    const float* floats = static_cast<const float*>(in);
    int isum = 0, imax = INT_MIN;
    float fmin_ = INFINITY, fmax_ = -INFINITY;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        if (kOp == 0) isum += ints[i];
        if (kOp == 1) imax = max(imax, ints[i]);
        if (kOp == 2) fmin_ = fminf(fmin_, floats[i]);
        if (kOp == 3) fmax_ = fmaxf(fmax_, floats[i]);
    }
    for (int d = 16; d > 0; d >>= 1) {
        if (kOp == 0) isum += __shfl_xor_sync(prim::kFullMask, isum, d);
        if (kOp == 1) imax = max(imax, __shfl_xor_sync(prim::kFullMask, imax, d));
        if (kOp == 2) fmin_ = fminf(fmin_, __shfl_xor_sync(prim::kFullMask, fmin_, d));
        if (kOp == 3) fmax_ = fmaxf(fmax_, __shfl_xor_sync(prim::kFullMask, fmax_, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (kOp == 0) atomicAdd(reinterpret_cast<int*>(acc), isum);
        if (kOp == 1) atomicMax(reinterpret_cast<int*>(acc), imax);
        if (kOp == 2) atomicMin(acc, ordered(fmin_));
        if (kOp == 3) atomicMax(acc, ordered(fmax_));
    }
}

__global__ void finish_float_reduce(unsigned* acc) { *reinterpret_cast<float*>(acc) = unordered(*acc); }

} // namespace

void prim_exclusive_scan(MemManager& mem, const void* in, int n, int elem_bytes, void* out) {
    if (elem_bytes == 8) {
        auto tmp = mem.alloc<unsigned long long>(prim::scan_scratch_elems<unsigned long long>(n));
        prim::exclusive_scan<unsigned long long>(LoadU64{static_cast<const unsigned long long*>(in)}, n,
                                                 static_cast<unsigned long long*>(out), tmp, (unsigned long long*)nullptr);
        mem.free(tmp);
    } else {
        int* tmp = mem.alloc<int>(prim::scan_scratch_elems<int>(n));
        prim::exclusive_scan<int>(prim::LoadInt{static_cast<const int*>(in)}, n, static_cast<int*>(out), tmp, (int*)nullptr);
        mem.free(tmp);
    }
}

void prim_reduce(MemManager& mem, const void* in, int n, int op, void* out) {
    // identity first (the reference passes it as `init`, src/parallel.cuh:44), then one grid-stride pass with atomics
    auto acc = static_cast<unsigned*>(out);
    const unsigned identity[4] = {0u, 0x80000000u /* INT_MIN */, 0xFF800000u /* ordered(+inf) */, 0x007FFFFFu /* ordered(-inf) */};
    HGB_CUDA(cudaMemcpyAsync(acc, &identity[op], sizeof(unsigned), cudaMemcpyHostToDevice, 0));
    (void)mem;
    if (n > 0) {
        const int blocks = std::min((n + 255) / 256, sm_count() * 8);
        if (op == 0) reduce_kernel<0><<<blocks, 256>>>(in, n, acc);
        else if (op == 1) reduce_kernel<1><<<blocks, 256>>>(in, n, acc);
        else if (op == 2) reduce_kernel<2><<<blocks, 256>>>(in, n, acc);
        else reduce_kernel<3><<<blocks, 256>>>(in, n, acc);
        count_launch();
    }
    if (op >= 2) { finish_float_reduce<<<1, 1>>>(acc); count_launch(); }
    HGB_CUDA(cudaGetLastError());
}

int prim_partition(MemManager& mem, const int* in, const int* flags, int n, int* out) {
    if (n <= 0) return 0;
    int* kept_before = mem.alloc<int>(size_t(n) + 1);
    int* tmp = mem.alloc<int>(prim::scan_scratch_elems<int>(n));
    prim::exclusive_scan<int>(FlagSet{flags}, n, kept_before, tmp, (int*)nullptr);
    partition_scatter<<<(n + 255) / 256, 256>>>(in, flags, kept_before, n, out); count_launch();
    HGB_CUDA(cudaGetLastError());
    int kept = 0;
    HGB_CUDA(cudaMemcpy(&kept, kept_before + n, sizeof(int), cudaMemcpyDeviceToHost));
    mem.free(kept_before);
    mem.free(tmp);
    return kept;
}

void prim_sort_pairs(MemManager& mem, int* keys, int* vals, int n, int bits) {
    if (n <= 0 || bits <= 0) return;
    int* keys_alt = mem.alloc<int>(n);
    int* vals_alt = mem.alloc<int>(n);
    int* tmp = mem.alloc<int>(prim::sort_scratch_ints(n));
    if (prim::sort_pairs(keys, vals, keys_alt, vals_alt, n, bits, tmp)) {
        HGB_CUDA(cudaMemcpyAsync(keys, keys_alt, sizeof(int) * size_t(n), cudaMemcpyDeviceToDevice, 0));
        HGB_CUDA(cudaMemcpyAsync(vals, vals_alt, sizeof(int) * size_t(n), cudaMemcpyDeviceToDevice, 0));
    }
    HGB_CUDA(cudaStreamSynchronize(0));
    mem.free(keys_alt); mem.free(vals_alt); mem.free(tmp);
}

} // namespace hagrid
