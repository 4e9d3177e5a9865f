// The toolkit's CUB (2.8, shipped with CUDA 12.9) behind a tiny C interface, for tools/gpu_primitives_bench.py: the
// bar SURVEY.md section 2.3 names for the hand-written primitives. Test/measurement infrastructure, not product code.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

static void* g_tmp = nullptr;
static size_t g_tmp_bytes = 0;
static void* scratch(size_t bytes) {
    if (bytes > g_tmp_bytes) { cudaFree(g_tmp); cudaMalloc(&g_tmp, bytes); g_tmp_bytes = bytes; }
    return g_tmp;
}

extern "C" int cub_exclusive_sum_i32(const int* in, int* out, int n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n);
    void* tmp = scratch(bytes);
    return int(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n));
}

extern "C" int cub_exclusive_sum_u64(const unsigned long long* in, unsigned long long* out, int n) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n);
    void* tmp = scratch(bytes);
    return int(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n));
}

extern "C" int cub_sort_pairs(int* keys, int* vals, int* keys_alt, int* vals_alt, int n, int bits) {
    cub::DoubleBuffer<int> k(keys, keys_alt), v(vals, vals_alt);
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, k, v, n, 0, bits);
    void* tmp = scratch(bytes);
    const int rc = int(cub::DeviceRadixSort::SortPairs(tmp, bytes, k, v, n, 0, bits));
    return rc ? -rc : (k.Current() == keys_alt ? 1 : 0);
}

extern "C" int cub_partition_flagged(const int* in, const int* flags, int* out, int* num_selected, int n) {
    size_t bytes = 0;
    cub::DevicePartition::Flagged(nullptr, bytes, in, flags, out, num_selected, n);
    void* tmp = scratch(bytes);
    return int(cub::DevicePartition::Flagged(tmp, bytes, in, flags, out, num_selected, n));
}

extern "C" int cub_reduce_sum_i32(const int* in, int* out, int n) {
    size_t bytes = 0;
    cub::DeviceReduce::Sum(nullptr, bytes, in, out, n);
    void* tmp = scratch(bytes);
    return int(cub::DeviceReduce::Sum(tmp, bytes, in, out, n));
}
