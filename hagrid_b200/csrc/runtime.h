// Internal runtime helpers shared by the kernels' host code.
#pragma once

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace hagrid {

/// Any CUDA failure is fatal, as in the reference (src/common.h:101-108).
inline void check_cuda(cudaError_t err, const char* what, const char* file, int line) {
    if (err != cudaSuccess) {
        std::fprintf(stderr, "%s(%d): %s failed: %s\n", file, line, what, cudaGetErrorString(err));
        std::abort();
    }
}

/// Number of kernels this library has launched since it was loaded (bench.py reports
/// the count inside its timed region as `gpu_launches`).
extern std::atomic<unsigned long long> g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

/// Multiprocessors of the current device (queried once per device): grids of grid-stride kernels are sized by it
int sm_count();

struct Tri; struct Ray; struct Hit;
/// Stream-taking forms of generate_bounce_rays / count_hits (hgb_api.h), used by the two-wave frame; without `keys` ray i's
/// random stream is named first_key + i (a chunk of a buffer passes where it starts)
void generate_bounce_rays_on(cudaStream_t stream, const Tri* tris, int num_tris, const Ray* rays, const Hit* hits, int num_rays,
                             float offset, float tmax, unsigned seed, Ray* out, const int* keys, int first_key);
void count_hits_on(cudaStream_t stream, const Hit* hits, int num_hits, unsigned long long* counters);

} // namespace hagrid

#define HGB_CUDA(call) ::hagrid::check_cuda((call), #call, __FILE__, __LINE__)
