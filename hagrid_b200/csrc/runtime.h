// Internal runtime helpers shared by the kernels' host code.
#pragma once

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

} // namespace hagrid

#define HGB_CUDA(call) ::hagrid::check_cuda((call), #call, __FILE__, __LINE__)
