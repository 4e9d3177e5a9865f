// Second-wave rays on the device (BASELINE.json config C5): one diffuse bounce off the
// primary hit points. The reference has no such stage; its front end stops at primary rays
// (src/main.cpp:52-66), so this file defines the operation and the test suite's CPU checker restates it
// with the same operation order. Every float operation is a single IEEE
// round-to-nearest instruction (no contraction, no approximate reciprocal, no trigonometry), so the CPU
// restatement and this kernel agree bit for bit.
#include <algorithm>

#include "hgb_api.h"
#include "runtime.h"

namespace hagrid {

namespace {

/// Integer mixer of the counter-based generator (two multiply/xor-shift rounds)
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

/// k-th uniform draw in [0, 1) of the stream `base`: 24 random bits, exact in float
__device__ __forceinline__ float draw(uint32_t base, uint32_t k) {
    return __fmul_rn(__uint2float_rn(mix32(base + k) >> 8), 0x1p-24f);
}

constexpr int kDiskTries = 8;

__global__ void __launch_bounds__(256)
bounce_rays(const Tri* __restrict__ tris, int num_tris, const Ray* rays, const Hit* __restrict__ hits, int num_rays,
            float offset, float tmax, uint32_t seed, const int* __restrict__ keys, int first_key, Ray* out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= num_rays) return;
    const float4 a = reinterpret_cast<const float4*>(rays + i)[0];     // org, tmin
    const float4 b = reinterpret_cast<const float4*>(rays + i)[1];     // dir, tmax
    const float4 h = reinterpret_cast<const float4*>(hits)[i];         // id, t, u, v
    const int id = __float_as_int(h.x);
    float4* dst = reinterpret_cast<float4*>(out + i);
    float nx = 0.0f, ny = 0.0f, nz = 0.0f, len = 0.0f;
    if (id >= 0 && id < num_tris) {
        const float4* t = reinterpret_cast<const float4*>(tris + id);
        nx = __ldg(t).w; ny = __ldg(t + 1).w; nz = __ldg(t + 2).w;
        len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
    }
    if (!(len > 0.0f)) {        // miss (or a degenerate triangle): the ray is emitted again unchanged
        dst[0] = a; dst[1] = b;
        return;
    }
    nx = __fdiv_rn(nx, len); ny = __fdiv_rn(ny, len); nz = __fdiv_rn(nz, len);
    // the shading normal faces the incoming ray
    const float facing = __fadd_rn(__fadd_rn(__fmul_rn(nx, b.x), __fmul_rn(ny, b.y)), __fmul_rn(nz, b.z));
    if (facing > 0.0f) { nx = -nx; ny = -ny; nz = -nz; }
    // origin: hit point pushed off the surface
    const float px = __fadd_rn(__fadd_rn(a.x, __fmul_rn(b.x, h.y)), __fmul_rn(nx, offset));
    const float py = __fadd_rn(__fadd_rn(a.y, __fmul_rn(b.y, h.y)), __fmul_rn(ny, offset));
    const float pz = __fadd_rn(__fadd_rn(a.z, __fmul_rn(b.z, h.y)), __fmul_rn(nz, offset));
    // uniform point of the unit disk by rejection, lifted to the hemisphere: cosine-weighted direction
    // the random stream belongs to the ray, not to its place in this buffer: a shard of a frame passes the rays' indices
    // in the whole frame as keys and gets the rays the unsharded frame would get
    const uint32_t base = mix32(seed ^ mix32(uint32_t(keys ? __ldg(keys + i) : first_key + i)));
    float dx = 0.0f, dy = 0.0f, s = 0.0f;
    for (int k = 0; k < kDiskTries; k++) {
        const float x = __fsub_rn(__fmul_rn(2.0f, draw(base, 2 * k)), 1.0f);
        const float y = __fsub_rn(__fmul_rn(2.0f, draw(base, 2 * k + 1)), 1.0f);
        const float q = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
        if (q < 1.0f) { dx = x; dy = y; s = q; break; }
    }
    const float dz = __fsqrt_rn(__fsub_rn(1.0f, s));
    // tangent frame: t1 = normalize(n x axis), t2 = n x t1; axis = y where n is close to x, x elsewhere
    float t1x, t1y, t1z;
    if (fabsf(nx) > 0.9f) { t1x = -nz; t1y = 0.0f; t1z = nx; }
    else                  { t1x = 0.0f; t1y = nz; t1z = -ny; }
    const float tl = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(t1x, t1x), __fmul_rn(t1y, t1y)), __fmul_rn(t1z, t1z)));
    t1x = __fdiv_rn(t1x, tl); t1y = __fdiv_rn(t1y, tl); t1z = __fdiv_rn(t1z, tl);
    const float t2x = __fsub_rn(__fmul_rn(ny, t1z), __fmul_rn(nz, t1y));
    const float t2y = __fsub_rn(__fmul_rn(nz, t1x), __fmul_rn(nx, t1z));
    const float t2z = __fsub_rn(__fmul_rn(nx, t1y), __fmul_rn(ny, t1x));
    const float ox = __fadd_rn(__fadd_rn(__fmul_rn(t1x, dx), __fmul_rn(t2x, dy)), __fmul_rn(nx, dz));
    const float oy = __fadd_rn(__fadd_rn(__fmul_rn(t1y, dx), __fmul_rn(t2y, dy)), __fmul_rn(ny, dz));
    const float oz = __fadd_rn(__fadd_rn(__fmul_rn(t1z, dx), __fmul_rn(t2z, dy)), __fmul_rn(nz, dz));
    dst[0] = make_float4(px, py, pz, 0.0f);
    dst[1] = make_float4(ox, oy, oz, tmax);
}

/// counters[0] += hits that name a primitive (id >= 0), counters[1] += sum of (id + 1): the per-frame figures a sharded
/// frame all-reduces. One 16-byte load per hit, a warp reduction, two atomics per block.
__global__ void __launch_bounds__(256)
count_hits_kernel(const Hit* __restrict__ hits, int num_hits, unsigned long long* __restrict__ counters) {
    unsigned long long found = 0, sum = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < num_hits; i += gridDim.x * 256) {
        const int id = __float_as_int(__ldg(reinterpret_cast<const float4*>(hits) + i).x);
        found += id >= 0;
        sum += (unsigned long long)(long long)(id + 1);
    }
    for (int d = 16; d > 0; d >>= 1) {
        found += __shfl_xor_sync(0xFFFFFFFFu, found, d);
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    }
    __shared__ unsigned long long part[2][8];
    if ((threadIdx.x & 31) == 0) { part[0][threadIdx.x >> 5] = found; part[1][threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned long long total = 0;
        for (int w = 0; w < 8; w++) total += part[threadIdx.x][w];
        atomicAdd(counters + threadIdx.x, total);
    }
}

} // namespace

void generate_bounce_rays_on(cudaStream_t stream, const Tri* tris, int num_tris, const Ray* rays, const Hit* hits, int num_rays,
                             float offset, float tmax, unsigned seed, Ray* out, const int* keys, int first_key) {
    if (num_rays <= 0) return;
    bounce_rays<<<(num_rays + 255) / 256, 256, 0, stream>>>(tris, num_tris, rays, hits, num_rays, offset, tmax, seed, keys, first_key, out); count_launch();
    HGB_CUDA(cudaGetLastError());
}

void count_hits_on(cudaStream_t stream, const Hit* hits, int num_hits, unsigned long long* counters) {
    if (num_hits <= 0) return;
    const int blocks = std::min((num_hits + 255) / 256, sm_count() * 8);
    count_hits_kernel<<<blocks, 256, 0, stream>>>(hits, num_hits, counters); count_launch();
    HGB_CUDA(cudaGetLastError());
}

void generate_bounce_rays(const Tri* tris, int num_tris, const Ray* rays, const Hit* hits, int num_rays,
                          float offset, float tmax, unsigned seed, Ray* out, const int* keys) {
    generate_bounce_rays_on(0, tris, num_tris, rays, hits, num_rays, offset, tmax, seed, out, keys, 0);
}

void count_hits(const Hit* hits, int num_hits, unsigned long long* counters) { count_hits_on(0, hits, num_hits, counters); }

} // namespace hagrid
