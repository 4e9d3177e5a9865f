// Device side of the .rays format: 24-byte {org, dir} records <-> 32-byte Ray{org, tmin, dir, tmax}.
#include "formats.h"
#include "runtime.h"

namespace hagrid {

namespace {

__global__ void __launch_bounds__(256) expand_records(const float* __restrict__ rec, long long count, float tmin, float tmax,
                                                      Ray* __restrict__ rays) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= count) return;
    const float2 a = reinterpret_cast<const float2*>(rec + 6 * i)[0];      // records are 8-byte aligned
    const float2 b = reinterpret_cast<const float2*>(rec + 6 * i)[1];
    const float2 c = reinterpret_cast<const float2*>(rec + 6 * i)[2];
    float4* out = reinterpret_cast<float4*>(rays + i);
    out[0] = make_float4(a.x, a.y, b.x, tmin);
    out[1] = make_float4(b.y, c.x, c.y, tmax);
}

__global__ void __launch_bounds__(256) pack_records(const Ray* __restrict__ rays, long long count, float* __restrict__ rec) {
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= count) return;
    const float4 a = reinterpret_cast<const float4*>(rays + i)[0];
    const float4 b = reinterpret_cast<const float4*>(rays + i)[1];
    float2* out = reinterpret_cast<float2*>(rec + 6 * i);
    out[0] = make_float2(a.x, a.y); out[1] = make_float2(a.z, b.x); out[2] = make_float2(b.y, b.z);
}

} // namespace

void expand_ray_records(const float* dev_records, long long count, float tmin, float tmax, Ray* rays) {
    if (count <= 0) return;
    expand_records<<<unsigned((count + 255) / 256), 256>>>(dev_records, count, tmin, tmax, rays); count_launch();
    HGB_CUDA(cudaGetLastError());
}

void pack_ray_records(const Ray* rays, long long count, float* dev_records) {
    if (count <= 0) return;
    pack_records<<<unsigned((count + 255) / 256), 256>>>(rays, count, dev_records); count_launch();
    HGB_CUDA(cudaGetLastError());
}

} // namespace hagrid
