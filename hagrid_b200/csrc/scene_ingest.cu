// Triangle setup on the device (load_model, src/main.cpp:255-268).
#include "runtime.h"
#include "scene_ingest.h"

namespace hagrid {

namespace {

__global__ void __launch_bounds__(256) make_tris(const float* __restrict__ vertices, const int* __restrict__ indices,
                                                 int num_tris, Tri* __restrict__ tris) {
    const int id = blockIdx.x * 256 + threadIdx.x;
    if (id >= num_tris) return;
    const int i0 = indices[3 * id], i1 = indices[3 * id + 1], i2 = indices[3 * id + 2];
    const float v0x = vertices[3 * i0], v0y = vertices[3 * i0 + 1], v0z = vertices[3 * i0 + 2];
    const float v1x = vertices[3 * i1], v1y = vertices[3 * i1 + 1], v1z = vertices[3 * i1 + 2];
    const float v2x = vertices[3 * i2], v2y = vertices[3 * i2 + 1], v2z = vertices[3 * i2 + 2];
    // the front end is x86-64 code without FMA: every product and difference is rounded on its own
    const float e1x = __fsub_rn(v0x, v1x), e1y = __fsub_rn(v0y, v1y), e1z = __fsub_rn(v0z, v1z);
    const float e2x = __fsub_rn(v2x, v0x), e2y = __fsub_rn(v2y, v0y), e2z = __fsub_rn(v2z, v0z);
    const float nx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e1z, e2y));
    const float ny = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e1x, e2z));
    const float nz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x));
    float4* out = reinterpret_cast<float4*>(tris + id);
    out[0] = make_float4(v0x, v0y, v0z, nx);
    out[1] = make_float4(e1x, e1y, e1z, ny);
    out[2] = make_float4(e2x, e2y, e2z, nz);
}

} // namespace

void setup_triangles(const vec3* dev_vertices, const int* dev_indices, int num_tris, Tri* tris) {
    if (num_tris <= 0) return;
    make_tris<<<(num_tris + 255) / 256, 256>>>(reinterpret_cast<const float*>(dev_vertices), dev_indices, num_tris, tris);
    count_launch();
    HGB_CUDA(cudaGetLastError());
}

} // namespace hagrid
