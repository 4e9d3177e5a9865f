// Device-side arithmetic contract of the path.
//
// The reference is compiled with --use_fast_math (src/CMakeLists.txt:4): FTZ,
// approximate reciprocal, and mul+add contraction chosen by the compiler.
// Every discrete decision of build and traversal (voxel truncation, tri/box
// SAT, `texit == tcell.x`, `hit.t <= texit`, the hit acceptance test) hangs on
// those roundings, so this library does NOT leave them to the optimiser: each
// float operation below is one PTX instruction with explicit rounding and
// .ftz, and the fused/unfused shape of every expression is written out by
// hand to match the SASS of the reference rebuilt for sm_100a
// (oracle/_ref/traverse_pid.sass; FMUL/FFMA/FADD sequence documented at each
// use). `.rn` on mul/add forbids ptxas from contracting them.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "hgb_types.h"

namespace hagrid {
namespace dev {

__device__ __forceinline__ float mul(float a, float b) {
    float r; asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float add(float a, float b) {
    float r; asm("add.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float sub(float a, float b) {
    float r; asm("sub.rn.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}
/// a * b + c, single rounding
__device__ __forceinline__ float fma(float a, float b, float c) {
    float r; asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
/// a * b - c * d the way nvcc contracts it: the second product is rounded, the first is fused
__device__ __forceinline__ float diff_of_products(float a, float b, float c, float d) {
    return fma(a, b, -mul(c, d));
}
/// ax*bx + ay*by + az*bz the way nvcc contracts it: fma(az, bz, fma(ax, bx, ay*by))
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return fma(az, bz, fma(ax, bx, mul(ay, by)));
}
/// MUFU.RCP (what `1.0f / x` becomes under --use_fast_math)
__device__ __forceinline__ float rcp(float x) {
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
/// MUFU.RCP based a / b (div.approx.ftz: what `a / b` becomes under --use_fast_math)
__device__ __forceinline__ float div_approx(float a, float b) {
    float r; asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}
/// float -> int, truncation toward zero, saturating, NaN -> 0 (F2I.FTZ.TRUNC)
__device__ __forceinline__ int trunc_to_int(float x) {
    int r; asm("cvt.rzi.ftz.s32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r;
}
__device__ __forceinline__ float int_to_float(int x) {
    float r; asm("cvt.rn.f32.s32 %0, %1;" : "=f"(r) : "r"(x)); return r;
}
/// x with the sign of x * y (src/common.h:45-47)
__device__ __forceinline__ float prodsign(float x, float y) {
    return __int_as_float(__float_as_int(x) ^ (__float_as_int(y) & 0x80000000));
}
/// 1 / x, +-inf for x == 0 (FTZ compare: denormals count as zero; src/common.h:40-42)
__device__ __forceinline__ float safe_rcp(float x) {
    return x != 0.0f ? rcp(x) : __int_as_float(0x7f800000 | (__float_as_int(x) & 0x80000000));
}
/// `a < b ? a : b` / `a > b ? a : b` (NaN in `a` yields b; src/common.h:23-25)
__device__ __forceinline__ float sel_min(float a, float b) { return a < b ? a : b; }
__device__ __forceinline__ float sel_max(float a, float b) { return a > b ? a : b; }

// ------------------------------------------------------------ vector loads
__device__ __forceinline__ float4 ldg4(const void* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ int4   ldg4i(const void* p) { return __ldg(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ uint4  ldg4u(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

/// Streaming accesses: every ray is read once and every hit written once, so neither should displace
/// the grid and the triangles from L1 (the loads skip L1 allocation, the stores are evict-first).
__device__ __forceinline__ float4 ldg4_stream(const void* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg4_stream(void* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct CellBox {
    int min_x, min_y, min_z, begin;
    int max_x, max_y, max_z, end;     // `end` < 0: sentinel-terminated list (SmallCell)
};

__device__ __forceinline__ CellBox load_cell_box(const Cell* cells, int id) {
    const int4 a = ldg4i(cells + id);
    const int4 b = ldg4i(reinterpret_cast<const int4*>(cells + id) + 1);
    CellBox c;
    c.min_x = a.x; c.min_y = a.y; c.min_z = a.z; c.begin = a.w;
    c.max_x = b.x; c.max_y = b.y; c.max_z = b.z; c.end = b.w;
    return c;
}

/// SmallCell word packing: {min.x | min.y<<16, min.z | max.x<<16, max.y | max.z<<16, begin}
/// (src/grid.h:162-176)
__device__ __forceinline__ CellBox load_cell_box(const SmallCell* cells, int id) {
    const uint4 w = ldg4u(cells + id);
    CellBox c;
    c.min_x = w.x & 0xFFFF; c.min_y = w.x >> 16; c.min_z = w.y & 0xFFFF;
    c.max_x = w.y >> 16;    c.max_y = w.z & 0xFFFF; c.max_z = w.z >> 16;
    c.begin = int(w.w);
    c.end = -1;
    return c;
}

__device__ __forceinline__ void store_cell(Cell* cells, int id, int min_x, int min_y, int min_z, int begin,
                                           int max_x, int max_y, int max_z, int end) {
    int4* p = reinterpret_cast<int4*>(cells + id);
    p[0] = make_int4(min_x, min_y, min_z, begin);
    p[1] = make_int4(max_x, max_y, max_z, end);
}

/// Voxel-map walk (src/grid.h:103-116): one top-level word, then `log_dim`
/// bits of each coordinate per hop until a leaf; returns the cell index.
__device__ __forceinline__ int lookup_cell(const uint32_t* __restrict__ entries, int shift,
                                           int top_x, int top_y, int vx, int vy, int vz) {
    uint32_t e = __ldg(entries + ((vx >> shift) + top_x * ((vy >> shift) + top_y * (vz >> shift))));
    uint32_t log_dim = e & 3u;
    int depth = int(log_dim);
    while (log_dim) {
        const int s = shift - depth;
        const uint32_t mask = (1u << log_dim) - 1u;
        const uint32_t kx = (uint32_t(vx) >> s) & mask;
        const uint32_t ky = (uint32_t(vy) >> s) & mask;
        const uint32_t kz = (uint32_t(vz) >> s) & mask;
        e = __ldg(entries + ((e >> 2) + kx + ((ky + (kz << log_dim)) << log_dim)));
        log_dim = e & 3u;
        depth += int(log_dim);
    }
    return int(e >> 2);
}

} // namespace dev
} // namespace hagrid
