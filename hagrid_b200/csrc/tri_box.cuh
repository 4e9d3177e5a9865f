// Triangle / axis-aligned box overlap (separating axes: the triangle plane and
// the nine edge x box-axis cross products; no bounding-box pre-test), the
// decision used by the reference's filter_refs and compute_split_masks through
// intersect_prim_cell (src/prims.h:161-264). Construction parity hangs on every
// rounding of this test, so the fused/unfused shape of each expression is spelled
// out with explicit single-instruction ops to match the reference's sm_100a SASS:
//   a*b + c*d -> fma(a, b, rn(c*d))          a*b - c*d -> fma(a, b, -rn(c*d))
//   x*x' + y*y' + z*z' -> fma(z, z', fma(x, x', rn(y*y')))
//   v - (max + min) * 0.5 -> fma(max + min, -0.5, v)
#pragma once

#include "device_math.cuh"

namespace hagrid {
namespace dev {

struct Float3 { float x, y, z; };

struct TriData {          // the 48-byte record, unpacked
    Float3 v0, e1, e2, n;
};

__device__ __forceinline__ TriData load_tri(const Tri* __restrict__ tris, int id) {
    const float4 a = ldg4(reinterpret_cast<const float4*>(tris + id) + 0);
    const float4 b = ldg4(reinterpret_cast<const float4*>(tris + id) + 1);
    const float4 c = ldg4(reinterpret_cast<const float4*>(tris + id) + 2);
    TriData t;
    t.v0 = {a.x, a.y, a.z}; t.e1 = {b.x, b.y, b.z}; t.e2 = {c.x, c.y, c.z};
    t.n = {a.w, b.w, c.w};
    return t;
}

/// Bounds of the triangle: v1 = v0 - e1, v2 = v0 + e2, select-based min/max
/// (src/prims.h:27-31 with src/common.h:23-25).
__device__ __forceinline__ void tri_bounds(const TriData& t, Float3& lo, Float3& hi) {
    const float v1x = sub(t.v0.x, t.e1.x), v1y = sub(t.v0.y, t.e1.y), v1z = sub(t.v0.z, t.e1.z);
    const float v2x = add(t.v0.x, t.e2.x), v2y = add(t.v0.y, t.e2.y), v2z = add(t.v0.z, t.e2.z);
    lo.x = sel_min(t.v0.x, sel_min(v1x, v2x)); hi.x = sel_max(t.v0.x, sel_max(v1x, v2x));
    lo.y = sel_min(t.v0.y, sel_min(v1y, v2y)); hi.y = sel_max(t.v0.y, sel_max(v1y, v2y));
    lo.z = sel_min(t.v0.z, sel_min(v1z, v2z)); hi.z = sel_max(t.v0.z, sel_max(v1z, v2z));
}

// Separating-axis tests for the axes e x (1,0,0), e x (0,1,0), e x (0,0,1);
// `a`, `b` are the two triangle vertices (box-centred) that bound the projection,
// `f` = |e|, `h` = box half size. True means "separated".
__device__ __forceinline__ bool separated_x(const Float3& h, const Float3& e, const Float3& f, const Float3& a, const Float3& b) {
    const float p0 = diff_of_products(e.y, a.z, e.z, a.y);
    const float p1 = diff_of_products(e.y, b.z, e.z, b.y);
    const float rad = fma(f.z, h.y, mul(f.y, h.z));
    return (fminf(p0, p1) > rad) | (fmaxf(p0, p1) < -rad);
}
__device__ __forceinline__ bool separated_y(const Float3& h, const Float3& e, const Float3& f, const Float3& a, const Float3& b) {
    const float p0 = diff_of_products(e.z, a.x, e.x, a.z);
    const float p1 = diff_of_products(e.z, b.x, e.x, b.z);
    const float rad = fma(f.z, h.x, mul(f.x, h.z));
    return (fminf(p0, p1) > rad) | (fmaxf(p0, p1) < -rad);
}
__device__ __forceinline__ bool separated_z(const Float3& h, const Float3& e, const Float3& f, const Float3& a, const Float3& b) {
    const float p0 = diff_of_products(e.x, a.y, e.y, a.x);
    const float p1 = diff_of_products(e.x, b.y, e.y, b.x);
    const float rad = fma(f.y, h.x, mul(f.x, h.y));
    return (fminf(p0, p1) > rad) | (fmaxf(p0, p1) < -rad);
}

__device__ __forceinline__ bool tri_overlaps_box(const TriData& t, const Float3& lo, const Float3& hi) {
    // Plane of the triangle against the two box corners extreme along n
    {
        const Float3 first = {t.n.x > 0.0f ? lo.x : hi.x, t.n.y > 0.0f ? lo.y : hi.y, t.n.z > 0.0f ? lo.z : hi.z};
        const Float3 last  = {t.n.x <= 0.0f ? lo.x : hi.x, t.n.y <= 0.0f ? lo.y : hi.y, t.n.z <= 0.0f ? lo.z : hi.z};
        const float d  = dot3(t.v0.x, t.v0.y, t.v0.z, t.n.x, t.n.y, t.n.z);
        const float d0 = sub(dot3(t.n.x, t.n.y, t.n.z, first.x, first.y, first.z), d);
        const float d1 = sub(dot3(t.n.x, t.n.y, t.n.z, last.x, last.y, last.z), d);
        if (!(mul(d1, d0) <= 0.0f)) return false;
    }
    const Float3 sum = {add(hi.x, lo.x), add(hi.y, lo.y), add(hi.z, lo.z)};
    const Float3 h = {mul(sub(hi.x, lo.x), 0.5f), mul(sub(hi.y, lo.y), 0.5f), mul(sub(hi.z, lo.z), 0.5f)};
    // Box-centred vertices: w = v - (hi + lo) * 0.5, contracted to one fma
    const Float3 w0 = {fma(sum.x, -0.5f, t.v0.x), fma(sum.y, -0.5f, t.v0.y), fma(sum.z, -0.5f, t.v0.z)};
    const Float3 w1 = {fma(sum.x, -0.5f, sub(t.v0.x, t.e1.x)), fma(sum.y, -0.5f, sub(t.v0.y, t.e1.y)),
                       fma(sum.z, -0.5f, sub(t.v0.z, t.e1.z))};
    const Float3 w2 = {fma(sum.x, -0.5f, add(t.v0.x, t.e2.x)), fma(sum.y, -0.5f, add(t.v0.y, t.e2.y)),
                       fma(sum.z, -0.5f, add(t.v0.z, t.e2.z))};

    const Float3 f1 = {fabsf(t.e1.x), fabsf(t.e1.y), fabsf(t.e1.z)};
    if (separated_x(h, t.e1, f1, w0, w2) || separated_y(h, t.e1, f1, w0, w2) || separated_z(h, t.e1, f1, w1, w2))
        return false;
    const Float3 f2 = {fabsf(t.e2.x), fabsf(t.e2.y), fabsf(t.e2.z)};
    if (separated_x(h, t.e2, f2, w0, w1) || separated_y(h, t.e2, f2, w0, w1) || separated_z(h, t.e2, f2, w1, w2))
        return false;
    const Float3 e3 = {add(t.e1.x, t.e2.x), add(t.e1.y, t.e2.y), add(t.e1.z, t.e2.z)};
    const Float3 f3 = {fabsf(e3.x), fabsf(e3.y), fabsf(e3.z)};
    if (separated_x(h, e3, f3, w0, w2) || separated_y(h, e3, f3, w0, w2) || separated_z(h, e3, f3, w0, w1))
        return false;
    return true;
}

} // namespace dev
} // namespace hagrid
