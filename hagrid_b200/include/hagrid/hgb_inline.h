// Public inline helpers of the path, usable from host and device code: the functions third-party code finds next
// to the types in the reference's headers (src/grid.h:78-176, src/ray.h:35-47, src/prims.h:262-306). They are written
// against this library's types (hgb_types.h) in ordinary C++ arithmetic -- what the caller's compiler makes of it,
// exactly like the reference's inline code; the library's own kernels use the rounding-pinned forms in csrc/
// (device_math.cuh, tri_box.cuh) and do not include this header.
#ifndef HGB_INLINE_H
#define HGB_INLINE_H

#include <math.h>

#include "hgb_types.h"

namespace hagrid {

// ---------------------------------------------------------------- scalars (src/common.h:40-47)
/// 1 / x, and an infinity of x's sign for x == 0
HGB_HD inline float safe_rcp(float x) { return x != 0.0f ? 1.0f / x : copysignf(INFINITY, x); }
/// x with the sign of x * y
HGB_HD inline float prodsign(float x, float y) { return as<float>(as<uint32_t>(x) ^ (as<uint32_t>(y) & 0x80000000u)); }

// ---------------------------------------------------------------- voxel map (src/grid.h:78-140)
/// Inclusive range of the cells of a dims.x * dims.y * dims.z lattice over `grid_bb` that `obj_bb` touches
HGB_HD inline Range compute_range(const ivec3& dims, const BBox& grid_bb, const BBox& obj_bb) {
    const vec3 scale = vec3(dims) / grid_bb.extents();
    const vec3 lo = (obj_bb.min - grid_bb.min) * scale, hi = (obj_bb.max - grid_bb.min) * scale;
    return Range(hagrid::max(int(lo.x), 0), hagrid::max(int(lo.y), 0), hagrid::max(int(lo.z), 0),
                 hagrid::min(int(hi.x), dims.x - 1), hagrid::min(int(hi.y), dims.y - 1), hagrid::min(int(hi.z), dims.z - 1));
}

/// Index of the cell that owns `voxel` (virtual-grid coordinates): top-level word, then `log_dim` bits of every
/// coordinate per hop until a leaf. `dims` are the top-level dimensions, `shift` = Grid::shift.
HGB_HD inline uint32_t lookup_entry(const Entry* entries, int shift, const ivec3& dims, const ivec3& voxel) {
    Entry e = entries[(voxel.x >> shift) + dims.x * ((voxel.y >> shift) + dims.y * (voxel.z >> shift))];
    int depth = int(e.log_dim);
    while (e.log_dim) {
        const int bits = int(e.log_dim), down = shift - depth, mask = (1 << bits) - 1;
        const int kx = (voxel.x >> down) & mask, ky = (voxel.y >> down) & mask, kz = (voxel.z >> down) & mask;
        e = entries[e.begin + kx + ((ky + (kz << bits)) << bits)];
        depth += int(e.log_dim);
    }
    return e.begin;
}

/// Calls f(ref) for the references of a cell in array order; returns the number of reference words read
template <typename F>
HGB_HD inline int foreach_ref(const Cell& cell, const int* ref_ids, F f) {
    for (int i = cell.begin; i < cell.end; i++) f(ref_ids[i]);
    return cell.end - cell.begin;
}

/// Same for a compressed cell: the list ends at a -1 sentinel, which counts as a word read (0 for an empty cell)
template <typename F>
HGB_HD inline int foreach_ref(const SmallCell& cell, const int* ref_ids, F f) {
    if (cell.begin < 0) return 0;
    int i = cell.begin;
    for (int ref = ref_ids[i++]; ref >= 0; ref = ref_ids[i++]) f(ref);
    return i - cell.begin;
}

// ---------------------------------------------------------------- ray / triangle (src/prims.h:266-295)
/// Closest-hit update of `hit` by triangle `id`: hit.t shrinks when the triangle is hit inside (ray.tmin, ray.tmax);
/// the caller lowers ray.tmax to hit.t between calls, as the traversal does. u, v stay untouched (the reference
/// only fills them under COMPUTE_UVS, which it never defines).
HGB_HD inline bool intersect_prim_ray(const Tri& tri, const Ray& ray, int id, Hit& hit) {
    const vec3 n = tri.normal();
    const vec3 c = tri.v0 - ray.org;
    const vec3 r = cross(ray.dir, c);
    const float det = dot(n, ray.dir), abs_det = fabsf(det);
    const float u = prodsign(dot(r, tri.e2), det), v = prodsign(dot(r, tri.e1), det), w = abs_det - u - v;
    const float eps = 1e-9f;
    if (!(u >= -eps && v >= -eps && w >= -eps)) return false;
    const float t = prodsign(dot(n, c), det);
    if (!(t >= abs_det * ray.tmin && abs_det * ray.tmax > t)) return false;
    hit.t = t * (1.0f / abs_det);
    hit.id = id;
    return true;
}

// ---------------------------------------------------------------- triangle / box (src/prims.h:161-264)
namespace detail {
/// Is e x (unit axis `axis`) a separating axis? `a`, `b`: the two box-centred vertices whose projections differ,
/// `f` = |e|, `h` = box half size.
template <int axis>
HGB_HD inline bool edge_axis_separates(const vec3& h, const vec3& e, const vec3& f, const vec3& a, const vec3& b) {
    float p0, p1, rad;
    if (axis == 0)      { p0 = e.y * a.z - e.z * a.y; p1 = e.y * b.z - e.z * b.y; rad = f.z * h.y + f.y * h.z; }
    else if (axis == 1) { p0 = e.z * a.x - e.x * a.z; p1 = e.z * b.x - e.x * b.z; rad = f.z * h.x + f.x * h.z; }
    else                { p0 = e.x * a.y - e.y * a.x; p1 = e.x * b.y - e.y * b.x; rad = f.y * h.x + f.x * h.y; }
    return fminf(p0, p1) > rad || fmaxf(p0, p1) < -rad;
}
} // namespace detail

/// Does the triangle touch the box? Separating axes: the triangle's plane and the nine edge x box-axis products (no
/// bounding-box pre-test: callers only ask for boxes inside the triangle's bounds, src/build.cu:140-216).
HGB_HD inline bool intersect_prim_cell(const Tri& tri, const BBox& box) {
    const vec3 n = tri.normal();
    const vec3 first(n.x > 0 ? box.min.x : box.max.x, n.y > 0 ? box.min.y : box.max.y, n.z > 0 ? box.min.z : box.max.z);
    const vec3 last(n.x > 0 ? box.max.x : box.min.x, n.y > 0 ? box.max.y : box.min.y, n.z > 0 ? box.max.z : box.min.z);
    const float d = dot(tri.v0, n);
    if (!((dot(n, last) - d) * (dot(n, first) - d) <= 0.0f)) return false;
    const vec3 centre = (box.max + box.min) * 0.5f, h = (box.max - box.min) * 0.5f;
    const vec3 w0 = tri.v0 - centre, w1 = tri.v0 - tri.e1 - centre, w2 = tri.v0 + tri.e2 - centre;
    const vec3 e3 = tri.e1 + tri.e2;
    const vec3 f1(fabsf(tri.e1.x), fabsf(tri.e1.y), fabsf(tri.e1.z)), f2(fabsf(tri.e2.x), fabsf(tri.e2.y), fabsf(tri.e2.z));
    const vec3 f3(fabsf(e3.x), fabsf(e3.y), fabsf(e3.z));
    using namespace detail;
    if (edge_axis_separates<0>(h, tri.e1, f1, w0, w2) || edge_axis_separates<1>(h, tri.e1, f1, w0, w2) || edge_axis_separates<2>(h, tri.e1, f1, w1, w2)) return false;
    if (edge_axis_separates<0>(h, tri.e2, f2, w0, w1) || edge_axis_separates<1>(h, tri.e2, f2, w0, w1) || edge_axis_separates<2>(h, tri.e2, f2, w1, w2)) return false;
    if (edge_axis_separates<0>(h, e3, f3, w0, w2) || edge_axis_separates<1>(h, e3, f3, w0, w2) || edge_axis_separates<2>(h, e3, f3, w0, w1)) return false;
    return true;
}

// ---------------------------------------------------------------- vectorised device accessors
#ifdef __CUDACC__
/// 32-byte ray as two 16-byte loads, 16-byte hit as one store (src/ray.h:35-47)
__device__ __forceinline__ Ray load_ray(const Ray* p) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    return Ray(vec3(a.x, a.y, a.z), a.w, vec3(b.x, b.y, b.z), b.w);
}
__device__ __forceinline__ void store_hit(Hit* p, const Hit& hit) {
    *reinterpret_cast<float4*>(p) = make_float4(__int_as_float(hit.id), hit.t, hit.u, hit.v);
}
/// 48-byte triangle as three 16-byte loads (src/prims.h:297-306)
__device__ __forceinline__ Tri load_prim(const Tri* p) {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1], c = reinterpret_cast<const float4*>(p)[2];
    return Tri(vec3(a.x, a.y, a.z), a.w, vec3(b.x, b.y, b.z), b.w, vec3(c.x, c.y, c.z), c.w);
}
/// Cells as 16-byte words (src/grid.h:142-176); a compressed cell packs min.x|min.y, min.z|max.x, max.y|max.z, begin
__device__ __forceinline__ Cell load_cell(const Cell* p) {
    const int4 a = reinterpret_cast<const int4*>(p)[0], b = reinterpret_cast<const int4*>(p)[1];
    return Cell(ivec3(a.x, a.y, a.z), a.w, ivec3(b.x, b.y, b.z), b.w);
}
__device__ __forceinline__ ivec3 load_cell_min(const Cell* p) {
    const int4 a = reinterpret_cast<const int4*>(p)[0];
    return ivec3(a.x, a.y, a.z);
}
__device__ __forceinline__ void store_cell(Cell* p, const Cell& c) {
    reinterpret_cast<int4*>(p)[0] = make_int4(c.min.x, c.min.y, c.min.z, c.begin);
    reinterpret_cast<int4*>(p)[1] = make_int4(c.max.x, c.max.y, c.max.z, c.end);
}
__device__ __forceinline__ SmallCell load_cell(const SmallCell* p) {
    const uint4 w = *reinterpret_cast<const uint4*>(p);
    return SmallCell(usvec3(w.x & 0xFFFF, w.x >> 16, w.y & 0xFFFF), usvec3(w.y >> 16, w.z & 0xFFFF, w.z >> 16), int(w.w));
}
__device__ __forceinline__ void store_cell(SmallCell* p, const SmallCell& c) {
    *reinterpret_cast<uint4*>(p) = make_uint4(c.min.x | (unsigned(c.min.y) << 16), c.min.z | (unsigned(c.max.x) << 16),
                                              c.max.y | (unsigned(c.max.z) << 16), unsigned(c.begin));
}
#endif // __CUDACC__

} // namespace hagrid
#endif // HGB_INLINE_H
