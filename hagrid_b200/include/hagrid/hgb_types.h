// hagrid_b200 — public POD types and host math of the irregular-grid path.
//
// Source-compatible with what the reference's front end consumes from its
// headers: the byte layouts below are the ABI of the path
//   vec3/ivec3 12 B (src/vec.h:56-74)      Ray   32 B (src/ray.h:9-20)
//   Hit  16 B (src/ray.h:23-33)            Tri   48 B (src/prims.h:13-16)
//   BBox 32 B (src/bbox.h:10-14)           Cell  32 B (src/grid.h:23-33)
//   SmallCell 16 B (src/grid.h:36-45)      Entry  4 B (src/grid.h:12-20)
// and are pinned by static_asserts at the end of this file. Everything here is
// host/device-neutral value code; the device-side loads, the ray/triangle and
// triangle/box arithmetic live next to the kernels in hagrid_b200/csrc.
#ifndef HGB_TYPES_H
#define HGB_TYPES_H

#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

#if defined(__CUDACC__)
#  ifndef HOST
#    define HOST __host__
#  endif
#  ifndef DEVICE
#    define DEVICE __device__
#  endif
#else
#  ifndef HOST
#    define HOST
#  endif
#  ifndef DEVICE
#    define DEVICE
#  endif
#endif
#define HGB_HD HOST DEVICE

namespace hagrid {

// ---------------------------------------------------------------- scalars
// `a < b ? a : b` on purpose: the path's NaN behaviour depends on it
// (src/common.h:23-25).
template <typename T> HGB_HD inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> HGB_HD inline T max(T a, T b) { return a > b ? a : b; }
template <typename T> HGB_HD inline T clamp(T v, T lo, T hi) { return min(hi, max(lo, v)); }

/// ceil(i / j) for positive j (src/common.h:18-20)
HGB_HD inline int round_div(int i, int j) { return (i + j - 1) / j; }

/// Bit-cast between same-sized types (src/common.h:32-37)
template <typename To, typename From>
HGB_HD inline To as(From from) {
    static_assert(sizeof(To) == sizeof(From), "as<> needs equal sizes");
    union { From f; To t; } pun;
    pun.f = from;
    return pun.t;
}

/// Number of bits needed to store `v` (0 -> 0, 2..3 -> 2, 4..7 -> 3, ...) with
/// the reference's one irregular value ilog2(1) == 0: its 5-step bisection over
/// [0, 32] cannot separate 0 from 1 (src/common.h:80-93). Used as the `bits`
/// argument of the build's radix sort.
template <typename T>
HGB_HD inline int ilog2(T v) {
    unsigned long long u = (unsigned long long)v;
    if (u <= 1) return 0;
    int bits = 0;
    while (u) { bits++; u >>= 1; }
    return bits;
}

// ---------------------------------------------------------------- vectors
template <typename T>
struct tvec2 {
    T x, y;
    HGB_HD tvec2() {}
    HGB_HD tvec2(T s) : x(s), y(s) {}
    HGB_HD tvec2(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> HGB_HD explicit tvec2(const tvec2<U>& o) : x(T(o.x)), y(T(o.y)) {}
};

template <typename T>
struct tvec3 {
    // r/g/b name the same storage: the front end's MTL parser writes colours through them
    // (src/load_obj.cpp:279-301); the record stays three packed scalars.
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    HGB_HD tvec3() {}
    HGB_HD tvec3(T s) : x(s), y(s), z(s) {}
    HGB_HD tvec3(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    template <typename U> HGB_HD explicit tvec3(const tvec3<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
};

#define HGB_VEC_OP(op)                                                                                        \
    template <typename T> HGB_HD inline tvec2<T> operator op(const tvec2<T>& a, const tvec2<T>& b) {          \
        return tvec2<T>(a.x op b.x, a.y op b.y); }                                                            \
    template <typename T> HGB_HD inline tvec2<T> operator op(const tvec2<T>& a, T b) {                        \
        return tvec2<T>(a.x op b, a.y op b); }                                                                \
    template <typename T> HGB_HD inline tvec2<T> operator op(T a, const tvec2<T>& b) {                        \
        return tvec2<T>(a op b.x, a op b.y); }                                                                \
    template <typename T> HGB_HD inline tvec3<T> operator op(const tvec3<T>& a, const tvec3<T>& b) {          \
        return tvec3<T>(a.x op b.x, a.y op b.y, a.z op b.z); }                                                \
    template <typename T> HGB_HD inline tvec3<T> operator op(const tvec3<T>& a, T b) {                        \
        return tvec3<T>(a.x op b, a.y op b, a.z op b); }                                                      \
    template <typename T> HGB_HD inline tvec3<T> operator op(T a, const tvec3<T>& b) {                        \
        return tvec3<T>(a op b.x, a op b.y, a op b.z); }
HGB_VEC_OP(+) HGB_VEC_OP(-) HGB_VEC_OP(*) HGB_VEC_OP(/)
HGB_VEC_OP(<<) HGB_VEC_OP(>>) HGB_VEC_OP(&) HGB_VEC_OP(|)
#undef HGB_VEC_OP

#define HGB_VEC_ASSIGN(op)                                                                                     \
    template <typename T> HGB_HD inline tvec2<T>& operator op##=(tvec2<T>& a, const tvec2<T>& b) { a = a op b; return a; } \
    template <typename T> HGB_HD inline tvec3<T>& operator op##=(tvec3<T>& a, const tvec3<T>& b) { a = a op b; return a; } \
    template <typename T> HGB_HD inline tvec2<T>& operator op##=(tvec2<T>& a, T b) { a = a op b; return a; }   \
    template <typename T> HGB_HD inline tvec3<T>& operator op##=(tvec3<T>& a, T b) { a = a op b; return a; }
HGB_VEC_ASSIGN(+) HGB_VEC_ASSIGN(-) HGB_VEC_ASSIGN(*) HGB_VEC_ASSIGN(/)
#undef HGB_VEC_ASSIGN

template <typename T> HGB_HD inline tvec2<T> min(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(min(a.x, b.x), min(a.y, b.y)); }
template <typename T> HGB_HD inline tvec2<T> max(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(max(a.x, b.x), max(a.y, b.y)); }
template <typename T> HGB_HD inline tvec3<T> min(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
template <typename T> HGB_HD inline tvec3<T> max(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
template <typename T> HGB_HD inline tvec3<T> clamp(const tvec3<T>& v, T lo, T hi) {
    return tvec3<T>(min(max(v.x, lo), hi), min(max(v.y, lo), hi), min(max(v.z, lo), hi));
}
template <typename T> HGB_HD inline T dot(const tvec2<T>& a, const tvec2<T>& b) { return a.x * b.x + a.y * b.y; }
template <typename T> HGB_HD inline T dot(const tvec3<T>& a, const tvec3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> HGB_HD inline T length(const tvec2<T>& a) { return std::sqrt(dot(a, a)); }
template <typename T> HGB_HD inline T length(const tvec3<T>& a) { return std::sqrt(dot(a, a)); }
template <typename T> HGB_HD inline tvec2<T> normalize(const tvec2<T>& a) { return a * (T(1) / length(a)); }
template <typename T> HGB_HD inline tvec3<T> normalize(const tvec3<T>& a) { return a * (T(1) / length(a)); }
template <typename T> HGB_HD inline tvec3<T> cross(const tvec3<T>& a, const tvec3<T>& b) {
    return tvec3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

/// Rotation of `v` by `angle` around the unit vector `axis`, as the
/// quaternion sandwich q v q* (used by the viewer only, src/vec.h:105-125).
template <typename T>
HGB_HD inline tvec3<T> rotate(const tvec3<T>& v, const tvec3<T>& axis, T angle) {
    const T s = std::sin(angle / 2), w = std::cos(angle / 2);
    const tvec3<T> q = axis * s;
    // t = q x v + w v ; result = v' = (q.v) q + w t + q x t ... written out:
    const tvec3<T> t(w * v.x + q.y * v.z - q.z * v.y,
                     w * v.y - q.x * v.z + q.z * v.x,
                     w * v.z + q.x * v.y - q.y * v.x);
    const T k = -(q.x * v.x + q.y * v.y + q.z * v.z);
    return tvec3<T>(k * -q.x + t.x * w + t.y * -q.z - t.z * -q.y,
                    k * -q.y - t.x * -q.z + t.y * w + t.z * -q.x,
                    k * -q.z + t.x * -q.y - t.y * -q.x + t.z * w);
}

template <int axis, typename T> HGB_HD inline T get(const tvec2<T>& v) { return axis == 0 ? v.x : v.y; }
template <int axis, typename T> HGB_HD inline T get(const tvec3<T>& v) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

typedef tvec2<float>          vec2;
typedef tvec2<int>            ivec2;
typedef tvec2<unsigned short> usvec2;
typedef tvec3<float>          vec3;
typedef tvec3<int>            ivec3;
typedef tvec3<unsigned short> usvec3;

// ---------------------------------------------------------------- boxes
/// Axis-aligned box padded to two float4 (src/bbox.h:10-14).
struct BBox {
    vec3 min; int pad0;
    vec3 max; int pad1;

    HGB_HD BBox() {}
    HGB_HD BBox(const vec3& p) : min(p), max(p) {}
    HGB_HD BBox(const vec3& lo, const vec3& hi) : min(lo), max(hi) {}

    HGB_HD BBox& extend(const vec3& p) { min = hagrid::min(min, p); max = hagrid::max(max, p); return *this; }
    HGB_HD BBox& extend(const BBox& b) { min = hagrid::min(min, b.min); max = hagrid::max(max, b.max); return *this; }
    HGB_HD vec3 extents() const { return max - min; }
    HGB_HD vec3 center() const { return 0.5f * (max + min); }
    HGB_HD bool is_empty() const { return min.x > max.x || min.y > max.y || min.z > max.z; }
    HGB_HD float half_area() const {
        const vec3 e = hagrid::max(extents(), vec3(0.0f));
        return e.x * (e.y + e.z) + e.y * e.z;
    }
    HGB_HD static BBox empty() { return BBox(vec3(FLT_MAX), vec3(-FLT_MAX)); }
    HGB_HD static BBox full() { return BBox(vec3(-FLT_MAX), vec3(FLT_MAX)); }
};

// ---------------------------------------------------------------- rays
/// org + t * dir, t in [tmin, tmax] (src/ray.h:9-20)
struct Ray {
    vec3 org; float tmin;
    vec3 dir; float tmax;
    HGB_HD Ray() {}
    HGB_HD Ray(const vec3& o, float t0, const vec3& d, float t1) : org(o), tmin(t0), dir(d), tmax(t1) {}
};

/// Closest hit (src/ray.h:23-33)
struct Hit {
    int id; float t, u, v;
    HGB_HD Hit() {}
    HGB_HD Hit(int id_, float t_, float u_, float v_) : id(id_), t(t_), u(u_), v(v_) {}
};

// ---------------------------------------------------------------- triangles
/// v0, e1 = v0 - v1, e2 = v2 - v0, n = e1 x e2 (unnormalised), interleaved so
/// that the record is three float4 (src/prims.h:13-16, src/main.cpp:259-267).
struct Tri {
    vec3 v0; float nx;
    vec3 e1; float ny;
    vec3 e2; float nz;
    HGB_HD Tri() {}
    HGB_HD Tri(const vec3& v0_, float nx_, const vec3& e1_, float ny_, const vec3& e2_, float nz_)
        : v0(v0_), nx(nx_), e1(e1_), ny(ny_), e2(e2_), nz(nz_) {}
    HGB_HD vec3 normal() const { return vec3(nx, ny, nz); }
    HGB_HD BBox bbox() const {
        const vec3 v1 = v0 - e1, v2 = v0 + e2;
        return BBox(hagrid::min(v0, hagrid::min(v1, v2)), hagrid::max(v0, hagrid::max(v1, v2)));
    }
};

// ---------------------------------------------------------------- grid
/// Voxel-map word: 2 bits log_dim (0 = leaf), 30 bits begin (src/grid.h:12-20)
struct Entry {
    enum { LOG_DIM_BITS = 2, BEGIN_BITS = 32 - LOG_DIM_BITS };
    uint32_t log_dim : LOG_DIM_BITS;
    uint32_t begin   : BEGIN_BITS;
};

HGB_HD inline Entry make_entry(uint32_t log_dim, uint32_t begin) {
    Entry e;
    e.log_dim = log_dim;
    e.begin = begin;
    return e;
}

/// Cell box in virtual-grid units + its reference range (src/grid.h:23-33)
struct Cell {
    ivec3 min; int begin;
    ivec3 max; int end;
    HGB_HD Cell() {}
    HGB_HD Cell(const ivec3& lo, int b, const ivec3& hi, int e) : min(lo), begin(b), max(hi), end(e) {}
};

/// 16-bit cell box; references end at a -1 sentinel (src/grid.h:36-45)
struct SmallCell {
    usvec3 min;
    usvec3 max;
    int begin;
    HGB_HD SmallCell() {}
    HGB_HD SmallCell(const usvec3& lo, const usvec3& hi, int b) : min(lo), max(hi), begin(b) {}
};

/// Host-side descriptor of a grid whose arrays live on the device (src/grid.h:48-62)
struct Grid {
    Entry* entries;
    int*   ref_ids;
    Cell*  cells;
    SmallCell* small_cells;
    BBox  bbox;
    ivec3 dims;
    int num_cells;
    int num_entries;
    int num_refs;
    int shift;
    std::vector<int> offsets;
};

/// Inclusive integer box of top-level cells (src/grid.h:65-75)
struct Range {
    int lx, ly, lz, hx, hy, hz;
    HGB_HD Range() {}
    HGB_HD Range(int lx_, int ly_, int lz_, int hx_, int hy_, int hz_)
        : lx(lx_), ly(ly_), lz(lz_), hx(hx_), hy(hy_), hz(hz_) {}
    HGB_HD int size() const { return (hx - lx + 1) * (hy - ly + 1) * (hz - lz + 1); }
};

/// Cleary's resolution heuristic: dims ~ extents * cbrt(density * n / volume)
/// (src/grid.h:96-101). Host use only in this library (IEEE arithmetic).
inline ivec3 compute_grid_dims(const BBox& bb, int num_prims, float density) {
    const vec3 e = bb.extents();
    const float volume = e.x * e.y * e.z;
    const float ratio = cbrtf(density * num_prims / volume);
    return max(ivec3(1), ivec3(int(e.x * ratio), int(e.y * ratio), int(e.z * ratio)));
}

static_assert(sizeof(vec3) == 12 && sizeof(ivec3) == 12 && sizeof(usvec3) == 6, "vec3 layout");
static_assert(sizeof(Ray) == 32 && offsetof(Ray, tmin) == 12 && offsetof(Ray, dir) == 16 && offsetof(Ray, tmax) == 28, "Ray layout");
static_assert(sizeof(Hit) == 16 && offsetof(Hit, t) == 4, "Hit layout");
static_assert(sizeof(Tri) == 48 && offsetof(Tri, nx) == 12 && offsetof(Tri, e1) == 16 && offsetof(Tri, e2) == 32 && offsetof(Tri, nz) == 44, "Tri layout");
static_assert(sizeof(BBox) == 32 && offsetof(BBox, max) == 16, "BBox layout");
static_assert(sizeof(Cell) == 32 && offsetof(Cell, begin) == 12 && offsetof(Cell, max) == 16 && offsetof(Cell, end) == 28, "Cell layout");
static_assert(sizeof(SmallCell) == 16 && offsetof(SmallCell, max) == 6 && offsetof(SmallCell, begin) == 12, "SmallCell layout");
static_assert(sizeof(Entry) == 4, "Entry layout");

} // namespace hagrid

#endif // HGB_TYPES_H
