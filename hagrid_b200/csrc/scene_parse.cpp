// Parallel Wavefront OBJ parser with the semantics of the reference's single-threaded loader
// (src/load_obj.cpp:78-239) for everything that reaches the tracer: positions and faces.
//
//   * the file is read once and cut into one chunk per thread at line boundaries;
//   * pass 1 (parallel): every chunk parses its lines with the reference's rules — `v` positions via strtof,
//     `f` faces via the reference's index grammar (v, v/t, v//n, v/t/n, negative = relative, at most 8 corners),
//     `vn` / `vt` are only counted (relative t / n indices are validated against the counts), g / o / s /
//     usemtl / mtllib are accepted, anything else is an error like in the reference (err_count > 0 => refused);
//   * the per-chunk vertex / normal / texcoord counts are prefix-summed, then pass 2 (parallel) resolves the
//     relative indices, validates them and writes the fan triangles (v0, v[i+1], v[i+2]) of every face to its
//     final place. Faces keep file order, which is the order load_model emits them in (src/main.cpp:251-270).
//
// Differences from the reference, all on inputs it mishandles: a position index beyond the vertices of the
// file is an error here (the reference reads out of bounds); a line of 1023 characters or more ends the parse
// like the reference's failed getline does (src/load_obj.cpp:103-105), but is reported as an error.
#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "scene_ingest.h"

namespace hagrid {

namespace {

constexpr int kMaxCorners = 8;      // Face::max_indices, src/load_obj.h
constexpr size_t kMaxLine = 1024;   // src/load_obj.cpp:102

struct RawFace {
    int v[kMaxCorners], t[kMaxCorners], n[kMaxCorners];
    int count;
    int verts_before, texs_before, norms_before;   // chunk-local counts when the face was read
};

struct Chunk {
    const char* begin; const char* end;
    std::vector<vec3> vertices;
    std::vector<RawFace> faces;
    int num_normals = 0, num_texcoords = 0;
    int first_vertex = 0, first_normal = 0, first_texcoord = 0;   // global counts before the chunk
    size_t first_tri = 0, num_tris = 0;
    std::string error;
};

// isspace / isdigit of the "C" locale (the only one the loader ever runs in), without the library call
inline bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline const char* skip_spaces(const char* p) { while (is_space(*p)) p++; return p; }

/// strtof for the numbers OBJ files hold, same value and same end pointer: plain decimals of at most 19 digits
/// with a small exponent are converted through one exact double operation (integer mantissa times or divided by
/// an exactly representable power of ten: the double is correctly rounded) and rounded to float. Rounding twice
/// can only differ from rounding once when that double sits exactly half-way between two floats; then, and for
/// everything else (inf, nan, hex floats, long mantissas, huge exponents, results outside the normal float range,
/// no digits at all), the C library decides.
float parse_float(const char* text, char** end) {
    static const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                      1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    const char* p = skip_spaces(text);
    bool negative = false;
    if (*p == '-' || *p == '+') { negative = *p == '-'; p++; }
    uint64_t mantissa = 0;
    int digits = 0, exp10 = 0;
    bool any = false;
    while (*p >= '0' && *p <= '9') {
        any = true;
        if (mantissa || *p != '0') { if (++digits > 19) return std::strtof(text, end); mantissa = mantissa * 10 + uint64_t(*p - '0'); }
        p++;
    }
    if (*p == '.') {
        p++;
        while (*p >= '0' && *p <= '9') {
            any = true;
            if (mantissa || *p != '0') { if (++digits > 19) return std::strtof(text, end); mantissa = mantissa * 10 + uint64_t(*p - '0'); }
            exp10--;
            p++;
        }
    }
    if (!any) return std::strtof(text, end);
    if (*p == 'e' || *p == 'E') {
        const char* q = p + 1;
        bool exp_negative = false;
        if (*q == '-' || *q == '+') { exp_negative = *q == '-'; q++; }
        if (*q >= '0' && *q <= '9') {
            int e = 0;
            while (*q >= '0' && *q <= '9') { if (e < 10000) e = e * 10 + (*q - '0'); q++; }
            exp10 += exp_negative ? -e : e;
            p = q;
        }
    } else if (*p == 'x' || *p == 'X') {
        return std::strtof(text, end);                       // "0x..." is a hexadecimal float for strtof
    }
    if (mantissa == 0) { *end = const_cast<char*>(p); return negative ? -0.0f : 0.0f; }
    if (mantissa >= (1ull << 53) || exp10 < -22 || exp10 > 22) return std::strtof(text, end);
    const double d = exp10 < 0 ? double(mantissa) / kPow10[-exp10] : double(mantissa) * kPow10[exp10];
    uint64_t bits;
    std::memcpy(&bits, &d, sizeof(bits));
    if ((bits & 0x1FFFFFFFull) == 0x10000000ull || d < 1.2e-38 || d > 3.4e38) return std::strtof(text, end);
    *end = const_cast<char*>(p);
    const float f = float(d);
    return negative ? -f : f;
}

/// strtol(base 10) for the indices of a face; long digit strings go to the C library
long parse_index(const char* text, char** end) {
    const char* p = skip_spaces(text);
    bool negative = false;
    if (*p == '-' || *p == '+') { negative = *p == '-'; p++; }
    if (!(*p >= '0' && *p <= '9')) return std::strtol(text, end, 10);
    long value = 0;
    int digits = 0;
    while (*p >= '0' && *p <= '9') { if (++digits > 17) return std::strtol(text, end, 10); value = value * 10 + (*p - '0'); p++; }
    *end = const_cast<char*>(p);
    return negative ? -value : value;
}

/// read_index, src/load_obj.cpp:42-76
bool read_corner(const char*& p, int& v, int& t, int& n) {
    const char* base = skip_spaces(p);
    if (!is_digit(*base) && *base != '-') return false;
    v = t = n = 0;
    char* next;
    v = int(parse_index(base, &next)); base = skip_spaces(next);
    if (*base == '/') {
        base++;
        if (*base != '/') { t = int(parse_index(base, &next)); base = next; }
        base = skip_spaces(base);
        if (*base == '/') { base++; n = int(parse_index(base, &next)); base = next; }
    }
    p = base;
    return true;
}

void parse_chunk(Chunk& c) {
    // Lines are terminated in place (the buffer is private to the parse and chunks do not overlap): the numbers of a
    // line must not run on into the next one, which is what the reference's getline into a line buffer guarantees.
    char* p = const_cast<char*>(c.begin);
    char* const chunk_end = const_cast<char*>(c.end);
    while (p < chunk_end) {
        char* eol = static_cast<char*>(std::memchr(p, '\n', size_t(chunk_end - p)));
        char* stop = eol ? eol : chunk_end;
        if (size_t(stop - p) >= kMaxLine - 1) { c.error = "line longer than 1022 characters"; return; }
        *stop = '\0';                    // the '\n', or the terminator behind the last line of the file
        const char* ptr = skip_spaces(p);
        p = eol ? eol + 1 : chunk_end;
        if (*ptr == '\0' || *ptr == '#') continue;
        {   // remove_eol, src/load_obj.cpp:22-30
            char* q = stop - 1;
            while (q > ptr && is_space(*q)) *q-- = '\0';
        }
        if (*ptr == 'v') {
            if (ptr[1] == ' ' || ptr[1] == '\t') {
                char* next;
                vec3 v;
                v.x = parse_float(ptr + 1, &next); v.y = parse_float(next, &next); v.z = parse_float(next, &next);
                c.vertices.push_back(v);
            } else if (ptr[1] == 'n') c.num_normals++;
            else if (ptr[1] == 't') c.num_texcoords++;
            else { c.error = "invalid vertex"; return; }
        } else if (*ptr == 'f' && is_space(ptr[1])) {
            RawFace f;
            f.count = 0;
            const char* q = ptr + 2;
            while (f.count < kMaxCorners && read_corner(q, f.v[f.count], f.t[f.count], f.n[f.count])) f.count++;
            if (f.count < 3) { c.error = "invalid face"; return; }
            f.verts_before = int(c.vertices.size()); f.texs_before = c.num_texcoords; f.norms_before = c.num_normals;
            c.faces.push_back(f);
        } else if ((*ptr == 'g' || *ptr == 'o' || *ptr == 's') && is_space(ptr[1])) {
        } else if ((!std::strncmp(ptr, "usemtl", 6) || !std::strncmp(ptr, "mtllib", 6)) && is_space(ptr[6])) {
        } else { c.error = std::string("unknown command ") + ptr; return; }
    }
}

/// Relative -> absolute indices and validation (src/load_obj.cpp:172-187); sizes include the dummy element 0.
bool resolve_face(const Chunk& c, const RawFace& f, int total_vertices, int* abs_v) {
    const int nv = 1 + c.first_vertex + f.verts_before, nt = 1 + c.first_texcoord + f.texs_before, nn = 1 + c.first_normal + f.norms_before;
    for (int i = 0; i < f.count; i++) {
        const int v = f.v[i] < 0 ? nv + f.v[i] : f.v[i];
        const int t = f.t[i] < 0 ? nt + f.t[i] : f.t[i];
        const int n = f.n[i] < 0 ? nn + f.n[i] : f.n[i];
        if (v <= 0 || t < 0 || n < 0 || v > total_vertices) return false;     // positive indices may point ahead in the file
        abs_v[i] = v;
    }
    return true;
}

template <typename F>
void run_parallel(int count, F&& body) {
    std::vector<std::thread> pool;
    for (int i = 1; i < count; i++) pool.emplace_back([&body, i] { body(i); });
    body(0);
    for (auto& t : pool) t.join();
}

} // namespace

bool parse_obj(const std::string& path, int threads, ObjGeometry& out) {
    out.vertices.clear(); out.indices.clear(); out.error.clear();
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { out.error = "cannot open " + path; return false; }
    struct stat info;
    if (::fstat(fd, &info) != 0 || info.st_size < 0) { ::close(fd); out.error = "cannot read " + path; return false; }
    const long size = long(info.st_size);
    if (threads <= 0) threads = int(std::thread::hardware_concurrency());
    threads = std::max(1, std::min(threads, int(size / (1 << 16)) + 1));
    // the file goes into an uninitialised buffer, every thread reading its own slice: the copy out of the page cache
    // and the page faults of the fresh buffer are the larger part of a cold sequential read
    std::unique_ptr<char[]> text(new char[size_t(size) + 1]);
    std::vector<char> slice_ok(size_t(threads), 0);
    run_parallel(threads, [&](int i) {
        size_t at = size_t(size) * size_t(i) / size_t(threads);
        const size_t stop = size_t(size) * size_t(i + 1) / size_t(threads);
        while (at < stop) {
            const ssize_t n = ::pread(fd, text.get() + at, stop - at, off_t(at));
            if (n <= 0) return;
            at += size_t(n);
        }
        slice_ok[size_t(i)] = 1;
    });
    ::close(fd);
    for (char ok : slice_ok)
        if (!ok) { out.error = "cannot read " + path; return false; }
    text[size_t(size)] = '\0';

    std::vector<Chunk> chunks;
    chunks.resize(static_cast<size_t>(threads));
    {
        const char* base = text.get();
        const char* end = base + size;
        const char* cur = base;
        for (int i = 0; i < threads; i++) {
            const char* want = i + 1 == threads ? end : base + size_t(size) * size_t(i + 1) / size_t(threads);
            if (want < cur) want = cur;
            if (want < end) {
                const char* nl = static_cast<const char*>(std::memchr(want, '\n', size_t(end - want)));
                want = nl ? nl + 1 : end;
            }
            chunks[size_t(i)].begin = cur; chunks[size_t(i)].end = want;
            cur = want;
        }
    }
    run_parallel(threads, [&](int i) { parse_chunk(chunks[size_t(i)]); });
    for (const Chunk& c : chunks)
        if (!c.error.empty()) { out.error = c.error; return false; }

    int nv = 0, nn = 0, nt = 0;
    size_t tris = 0;
    for (Chunk& c : chunks) {
        c.first_vertex = nv; c.first_normal = nn; c.first_texcoord = nt; c.first_tri = tris;
        nv += int(c.vertices.size()); nn += c.num_normals; nt += c.num_texcoords;
        for (const RawFace& f : c.faces) c.num_tris += size_t(f.count - 2);
        tris += c.num_tris;
    }
    if (tris > size_t(0x7fffffff) / 3) { out.error = "too many triangles"; return false; }
    out.vertices.resize(size_t(nv) + 1);
    out.vertices[0] = vec3(0.0f, 0.0f, 0.0f);          // dummy vertex, src/load_obj.cpp:96
    out.indices.resize(tris * 3);
    run_parallel(threads, [&](int i) {
        Chunk& c = chunks[size_t(i)];
        if (!c.vertices.empty()) std::memcpy(&out.vertices[size_t(c.first_vertex) + 1], c.vertices.data(), sizeof(vec3) * c.vertices.size());
        int* dst = out.indices.data() + 3 * c.first_tri;
        for (const RawFace& f : c.faces) {
            int v[kMaxCorners];
            if (!resolve_face(c, f, nv, v)) { c.error = "invalid indices"; return; }
            for (int k = 0; k < f.count - 2; k++) { *dst++ = v[0]; *dst++ = v[k + 1]; *dst++ = v[k + 2]; }
        }
    });
    for (const Chunk& c : chunks)
        if (!c.error.empty()) { out.error = c.error; return false; }
    return true;
}

} // namespace hagrid
