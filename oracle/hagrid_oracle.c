/* CPU restatement of the reference's irregular-grid path — TEST INFRASTRUCTURE ONLY.
 *
 * Plain sequential C following cg-saarland/hagrid's algorithms stage by stage
 * (file:line citations per function, paths relative to /root/reference). Nothing
 * in the product imports, links or calls this file: only tests/, the smoke test
 * and bench.py's cpu_baseline leg do, and only as the checker / reported baseline.
 *
 * Parity status: PINNED against outputs of the reference itself — the reference
 * rebuilt for sm_100a (oracle/build_ref.sh) was run on a B200 and its stage-by-stage
 * grids and hit buffers are committed under tests/golden/ (made by
 * tests/golden/make_golden.py). The reference has no tests or golden vectors of its
 * own.
 *
 * Arithmetic: the reference's device code is compiled with --use_fast_math. This
 * restatement reproduces what can be reproduced on a CPU — flush-to-zero (MXCSR
 * FTZ|DAZ, set by og_init) and the fused/unfused shape of every expression (fmaf
 * where the device fuses) — but not the approximate reciprocal (MUFU.RCP) and the
 * fast cbrtf, which differ from IEEE in the last bit. Hence: the integer stages
 * (flatten, expand, compress) and the SAH merge are bit-exact given the same
 * input; build_grid and traversal agree except where a decision sits within one
 * ulp of a threshold (documented tolerances in tests/).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>
#include <pmmintrin.h>

#define OG_MAX_LEVELS 32

typedef struct { int min[3]; int begin; int max[3]; int end; } og_cell;        /* src/grid.h:23-33 */
typedef struct { uint16_t min[3]; uint16_t max[3]; int begin; } og_small_cell;  /* src/grid.h:36-45 */
typedef struct { float v0[3], nx, e1[3], ny, e2[3], nz; } og_tri;               /* src/prims.h:13-16 */
typedef struct { float org[3], tmin, dir[3], tmax; } og_ray;                    /* src/ray.h:9-20 */
typedef struct { int id; float t, u, v; } og_hit;                               /* src/ray.h:23-33 */

typedef struct og_grid {
    float bbox_min[3], bbox_max[3];
    int dims[3];
    int shift;
    int num_cells, num_entries, num_refs;
    int compressed;
    int num_offsets;
    int offsets[OG_MAX_LEVELS];
    uint32_t* entries;
    og_cell* cells;
    og_small_cell* small_cells;
    int* refs;
} og_grid;

void og_init(void) {
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }
static float sel_min(float a, float b) { return a < b ? a : b; }    /* src/common.h:23-25 */
static float sel_max(float a, float b) { return a > b ? a : b; }
/* x*x' + y*y' + z*z' as the device contracts it: fma(z, z', fma(x, x', y*y')) */
static float dot3(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(az, bz, fmaf(ax, bx, ay * by)); }
/* a*b - c*d as the device contracts it */
static float dop(float a, float b, float c, float d) { return fmaf(a, b, -(c * d)); }
static float prodsign(float x, float y) {                                /* src/common.h:45-47 */
    uint32_t ux, uy; memcpy(&ux, &x, 4); memcpy(&uy, &y, 4);
    ux ^= uy & 0x80000000u; memcpy(&x, &ux, 4); return x;
}
static float safe_rcp(float x) { return x != 0 ? 1.0f / x : copysignf(INFINITY, x); }   /* src/common.h:40-42 */
static int f2i(float x) {                /* cvt.rzi.s32.f32: saturating, NaN -> 0 */
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (int)-2147483648LL;
    return (int)x;
}

og_grid* og_grid_new(void) { return (og_grid*)calloc(1, sizeof(og_grid)); }

void og_grid_free(og_grid* g) {
    if (!g) return;
    free(g->entries); free(g->cells); free(g->small_cells); free(g->refs); free(g);
}

/* Wraps host copies of a grid (e.g. a golden fixture) */
og_grid* og_grid_from_arrays(const float* bbox_min, const float* bbox_max, const int* dims, int shift,
                             int num_cells, int num_entries, int num_refs, int compressed,
                             int num_offsets, const int* offsets,
                             const uint32_t* entries, const void* cells, const int* refs) {
    og_grid* g = og_grid_new();
    memcpy(g->bbox_min, bbox_min, 12); memcpy(g->bbox_max, bbox_max, 12); memcpy(g->dims, dims, 12);
    g->shift = shift; g->num_cells = num_cells; g->num_entries = num_entries; g->num_refs = num_refs;
    g->compressed = compressed; g->num_offsets = num_offsets;
    memcpy(g->offsets, offsets, sizeof(int) * num_offsets);
    g->entries = (uint32_t*)malloc(4 * (size_t)imax(num_entries, 1)); memcpy(g->entries, entries, 4 * (size_t)num_entries);
    g->refs = (int*)malloc(4 * (size_t)imax(num_refs, 1)); memcpy(g->refs, refs, 4 * (size_t)num_refs);
    if (compressed) {
        g->small_cells = (og_small_cell*)malloc(sizeof(og_small_cell) * (size_t)imax(num_cells, 1));
        memcpy(g->small_cells, cells, sizeof(og_small_cell) * (size_t)num_cells);
    } else {
        g->cells = (og_cell*)malloc(sizeof(og_cell) * (size_t)imax(num_cells, 1));
        memcpy(g->cells, cells, sizeof(og_cell) * (size_t)num_cells);
    }
    return g;
}

/* ------------------------------------------------------------------ voxel map lookup (src/grid.h:103-116) */
static int lookup(const uint32_t* entries, int shift, int top_x, int top_y, int vx, int vy, int vz) {
    uint32_t e = entries[(vx >> shift) + top_x * ((vy >> shift) + top_y * (vz >> shift))];
    uint32_t log_dim = e & 3u;
    int d = (int)log_dim;
    while (log_dim) {
        uint32_t mask = (1u << log_dim) - 1u;
        uint32_t kx = ((uint32_t)vx >> (shift - d)) & mask, ky = ((uint32_t)vy >> (shift - d)) & mask, kz = ((uint32_t)vz >> (shift - d)) & mask;
        e = entries[(e >> 2) + kx + ((ky + (kz << log_dim)) << log_dim)];
        log_dim = e & 3u;
        d += (int)log_dim;
    }
    return (int)(e >> 2);
}

/* ------------------------------------------------------------------ ray/triangle (src/prims.h:266-295) */
static void intersect_tri(const og_tri* t, const float* org, const float* dir, float tmin, int id, og_hit* hit) {
    const float cx = t->v0[0] - org[0], cy = t->v0[1] - org[1], cz = t->v0[2] - org[2];
    const float rx = dop(dir[1], cz, dir[2], cy), ry = dop(dir[2], cx, dir[0], cz), rz = dop(dir[0], cy, dir[1], cx);
    const float det = dot3(t->nx, t->ny, t->nz, dir[0], dir[1], dir[2]);
    const float abs_det = fabsf(det);
    const float u = prodsign(dot3(rx, ry, rz, t->e2[0], t->e2[1], t->e2[2]), det);
    const float v = prodsign(dot3(rx, ry, rz, t->e1[0], t->e1[1], t->e1[2]), det);
    const float w = (abs_det - u) - v;
    const float eps = 1e-9f;
    if (u >= -eps && v >= -eps && w >= -eps) {
        const float tt = prodsign(dot3(t->nx, t->ny, t->nz, cx, cy, cz), det);
        if (tt >= abs_det * tmin && abs_det * hit->t > tt) {
            hit->t = tt * (1.0f / abs_det);     /* device: MUFU.RCP, may differ in the last bit */
            hit->id = id;
        }
    }
}

/* ------------------------------------------------------------------ traversal (src/traverse.cu:28-95) */
/* statistics only (tools/warp_sim.py): when set, the reference count of every visited cell is appended here */
static __thread short* og_record = NULL;
static __thread int* og_record_cells = NULL;
static __thread int og_record_max = 0;

static void traverse_one(const og_grid* g, const og_tri* tris, const og_ray* ray, og_hit* out, int prim_id_mode) {
    /* constants as setup_traversal computes them on the host (src/traverse.cu:97-101) */
    const int dims[3] = {g->dims[0] << g->shift, g->dims[1] << g->shift, g->dims[2] << g->shift};
    float ginv[3], csize[3];
    for (int k = 0; k < 3; k++) {
        const float ext = g->bbox_max[k] - g->bbox_min[k];
        ginv[k] = (float)dims[k] / ext;
        csize[k] = ext / (float)dims[k];
    }
    const float* org = ray->org; const float* dir = ray->dir;
    const float inv[3] = {safe_rcp(dir[0]), safe_rcp(dir[1]), safe_rcp(dir[2])};
    float t0[3], t1[3];
    for (int k = 0; k < 3; k++) {
        const float a = (g->bbox_min[k] - org[k]) * inv[k], b = (g->bbox_max[k] - org[k]) * inv[k];
        t0[k] = sel_min(a, b); t1[k] = sel_max(a, b);
    }
    const float tstart = fmaxf(fmaxf(t0[0], fmaxf(t0[1], t0[2])), ray->tmin);
    const float tend = fminf(fminf(t1[0], fminf(t1[1], t1[2])), ray->tmax);
    og_hit hit = {-1, ray->tmax, 0, 0};
    int steps = 0, visited = 0;
    if (!(tstart > tend)) {
        int voxel[3];
        for (int k = 0; k < 3; k++)
            voxel[k] = imin(dims[k] - 1, imax(0, f2i((fmaf(dir[k], tstart, org[k]) - g->bbox_min[k]) * ginv[k])));
        for (;;) {
            const int cell_id = lookup(g->entries, g->shift, g->dims[0], g->dims[1], voxel[0], voxel[1], voxel[2]);
            int cmin[3], cmax[3], begin, end;
            if (g->compressed) {
                const og_small_cell* c = g->small_cells + cell_id;
                for (int k = 0; k < 3; k++) { cmin[k] = c->min[k]; cmax[k] = c->max[k]; }
                begin = c->begin; end = -1;
            } else {
                const og_cell* c = g->cells + cell_id;
                for (int k = 0; k < 3; k++) { cmin[k] = c->min[k]; cmax[k] = c->max[k]; }
                begin = c->begin; end = c->end;
            }
            if (og_record_cells && visited < og_record_max) og_record_cells[visited] = cell_id;
            if (og_record && visited < og_record_max) og_record[visited] = (short)(g->compressed ? -2 : imin(end - begin, 32767));
            visited++;
            int point[3]; float tc[3];
            for (int k = 0; k < 3; k++) {
                point[k] = dir[k] >= 0.0f ? cmax[k] : cmin[k];
                tc[k] = (fmaf((float)point[k], csize[k], g->bbox_min[k]) - org[k]) * inv[k];
            }
            const float texit = fminf(tc[0], fminf(tc[1], tc[2]));
            for (int k = 0; k < 3; k++) {
                const int exit_voxel = f2i((fmaf(dir[k], texit, org[k]) - g->bbox_min[k]) * ginv[k]);
                const int next = texit == tc[k] ? point[k] + (dir[k] >= 0.0f ? 0 : -1) : exit_voxel;
                voxel[k] = dir[k] >= 0.0f ? imax(next, voxel[k]) : imin(next, voxel[k]);
            }
            if (g->compressed) {            /* sentinel-terminated list (src/grid.h:130-140) */
                int cur = begin;
                int ref = cur >= 0 ? g->refs[cur++] : -1;
                while (ref >= 0) {
                    intersect_tri(tris + ref, org, dir, ray->tmin, ref, &hit);
                    ref = g->refs[cur++];
                }
                steps += 1 + (cur - begin);
            } else {                        /* counted list (src/grid.h:118-128) */
                for (int cur = begin; cur < end; cur++) intersect_tri(tris + g->refs[cur], org, dir, ray->tmin, g->refs[cur], &hit);
                steps += 1 + (end - begin);
            }
            if (hit.t <= texit || voxel[0] < 0 || voxel[0] >= dims[0] || voxel[1] < 0 || voxel[1] >= dims[1] ||
                voxel[2] < 0 || voxel[2] >= dims[2])
                break;
        }
    }
    if (prim_id_mode != 1) hit.id = steps;  /* src/traverse.cu:93 */
    if (prim_id_mode == 2) hit.u = (float)visited;   /* statistics for DESIGN.md: cells visited (not a reference output) */
    *out = hit;
}

/* mode 0: Hit.id = step count (reference verbatim), 1: primitive id, 2: mode 0 plus Hit.u = cells visited. `threads` worker
 * threads pull chunks of 256 rays from a shared counter (pthreads; used by bench.py's
 * cpu_baseline leg to occupy all host cores). */
typedef struct {
    const og_grid* g; const og_tri* tris; const og_ray* rays; og_hit* hits; int num_rays, mode;
    volatile int next;
} og_job;

static void* traverse_worker(void* arg) {
    og_job* job = (og_job*)arg;
    og_init();
    for (;;) {
        const int begin = __sync_fetch_and_add(&job->next, 256);
        if (begin >= job->num_rays) break;
        const int end = imin(begin + 256, job->num_rays);
        for (int i = begin; i < end; i++) traverse_one(job->g, job->tris, job->rays + i, job->hits + i, job->mode);
    }
    return NULL;
}

/* statistics only: counts[i * max_steps + k] = references in the k-th cell ray i visits, -1 after the last one */
void og_traverse_record(const og_grid* g, const og_tri* tris, const og_ray* rays, short* counts, int max_steps, int num_rays) {
    og_init();
    for (int i = 0; i < num_rays; i++) {
        og_hit h;
        for (int k = 0; k < max_steps; k++) counts[(size_t)i * max_steps + k] = -1;
        og_record = counts + (size_t)i * max_steps; og_record_max = max_steps;
        traverse_one(g, tris, rays + i, &h, 1);
    }
    og_record = NULL;
}

/* statistics only: cells[i * max_steps + k] = index of the k-th cell ray i visits, -1 after the last one */
void og_traverse_record_cells(const og_grid* g, const og_tri* tris, const og_ray* rays, int* cells, int max_steps, int num_rays) {
    og_init();
    for (int i = 0; i < num_rays; i++) {
        og_hit h;
        for (int k = 0; k < max_steps; k++) cells[(size_t)i * max_steps + k] = -1;
        og_record_cells = cells + (size_t)i * max_steps; og_record_max = max_steps;
        traverse_one(g, tris, rays + i, &h, 1);
    }
    og_record_cells = NULL;
}

void og_traverse(const og_grid* g, const og_tri* tris, const og_ray* rays, og_hit* hits, int num_rays, int mode, int threads) {
    og_job job = {g, tris, rays, hits, num_rays, mode, 0};
    if (threads <= 1) { traverse_worker(&job); return; }
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    for (int t = 0; t < threads; t++) pthread_create(&tid[t], NULL, traverse_worker, &job);
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
}

/* ------------------------------------------------------------------ triangle / box (src/prims.h:161-264) */
static void tri_bounds(const og_tri* t, float* lo, float* hi) {         /* src/prims.h:27-31 */
    for (int k = 0; k < 3; k++) {
        const float v1 = t->v0[k] - t->e1[k], v2 = t->v0[k] + t->e2[k];
        lo[k] = sel_min(t->v0[k], sel_min(v1, v2));
        hi[k] = sel_max(t->v0[k], sel_max(v1, v2));
    }
}

static int separated(int axis, const float* h, const float* e, const float* f, const float* a, const float* b) {
    float p0, p1, rad;
    if (axis == 0)      { p0 = dop(e[1], a[2], e[2], a[1]); p1 = dop(e[1], b[2], e[2], b[1]); rad = fmaf(f[2], h[1], f[1] * h[2]); }
    else if (axis == 1) { p0 = dop(e[2], a[0], e[0], a[2]); p1 = dop(e[2], b[0], e[0], b[2]); rad = fmaf(f[2], h[0], f[0] * h[2]); }
    else                { p0 = dop(e[0], a[1], e[1], a[0]); p1 = dop(e[0], b[1], e[1], b[0]); rad = fmaf(f[1], h[0], f[0] * h[1]); }
    return fminf(p0, p1) > rad || fmaxf(p0, p1) < -rad;
}

static int tri_overlaps_box(const og_tri* t, const float* lo, const float* hi) {
    const float n[3] = {t->nx, t->ny, t->nz};
    float first[3], last[3];
    for (int k = 0; k < 3; k++) { first[k] = n[k] > 0 ? lo[k] : hi[k]; last[k] = n[k] <= 0 ? lo[k] : hi[k]; }
    const float d = dot3(t->v0[0], t->v0[1], t->v0[2], n[0], n[1], n[2]);
    const float d0 = dot3(n[0], n[1], n[2], first[0], first[1], first[2]) - d;
    const float d1 = dot3(n[0], n[1], n[2], last[0], last[1], last[2]) - d;
    if (!(d1 * d0 <= 0.0f)) return 0;
    float h[3], w0[3], w1[3], w2[3], f1[3], f2[3], e3[3], f3[3];
    for (int k = 0; k < 3; k++) {
        const float sum = hi[k] + lo[k];
        h[k] = (hi[k] - lo[k]) * 0.5f;
        w0[k] = fmaf(sum, -0.5f, t->v0[k]);
        w1[k] = fmaf(sum, -0.5f, t->v0[k] - t->e1[k]);
        w2[k] = fmaf(sum, -0.5f, t->v0[k] + t->e2[k]);
        f1[k] = fabsf(t->e1[k]); f2[k] = fabsf(t->e2[k]);
        e3[k] = t->e1[k] + t->e2[k]; f3[k] = fabsf(e3[k]);
    }
    if (separated(0, h, t->e1, f1, w0, w2) || separated(1, h, t->e1, f1, w0, w2) || separated(2, h, t->e1, f1, w1, w2)) return 0;
    if (separated(0, h, t->e2, f2, w0, w1) || separated(1, h, t->e2, f2, w0, w1) || separated(2, h, t->e2, f2, w1, w2)) return 0;
    if (separated(0, h, e3, f3, w0, w2) || separated(1, h, e3, f3, w0, w2) || separated(2, h, e3, f3, w0, w1)) return 0;
    return 1;
}

/* ------------------------------------------------------------------ build_grid (src/build.cu:470-760) */
typedef struct { int* ref_ids; int* cell_ids; int num_refs, num_kept; og_cell* cells; uint32_t* entries; int num_cells; } og_level;

static void cell_box(const og_grid* g, const float* csize, const og_cell* c, float* lo, float* hi) {
    for (int k = 0; k < 3; k++) {
        lo[k] = fmaf((float)c->min[k], csize[k], g->bbox_min[k]);
        hi[k] = fmaf((float)c->max[k], csize[k], g->bbox_min[k]);
    }
}

/* CUB DevicePartition::Flagged semantics: selected items first, in order;
 * rejected items at the rear in REVERSE order. */
static void partition_flagged(const int* in, const int* flags, int n, int* out) {
    int front = 0, back = n - 1;
    for (int i = 0; i < n; i++) { if (flags[i]) out[front++] = in[i]; else out[back--] = in[i]; }
}

static int cmp_pair(const void* a, const void* b) {
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

og_grid* og_build(const og_tri* tris, int n, float top_density, float snd_density) {
    og_init();
    og_grid* g = og_grid_new();
    /* scene box, Cleary dims rounded up to even, box grown by 0.1 % (src/build.cu:723-737, src/grid.h:96-101) */
    float lo[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, hi[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    float* tlo = (float*)malloc(sizeof(float) * 3 * (size_t)n); float* thi = (float*)malloc(sizeof(float) * 3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        tri_bounds(tris + i, tlo + 3 * i, thi + 3 * i);
        for (int k = 0; k < 3; k++) { lo[k] = sel_min(lo[k], tlo[3 * i + k]); hi[k] = sel_max(hi[k], thi[3 * i + k]); }
    }
    float ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    {   /* outside the reference's domain: a flat scene box (zero volume, src/grid.h:96-101 divides by it) gets a
           thickness of 0.1 % of its largest extent — same rule as hagrid_b200/csrc/grid_build.cu */
        const float widest = fmaxf(ext[0], fmaxf(ext[1], ext[2]));
        const float pad = (widest > 0.0f ? widest : 1.0f) * 0.0005f;
        for (int k = 0; k < 3; k++)
            if (!(ext[k] > 0.0f)) { lo[k] -= pad; hi[k] += pad; ext[k] = hi[k] - lo[k]; }
    }
    {
        const float volume = ext[0] * ext[1] * ext[2];
        const float ratio = cbrtf(top_density * n / volume);
        for (int k = 0; k < 3; k++) {
            int d = imax(1, (int)(ext[k] * ratio));
            g->dims[k] = d + (d & 1);
        }
    }
    for (int k = 0; k < 3; k++) { g->bbox_min[k] = lo[k] - ext[k] * 0.001f; g->bbox_max[k] = hi[k] + ext[k] * 0.001f; }
    const int* dims = g->dims;
    const int num_top = dims[0] * dims[1] * dims[2];

    /* reference emission, primitive-major, x fastest (src/build.cu:57-136, src/grid.h:84-93) */
    float inv[3];
    for (int k = 0; k < 3; k++) inv[k] = (float)dims[k] * (1.0f / (g->bbox_max[k] - g->bbox_min[k]));   /* device: I2F * MUFU.RCP */
    int* range = (int*)malloc(sizeof(int) * 6 * (size_t)n);
    size_t total = 0;
    for (int i = 0; i < n; i++) {
        int* r = range + 6 * i;
        for (int k = 0; k < 3; k++) {
            r[k] = imax(f2i((tlo[3 * i + k] - g->bbox_min[k]) * inv[k]), 0);
            r[3 + k] = imin(f2i((thi[3 * i + k] - g->bbox_min[k]) * inv[k]), dims[k] - 1);
        }
        total += (size_t)imax(0, (r[3] - r[0] + 1) * (r[4] - r[1] + 1) * (r[5] - r[2] + 1));
    }
    og_level* levels = (og_level*)calloc(OG_MAX_LEVELS, sizeof(og_level));
    int num_levels = 0;
    og_level* L = &levels[num_levels++];
    L->num_refs = (int)total; L->num_kept = (int)total;
    L->ref_ids = (int*)malloc(4 * (total + 1)); L->cell_ids = (int*)malloc(4 * (total + 1));
    int* refs_per_cell = (int*)calloc((size_t)num_top, 4);
    {
        size_t cur = 0;
        for (int i = 0; i < n; i++) {
            const int* r = range + 6 * i;
            for (int z = r[2]; z <= r[5]; z++) for (int y = r[1]; y <= r[4]; y++) for (int x = r[0]; x <= r[3]; x++) {
                const int cell = x + dims[0] * (y + dims[1] * z);
                L->ref_ids[cur] = i; L->cell_ids[cur] = cell; cur++;
                refs_per_cell[cell]++;
            }
        }
    }
    /* per-top-cell octree depth (src/build.cu:256-270) */
    int* log_dims = (int*)malloc(4 * (size_t)num_top);
    int shift = 0;
    {
        float e[3];
        for (int k = 0; k < 3; k++) e[k] = (g->bbox_max[k] - g->bbox_min[k]) * (1.0f / (float)dims[k]);
        const float volume = (e[0] * e[1]) * e[2];
        for (int i = 0; i < num_top; i++) {
            const float ratio = cbrtf((snd_density * (float)refs_per_cell[i]) * (1.0f / volume));
            const int m = imax(imax(1, f2i(e[0] * ratio)), imax(imax(1, f2i(e[1] * ratio)), imax(1, f2i(e[2] * ratio))));
            int d = 0; while ((1 << d) < m) d++;
            log_dims[i] = d;
            shift = imax(shift, d);
        }
    }
    float csize[3];
    for (int k = 0; k < 3; k++) csize[k] = (g->bbox_max[k] - g->bbox_min[k]) / (float)(dims[k] << shift);   /* host IEEE (src/build.cu:509) */

    /* top cells + exact filter (src/build.cu:332-351, 140-158) */
    L->num_cells = num_top;
    L->cells = (og_cell*)malloc(sizeof(og_cell) * (size_t)num_top);
    L->entries = (uint32_t*)calloc((size_t)num_top + 1, 4);
    for (int i = 0; i < num_top; i++) {
        og_cell* c = L->cells + i;
        const int x = i % dims[0], y = (i / dims[0]) % dims[1], z = i / (dims[0] * dims[1]);
        c->min[0] = x << shift; c->min[1] = y << shift; c->min[2] = z << shift;
        for (int k = 0; k < 3; k++) c->max[k] = c->min[k] + (1 << shift);
        c->begin = c->end = 0;
    }
    for (int i = 0; i < L->num_refs; i++) {
        float blo[3], bhi[3];
        cell_box(g, csize, L->cells + L->cell_ids[i], blo, bhi);
        if (!tri_overlaps_box(tris + L->ref_ids[i], blo, bhi)) { L->ref_ids[i] = -1; L->cell_ids[i] = -1; }
    }

    /* octree levels (build_iter, src/build.cu:528-619) */
    for (;;) {
        L = &levels[num_levels - 1];
        const int nr = L->num_refs, nc = L->num_cells;
        /* compute_dims + update_log_dims */
        for (int i = 0; i < nr; i++) {
            const int c = L->cell_ids[i];
            if (c < 0) continue;
            const int* m = L->cells[c].min;
            const int top = (m[0] >> shift) + dims[0] * ((m[1] >> shift) + dims[1] * (m[2] >> shift));
            L->entries[c] = (uint32_t)imin(log_dims[top], 1);
        }
        for (int i = 0; i < num_top; i++) log_dims[i] = imax(0, log_dims[i] - 1);
        int* kept = (int*)malloc(4 * ((size_t)nr + 1));
        for (int i = 0; i < nr; i++) kept[i] = L->cell_ids[i] >= 0 && (L->entries[L->cell_ids[i]] & 3u) == 0;
        int num_new_cells = 0;
        for (int i = 0; i < nc; i++) {
            const uint32_t ld = L->entries[i] & 3u;
            L->entries[i] = ld | ((uint32_t)(ld ? num_new_cells : i) << 2);
            if (ld) num_new_cells += 8;
        }
        int* pr = (int*)malloc(4 * ((size_t)nr + 1)); int* pc = (int*)malloc(4 * ((size_t)nr + 1));
        partition_flagged(L->ref_ids, kept, nr, pr);
        partition_flagged(L->cell_ids, kept, nr, pc);
        int num_kept = 0; for (int i = 0; i < nr; i++) num_kept += kept[i];
        free(kept); free(L->ref_ids); free(L->cell_ids);
        L->ref_ids = pr; L->cell_ids = pc; L->num_kept = num_kept;
        if (num_new_cells == 0) break;

        /* split masks (src/build.cu:160-216) and child references (src/build.cu:219-243) */
        const int num_split = nr - num_kept;
        uint8_t* masks = (uint8_t*)malloc((size_t)num_split + 1);
        size_t num_new_refs = 0;
        for (int s = 0; s < num_split; s++) {
            const int c = pc[num_kept + s];
            int mask = 0;
            if (c >= 0) {
                const og_tri* t = tris + pr[num_kept + s];
                float cmin[3], cmax[3], mid[3], blo[3], bhi[3];
                cell_box(g, csize, L->cells + c, cmin, cmax);
                for (int k = 0; k < 3; k++) mid[k] = (cmin[k] + cmax[k]) * 0.5f;
                tri_bounds(t, blo, bhi);
                mask = 0xFF;
                if (blo[0] > cmax[0] || bhi[0] < cmin[0]) mask = 0;
                if (blo[0] > mid[0]) mask &= 0xAA;
                if (bhi[0] < mid[0]) mask &= 0x55;
                if (blo[1] > cmax[1] || bhi[1] < cmin[1]) mask = 0;
                if (blo[1] > mid[1]) mask &= 0xCC;
                if (bhi[1] < mid[1]) mask &= 0x33;
                if (blo[2] > cmax[2] || bhi[2] < cmin[2]) mask = 0;
                if (blo[2] > mid[2]) mask &= 0xF0;
                if (bhi[2] < mid[2]) mask &= 0x0F;
                for (int i = 0; i < 8; i++) {
                    if (!(mask & (1 << i))) continue;
                    float olo[3], ohi[3];
                    for (int k = 0; k < 3; k++) { olo[k] = (i >> k) & 1 ? mid[k] : cmin[k]; ohi[k] = (i >> k) & 1 ? cmax[k] : mid[k]; }
                    if (!tri_overlaps_box(t, olo, ohi)) mask &= ~(1 << i);
                }
            }
            masks[s] = (uint8_t)mask;
            num_new_refs += (size_t)__builtin_popcount(mask);
        }
        og_level* N = &levels[num_levels++];
        N->num_refs = N->num_kept = (int)num_new_refs;
        N->ref_ids = (int*)malloc(4 * (num_new_refs + 1)); N->cell_ids = (int*)malloc(4 * (num_new_refs + 1));
        {
            size_t cur = 0;
            for (int s = 0; s < num_split; s++)
                for (int i = 0; i < 8; i++)
                    if (masks[s] & (1 << i)) {
                        N->ref_ids[cur] = pr[num_kept + s];
                        N->cell_ids[cur] = (int)(L->entries[pc[num_kept + s]] >> 2) + i;
                        cur++;
                    }
        }
        free(masks);
        /* child cells (src/build.cu:354-383) */
        N->num_cells = num_new_cells;
        N->cells = (og_cell*)malloc(sizeof(og_cell) * (size_t)num_new_cells);
        N->entries = (uint32_t*)calloc((size_t)num_new_cells + 1, 4);
        for (int c = 0; c < nc; c++) {
            if (!(L->entries[c] & 3u)) continue;
            const og_cell* p = L->cells + c;
            const int inc = (p->max[0] - p->min[0]) >> 1, first = (int)(L->entries[c] >> 2);
            for (int i = 0; i < 8; i++) {
                og_cell* q = N->cells + first + i;
                q->min[0] = p->min[0] + (i & 1) * inc; q->min[1] = p->min[1] + ((i >> 1) & 1) * inc; q->min[2] = p->min[2] + (i >> 2) * inc;
                for (int k = 0; k < 3; k++) q->max[k] = q->min[k] + inc;
                q->begin = q->end = 0;
            }
        }
    }
    free(log_dims); free(refs_per_cell); free(range); free(tlo); free(thi);

    /* concat_levels (src/build.cu:621-716) */
    int total_refs = 0, total_cells = 0;
    for (int l = 0; l < num_levels; l++) { total_refs += levels[l].num_kept; total_cells += levels[l].num_cells; }
    int* start_cell = (int*)malloc(4 * ((size_t)total_cells + 1));
    int num_leaves = 0;
    for (int l = 0, off = 0; l < num_levels; off += levels[l].num_cells, l++)
        for (int c = 0; c < levels[l].num_cells; c++) { start_cell[off + c] = num_leaves; num_leaves += (levels[l].entries[c] & 3u) == 0; }
    start_cell[total_cells] = num_leaves;
    g->cells = (og_cell*)malloc(sizeof(og_cell) * (size_t)imax(num_leaves, 1));
    g->entries = (uint32_t*)malloc(4 * (size_t)total_cells);
    int64_t* pairs = (int64_t*)malloc(8 * (size_t)imax(total_refs, 1));
    for (int l = 0, off = 0, roff = 0; l < num_levels; l++) {
        const og_level* V = &levels[l];
        for (int c = 0; c < V->num_cells; c++) {
            const uint32_t e = V->entries[c];
            if ((e & 3u) == 0) {
                g->cells[start_cell[off + c]] = V->cells[c];
                g->entries[off + c] = (uint32_t)start_cell[off + (int)(e >> 2)] << 2;
            } else {
                g->entries[off + c] = (((e >> 2) + (uint32_t)(off + V->num_cells)) << 2) | (e & 3u);
            }
        }
        /* key = final cell, then position in the concatenated array: a stable sort by cell */
        for (int i = 0; i < V->num_kept; i++)
            pairs[roff + i] = ((int64_t)start_cell[off + V->cell_ids[i]] << 32) | (uint32_t)(roff + i);
        off += V->num_cells; roff += V->num_kept;
    }
    int* all_refs = (int*)malloc(4 * (size_t)imax(total_refs, 1));
    for (int l = 0, roff = 0; l < num_levels; roff += levels[l].num_kept, l++)
        memcpy(all_refs + roff, levels[l].ref_ids, 4 * (size_t)levels[l].num_kept);
    qsort(pairs, (size_t)total_refs, 8, cmp_pair);
    g->refs = (int*)malloc(4 * (size_t)imax(total_refs, 1));
    for (int i = 0; i < total_refs; i++) {
        const int cell = (int)(pairs[i] >> 32);
        g->refs[i] = all_refs[(uint32_t)pairs[i]];
        if (i == 0 || (int)(pairs[i - 1] >> 32) != cell) g->cells[cell].begin = i;     /* src/build.cu:453-468 */
        g->cells[cell].end = i + 1;
    }
    if (total_refs > 0) g->cells[(int)(pairs[0] >> 32)].begin = 0;
    free(pairs); free(all_refs); free(start_cell);
    g->shift = num_levels - 1;
    g->num_cells = num_leaves; g->num_entries = total_cells; g->num_refs = total_refs;
    g->num_offsets = num_levels;
    for (int l = 0, off = 0; l < num_levels; l++) { off += levels[l].num_cells; g->offsets[l] = off; }
    for (int l = 0; l < num_levels; l++) { free(levels[l].ref_ids); free(levels[l].cell_ids); free(levels[l].cells); free(levels[l].entries); }
    free(levels);
    return g;
}

/* ------------------------------------------------------------------ merge_grid (src/merge.cu:21-377) */
static int union_size(const int* p0, int c0, const int* p1, int c1) {     /* src/merge.cu:58-69 */
    int i = 0, j = 0, c = 0;
    while (i < c0 && j < c1) { const int a = p0[i], b = p1[j]; i += a <= b; j += a >= b; c++; }
    return c + (c1 - j) + (c0 - i);
}

static void merge_lists(const int* p0, int c0, const int* p1, int c1, int* q) {   /* src/merge.cu:72-88 */
    int i = 0, j = 0;
    while (i < c0 && j < c1) { const int a = p0[i], b = p1[j]; *q++ = a < b ? a : b; i += a <= b; j += a >= b; }
    while (i < c0) *q++ = p0[i++];
    while (j < c1) *q++ = p1[j++];
}

static void merge_pass(og_grid* g, int axis, int empty_mask, const float* cs) {
    const int nc = g->num_cells, shift = g->shift;
    const int vd[3] = {g->dims[0] << shift, g->dims[1] << shift, g->dims[2] << shift};
    const int a1x = (axis + 1) % 3, a2x = (axis + 2) % 3;
    int* counts = (int*)malloc(4 * ((size_t)nc + 1)); int* nexts = (int*)malloc(4 * ((size_t)nc + 1)); int* prevs = (int*)malloc(4 * ((size_t)nc + 1));
    for (int i = 0; i < nc; i++) prevs[i] = -1;
    for (int id = 0; id < nc; id++) {                                             /* src/merge.cu:92-143 */
        const og_cell* c1 = g->cells + id;
        const int n1 = c1->end - c1->begin;
        int count = -(n1 + 1), next_id = -1;
        const int pos = c1->min[axis];
        const int shifted = (pos >> shift) & empty_mask, on_top = !(pos & ((1 << shift) - 1));
        if ((!shifted || !on_top) && c1->max[axis] < vd[axis]) {
            int v[3] = {c1->min[0], c1->min[1], c1->min[2]};
            v[axis] = c1->max[axis];
            next_id = lookup(g->entries, shift, g->dims[0], g->dims[1], v[0], v[1], v[2]);
            const og_cell* c2 = g->cells + next_id;
            if (c1->max[axis] == c2->min[axis] && c1->min[a1x] == c2->min[a1x] && c1->min[a2x] == c2->min[a2x] &&
                c1->max[a1x] == c2->max[a1x] && c1->max[a2x] == c2->max[a2x]) {
                const int n2 = c2->end - c2->begin;
                float e1[3], e2[3], A1, A2, A;
                for (int k = 0; k < 3; k++) { e1[k] = (float)(c1->max[k] - c1->min[k]) * cs[k]; e2[k] = (float)(c2->max[k] - c2->min[k]) * cs[k]; }
                /* per-axis rounding of the half areas, as compiled (see grid_merge.cu / oracle/_ref/merge.sass) */
                if (axis == 0) {
                    const float s = e1[1] + e1[2], p = e1[1] * e1[2];
                    A1 = fmaf(e1[0], s, p); A2 = fmaf(e2[0], s, p); A = (A1 + A2) - p;
                } else if (axis == 1) {
                    A1 = fmaf(e1[1], e1[2], e1[0] * (e1[1] + e1[2])); A2 = fmaf(e2[1], e1[2], e1[0] * (e2[1] + e1[2]));
                    A = fmaf(-e1[2], e1[0], A1 + A2);
                } else {
                    A1 = fmaf(e1[1], e1[2], e1[0] * (e1[1] + e1[2])); A2 = fmaf(e1[1], e2[2], e1[0] * (e1[1] + e2[2]));
                    A = fmaf(-e1[0], e1[1], A1 + A2);
                }
                const float apart = fmaf(A1, (float)n1 + 1.0f, A2 * ((float)n2 + 1.0f));
                if (A * ((float)imax(n1, n2) + 1.0f) <= apart) {
                    const int n = union_size(g->refs + c1->begin, n1, g->refs + c2->begin, n2);
                    if (A * ((float)n + 1.0f) <= apart) count = n;
                }
            }
        }
        counts[id] = count;
        next_id = count >= 0 ? next_id : -1;
        nexts[id] = next_id;
        if (next_id >= 0) prevs[next_id] = id;
    }
    int* flags = (int*)calloc((size_t)nc + 1, 4);
    for (int id = 0; id < nc; id++) {                                             /* src/merge.cu:146-170 */
        if (prevs[id] >= 0) continue;
        flags[id] = 1;
        int k = 1;
        for (int nx = nexts[id]; nx >= 0; nx = nexts[nx], k++) flags[nx] = k % 2 ? 0 : 1;
    }
    int* new_ids = (int*)malloc(4 * ((size_t)nc + 1));
    int new_cells_n = 0; size_t new_refs_n = 0;
    for (int id = 0; id < nc; id++) if (flags[id]) { new_cells_n++; new_refs_n += (size_t)(counts[id] >= 0 ? counts[id] : -(counts[id] + 1)); }
    og_cell* out_cells = (og_cell*)malloc(sizeof(og_cell) * (size_t)imax(new_cells_n, 1));
    int* out_refs = (int*)malloc(4 * (size_t)(new_refs_n + 1));
    int ci = 0, ri = 0;
    for (int id = 0; id < nc; id++) {                                             /* src/merge.cu:190-278 */
        if (!flags[id]) continue;
        const og_cell* c = g->cells + id;
        og_cell o = *c;
        o.begin = ri;
        new_ids[id] = ci;
        if (counts[id] >= 0) {
            const int nid = nexts[id];
            const og_cell* d = g->cells + nid;
            new_ids[nid] = ci;
            for (int k = 0; k < 3; k++) { o.min[k] = imin(c->min[k], d->min[k]); o.max[k] = imax(c->max[k], d->max[k]); }
            if (d->end > d->begin) merge_lists(g->refs + c->begin, c->end - c->begin, g->refs + d->begin, d->end - d->begin, out_refs + ri);
            else memcpy(out_refs + ri, g->refs + c->begin, 4 * (size_t)(c->end - c->begin));
            ri += counts[id];
        } else {
            memcpy(out_refs + ri, g->refs + c->begin, 4 * (size_t)(c->end - c->begin));
            ri += c->end - c->begin;
        }
        o.end = ri;
        out_cells[ci++] = o;
    }
    for (int i = 0; i < g->num_entries; i++)                                      /* src/merge.cu:281-290 */
        if ((g->entries[i] & 3u) == 0) g->entries[i] = (uint32_t)new_ids[g->entries[i] >> 2] << 2;
    free(g->cells); free(g->refs);
    g->cells = out_cells; g->refs = out_refs; g->num_cells = new_cells_n; g->num_refs = ri;
    free(counts); free(nexts); free(prevs); free(flags); free(new_ids);
}

void og_merge(og_grid* g, float alpha) {
    og_init();
    float cs[3];
    for (int k = 0; k < 3; k++) cs[k] = (g->bbox_max[k] - g->bbox_min[k]) / (float)(g->dims[k] << g->shift);
    if (alpha > 0) {
        int before, round = 0;
        do {
            before = g->num_cells;
            const int mask = round > 3 ? 0 : (1 << (round + 1)) - 1;
            merge_pass(g, 0, mask, cs); merge_pass(g, 1, mask, cs); merge_pass(g, 2, mask, cs);
            round++;
        } while (g->num_cells < alpha * before);
    }
}

/* ------------------------------------------------------------------ flatten_grid (src/flatten.cu:9-175) */
void og_flatten(og_grid* g) {
    const int shift = g->shift;
    uint32_t* E = g->entries;
    int* depth = (int*)calloc((size_t)g->num_entries + 1, 4);
    for (int level = shift; level >= 0; level--) {
        const int first = level > 0 ? g->offsets[level - 1] : 0, last = g->offsets[level];
        for (int i = first; i < last; i++) {                       /* collapse_entries, then compute_depths */
            if (!(E[i] & 3u)) continue;
            const uint32_t* kid = E + (E[i] >> 2);
            int same = 1; for (int k = 1; k < 8; k++) same &= kid[k] == kid[0];
            if (same) E[i] = kid[0];
        }
        for (int i = first; i < last; i++) {
            int d = 0;
            if (E[i] & 3u) { const int* kd = depth + (E[i] >> 2); for (int k = 0; k < 8; k++) d = imax(d, kd[k]); d++; }
            depth[i] = d;
        }
    }
    int* start = (int*)calloc((size_t)g->num_entries + 1, 4);
    int group_off[OG_MAX_LEVELS] = {0};
    int total = g->offsets[0];
    for (int level = 0; level < shift; level += 3) {
        const int first = level > 0 ? g->offsets[level - 1] : 0, last = g->offsets[level];
        int run = 0;
        for (int i = first; i < last; i++) { start[i] = run; run += depth[i] > 0 ? 1 << (imin(depth[i], 3) * 3) : 0; }
        group_off[level] = total; total += run;
    }
    uint32_t* out = (uint32_t*)malloc(4 * (size_t)total);
    for (int i = 0; i < g->offsets[0]; i++)
        out[i] = (E[i] & 3u) ? ((uint32_t)(g->offsets[0] + start[i]) << 2) | (uint32_t)imin(depth[i], 3) : E[i];
    int new_offsets[OG_MAX_LEVELS], nn = 0;
    for (int level = 0; level < shift; level += 3) {
        const int first = level > 0 ? g->offsets[level - 1] : 0, last = g->offsets[level];
        const int next_offset = level + 3 < shift ? group_off[level + 3] : 0;
        for (int id = first; id < last; id++) {
            const int d = imin(depth[id], 3);
            if (d == 0) continue;
            for (int i = 0; i < (1 << (3 * d)); i++) {
                uint32_t e = E[id];
                int x = 0, y = 0, z = 0, at = id;
                for (int lv = d - 1; lv >= 0; lv--) {
                    const int digit = (i >> (3 * lv)) & 7;
                    x |= (digit & 1) << lv; y |= ((digit >> 1) & 1) << lv; z |= (digit >> 2) << lv;
                    if (e & 3u) { at = (int)(e >> 2) + digit; e = E[at]; }
                }
                if (e & 3u) e = ((uint32_t)(next_offset + start[at]) << 2) | (uint32_t)imin(depth[at], 3);
                out[group_off[level] + start[id] + x + ((y + (z << d)) << d)] = e;
            }
        }
        new_offsets[nn++] = group_off[level];
    }
    new_offsets[nn++] = total;
    free(g->entries); free(depth); free(start);
    g->entries = out; g->num_entries = total; g->num_offsets = nn;
    memcpy(g->offsets, new_offsets, sizeof(int) * (size_t)nn);
}

/* ------------------------------------------------------------------ expand_grid (src/expand.cu:11-225) */
static int contains_all(const int* own, int n, const int* sub, int m) {      /* src/expand.cu:21-36 */
    if (m > n) return 0;
    if (m == 0) return 1;
    int i = 0, j = 0;
    do { const int a = own[i], b = sub[j]; if (b < a) return 0; j += a == b; i++; } while (i < n && j < m);
    return j == m;
}

static int face_growth(const og_grid* g, const og_cell* cells, const og_cell* cell, int axis, int dir, int* keep_going) {
    const int shift = g->shift;
    const int vd[3] = {g->dims[0] << shift, g->dims[1] << shift, g->dims[2] << shift};
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    if (dir ? cell->max[axis] >= vd[axis] : cell->min[axis] <= 0) return 0;
    int d = dir ? vd[axis] : -vd[axis], limit = d, step2 = vd[a2];
    int i = cell->min[a1], j = cell->min[a2];
    for (;;) {
        int v[3];
        v[axis] = dir ? cell->max[axis] : cell->min[axis] - 1; v[a1] = i; v[a2] = j;
        const og_cell* nx = cells + lookup(g->entries, shift, g->dims[0], g->dims[1], v[0], v[1], v[2]);
        if (dir) { limit = imin(limit, nx->max[axis] - cell->max[axis]); d = imin(d, limit); }
        else     { limit = imax(limit, nx->min[axis] - cell->min[axis]); d = imax(d, limit); }
        if (!contains_all(g->refs + cell->begin, cell->end - cell->begin, g->refs + nx->begin, nx->end - nx->begin)) { d = 0; break; }
        const int step1 = nx->max[a1] - i;
        step2 = imin(step2, nx->max[a2] - j);
        i += step1;
        if (i >= cell->max[a1]) { i = cell->min[a1]; j += step2; step2 = vd[a2]; if (j >= cell->max[a2]) break; }
    }
    *keep_going |= d == limit;
    return d;
}

void og_expand(og_grid* g, int iters) {
    if (iters == 0) return;
    const int nc = g->num_cells;
    og_cell* cur = g->cells;
    og_cell* other = (og_cell*)malloc(sizeof(og_cell) * (size_t)imax(nc, 1));
    memset(other, 0xCD, sizeof(og_cell) * (size_t)imax(nc, 1));
    int* flags = (int*)malloc(4 * (size_t)imax(nc, 1));
    for (int i = 0; i < nc; i++) flags[i] = -1;
    for (int it = 0; it < iters; it++)
        for (int axis = 0; axis < 3; axis++) {
            for (int id = 0; id < nc; id++) {
                if (!(flags[id] & (1 << axis))) continue;          /* skipped cells are NOT copied (src/expand.cu:154-155) */
                og_cell c = cur[id];
                int keep = 0;
                const int low = face_growth(g, cur, &c, axis, 0, &keep), high = face_growth(g, cur, &c, axis, 1, &keep);
                c.min[axis] += low; c.max[axis] += high;
                flags[id] = (keep ? 1 << axis : 0) | (flags[id] & ~(1 << axis));
                other[id] = c;
            }
            og_cell* t = cur; cur = other; other = t;
        }
    g->cells = cur;
    free(other); free(flags);
}

/* ------------------------------------------------------------------ compress_grid (src/compress.cu:6-63) */
int og_compress(og_grid* g) {
    for (int k = 0; k < 3; k++) if ((g->dims[k] << g->shift) >= (1 << 16)) return 0;
    const int nc = g->num_cells;
    size_t words = 0;
    for (int i = 0; i < nc; i++) { const int n = g->cells[i].end - g->cells[i].begin; words += (size_t)(n > 0 ? n + 1 : 0); }
    og_small_cell* sc = (og_small_cell*)malloc(sizeof(og_small_cell) * (size_t)imax(nc, 1));
    int* refs = (int*)malloc(4 * (words + 1));
    int at = 0;
    for (int i = 0; i < nc; i++) {
        const og_cell* c = g->cells + i;
        const int n = c->end - c->begin;
        for (int k = 0; k < 3; k++) { sc[i].min[k] = (uint16_t)c->min[k]; sc[i].max[k] = (uint16_t)c->max[k]; }
        sc[i].begin = n > 0 ? at : -1;
        if (n > 0) { memcpy(refs + at, g->refs + c->begin, 4 * (size_t)n); refs[at + n] = -1; at += n + 1; }
    }
    free(g->cells); free(g->refs);
    g->cells = NULL; g->small_cells = sc; g->refs = refs; g->num_refs = at; g->compressed = 1;
    return 1;
}

/* ------------------------------------------------------------------ front-end host functions (src/main.cpp:42-111)
 * Pinned against the reference's own code: tests/golden/frontend.npz is produced by oracle/ref_frontend.cpp,
 * which includes src/main.cpp unmodified (tests/golden/make_frontend_golden.py). */
static void normalize3(float* v) {                                             /* src/vec.h: a * (1 / length(a)) */
    const float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const float inv = 1.0f / len;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}
static void cross3(const float* a, const float* b, float* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

/* gen_camera, src/main.cpp:42-50; cam = eye, right, up, dir (member order of Camera, src/main.cpp:18-23) */
void og_gen_camera(const float* eye, const float* center, const float* up, float fov, float ratio, float* cam) {
    const float f = tanf(M_PI * fov / 360);
    float* right = cam + 3; float* upv = cam + 6; float* dir = cam + 9;
    for (int k = 0; k < 3; k++) { cam[k] = eye[k]; dir[k] = center[k] - eye[k]; }
    normalize3(dir);
    cross3(dir, up, right); normalize3(right);
    const float fr = f * ratio;
    for (int k = 0; k < 3; k++) right[k] *= fr;
    cross3(right, dir, upv); normalize3(upv);
    for (int k = 0; k < 3; k++) upv[k] *= f;
}

/* gen_rays, src/main.cpp:52-66 */
void og_gen_rays(const float* cam, float clip, int w, int h, og_ray* rays) {
    const float* eye = cam; const float* right = cam + 3; const float* up = cam + 6; const float* dir = cam + 9;
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        const float kx = 2 * x / (float)w - 1;
        const float ky = 1 - 2 * y / (float)h;
        og_ray* r = rays + (size_t)y * w + x;
        for (int k = 0; k < 3; k++) {
            r->org[k] = eye[k];
            r->dir[k] = (dir[k] + right[k] * kx) + up[k] * ky;
        }
        r->tmin = 0.0f; r->tmax = clip;
    }
}

/* update_surface<mode> + gradient, src/main.cpp:68-111, on a tightly packed BGRA image; mode 0 depth, 1 grey, 2 heat */
void og_update_surface(int mode, const og_hit* hits, float clip, int w, int h, uint8_t* bgra) {
    static const float g[5][3] = {{0, 0, 255}, {0, 255, 255}, {0, 128, 0}, {255, 255, 0}, {255, 0, 0}};
    for (size_t i = 0; i < (size_t)w * h; i++) {
        uint8_t* px = bgra + 4 * i;
        if (mode == 0) {
            const uint8_t c = (uint8_t)(int)(255.0f * hits[i].t / clip);
            px[0] = px[1] = px[2] = c;
        } else if (mode == 1) {
            const uint8_t c = (uint8_t)imin(255, hits[i].id);
            px[0] = px[1] = px[2] = c;
        } else {
            const float s = 1.0f / 5;
            const float k = imin(100, hits[i].id) / 100.0f;
            const int a = imin(4, (int)(k * 5)), b = imin(4, a + 1);
            const float t = (k - a * s) / s;
            float c[3];
            for (int q = 0; q < 3; q++) c[q] = (1.0f - t) * g[a][q] + t * g[b][q];
            px[0] = (uint8_t)(int)c[2]; px[1] = (uint8_t)(int)c[1]; px[2] = (uint8_t)(int)c[0];
        }
        px[3] = 255;
    }
}

/* ---- second-wave rays (config C5). The reference has no such stage (its front end stops at gen_rays,
 * src/main.cpp:52-66): the operation is DEFINED by include/hagrid_b200.h (hgb_generate_bounce_rays) and restated
 * here operation by operation in IEEE single precision (no fmaf, sqrtf and '/' correctly rounded), so the device
 * kernel (hagrid_b200/csrc/ray_bounce.cu) has to match bit for bit. */
static uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

static float draw24(uint32_t base, uint32_t k) { return (float)(mix32(base + k) >> 8) * 0x1p-24f; }

void og_bounce_rays(const og_tri* tris, int num_tris, const og_ray* rays, const og_hit* hits, int num_rays,
                    float offset, float tmax, uint32_t seed, og_ray* out) {
    for (int i = 0; i < num_rays; i++) {
        const og_ray r = rays[i];
        const og_hit h = hits[i];
        float n[3] = {0, 0, 0}, len = 0.0f;
        if (h.id >= 0 && h.id < num_tris) {
            n[0] = tris[h.id].nx; n[1] = tris[h.id].ny; n[2] = tris[h.id].nz;
            len = sqrtf((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2]);
        }
        if (!(len > 0.0f)) { out[i] = r; continue; }
        for (int k = 0; k < 3; k++) n[k] = n[k] / len;
        const float facing = (n[0] * r.dir[0] + n[1] * r.dir[1]) + n[2] * r.dir[2];
        if (facing > 0.0f) for (int k = 0; k < 3; k++) n[k] = -n[k];
        og_ray o;
        for (int k = 0; k < 3; k++) o.org[k] = (r.org[k] + r.dir[k] * h.t) + n[k] * offset;
        const uint32_t base = mix32(seed ^ mix32((uint32_t)i));
        float dx = 0.0f, dy = 0.0f, s = 0.0f;
        for (int k = 0; k < 8; k++) {
            const float x = 2.0f * draw24(base, 2 * k) - 1.0f;
            const float y = 2.0f * draw24(base, 2 * k + 1) - 1.0f;
            const float q = x * x + y * y;
            if (q < 1.0f) { dx = x; dy = y; s = q; break; }
        }
        const float dz = sqrtf(1.0f - s);
        float t1[3], t2[3];
        if (fabsf(n[0]) > 0.9f) { t1[0] = -n[2]; t1[1] = 0.0f; t1[2] = n[0]; }
        else                    { t1[0] = 0.0f; t1[1] = n[2]; t1[2] = -n[1]; }
        const float tl = sqrtf((t1[0] * t1[0] + t1[1] * t1[1]) + t1[2] * t1[2]);
        for (int k = 0; k < 3; k++) t1[k] = t1[k] / tl;
        t2[0] = n[1] * t1[2] - n[2] * t1[1];
        t2[1] = n[2] * t1[0] - n[0] * t1[2];
        t2[2] = n[0] * t1[1] - n[1] * t1[0];
        for (int k = 0; k < 3; k++) o.dir[k] = (t1[k] * dx + t2[k] * dy) + n[k] * dz;
        o.tmin = 0.0f; o.tmax = tmax;
        out[i] = o;
    }
}
