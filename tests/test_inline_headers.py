"""The public inline helpers of the compatibility headers (hagrid_b200/include/hagrid/hgb_inline.h: lookup_entry,
foreach_ref, compute_range, intersect_prim_ray, intersect_prim_cell -- what src/grid.h:78-140 and src/prims.h:262-295
offer to third-party code) compiled as HOST code with g++ and checked against the reference's golden grids and hits."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from util import Golden

ROOT = Path(__file__).resolve().parent.parent
SRC = r'''
#define HOST
#define DEVICE
#include "grid.h"
#include "ray.h"
#include "prims.h"
using namespace hagrid;
extern "C" {
void lookup_all(const Entry* entries, int shift, const int* dims, const int* voxels, int n, unsigned* out) {
    for (int i = 0; i < n; i++) out[i] = lookup_entry(entries, shift, ivec3(dims[0], dims[1], dims[2]), ivec3(voxels[3 * i], voxels[3 * i + 1], voxels[3 * i + 2]));
}
// closest hit of every ray among ALL triangles, in index order (brute force with the header's test)
void brute_force(const Tri* tris, int num_tris, const Ray* rays, int n, Hit* hits) {
    for (int i = 0; i < n; i++) {
        Ray ray = rays[i];
        Hit hit(-1, ray.tmax, 0, 0);
        for (int k = 0; k < num_tris; k++)
            if (intersect_prim_ray(tris[k], ray, k, hit)) ray.tmax = hit.t;
        hits[i] = hit;
    }
}
int sum_refs(const Cell* cells, int num_cells, const int* refs, long long* checksum) {
    int words = 0; long long sum = 0;
    for (int c = 0; c < num_cells; c++) words += foreach_ref(cells[c], refs, [&](int ref) { sum += ref; });
    *checksum = sum;
    return words;
}
int sum_small_refs(const SmallCell* cells, int num_cells, const int* refs, long long* checksum) {
    int words = 0; long long sum = 0;
    for (int c = 0; c < num_cells; c++) words += foreach_ref(cells[c], refs, [&](int ref) { sum += ref; });
    *checksum = sum;
    return words;
}
// every triangle must touch the world box of every cell that references it
int refs_touch_their_cells(const Tri* tris, const Cell* cells, int num_cells, const int* refs, const float* bmin, const float* csize) {
    int bad = 0;
    for (int c = 0; c < num_cells; c++) {
        const Cell& cell = cells[c];
        BBox box(vec3(bmin[0] + csize[0] * cell.min.x, bmin[1] + csize[1] * cell.min.y, bmin[2] + csize[2] * cell.min.z),
                 vec3(bmin[0] + csize[0] * cell.max.x, bmin[1] + csize[1] * cell.max.y, bmin[2] + csize[2] * cell.max.z));
        // grown by a hair: the build decides with the device's roundings, this test with the host's
        vec3 pad = (box.max - box.min) * 1e-4f;
        box.min = box.min - pad; box.max = box.max + pad;
        for (int i = cell.begin; i < cell.end; i++) bad += !intersect_prim_cell(tris[refs[i]], box);
    }
    return bad;
}
void range_of(const int* dims, const float* gmin, const float* gmax, const float* omin, const float* omax, int* out) {
    Range r = compute_range(ivec3(dims[0], dims[1], dims[2]), BBox(vec3(gmin[0], gmin[1], gmin[2]), vec3(gmax[0], gmax[1], gmax[2])),
                            BBox(vec3(omin[0], omin[1], omin[2]), vec3(omax[0], omax[1], omax[2])));
    out[0] = r.lx; out[1] = r.ly; out[2] = r.lz; out[3] = r.hx; out[4] = r.hy; out[5] = r.hz; out[6] = r.size();
}
}
'''


@pytest.fixture(scope="module")
def helpers(tmp_path_factory):
    d = tmp_path_factory.mktemp("inline")
    (d / "helpers.cpp").write_text(SRC)
    subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-I", str(ROOT / "hagrid_b200/include/hagrid"),
                    "-I/usr/local/cuda/include", str(d / "helpers.cpp"), "-o", str(d / "libhelpers.so")], check=True)
    return C.CDLL(str(d / "libhelpers.so"))


def p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("name", ["cornell32", "soup800"])
def test_lookup_entry_finds_the_cell_that_owns_the_voxel(helpers, name):
    g = Golden(name)
    info, entries, cells, refs = g.stage["expand"]
    dims = np.array(info["dims"], np.int32)
    rng = np.random.default_rng(3)
    vdims = dims << info["shift"]
    voxels = np.ascontiguousarray(rng.integers(0, vdims, size=(20000, 3)).astype(np.int32))
    out = np.empty(len(voxels), np.uint32)
    helpers.lookup_all(p(entries), info["shift"], p(dims), p(voxels), len(voxels), p(out))
    assert out.max() < info["num_cells"]
    # before expansion a cell's box is exactly the set of voxels it owns: check against the flatten stage
    info_f, entries_f, cells_f, _ = g.stage["flatten"]
    helpers.lookup_all(p(entries_f), info_f["shift"], p(dims), p(voxels), len(voxels), p(out))
    own = cells_f[out]
    assert ((voxels >= own["min"]) & (voxels < own["max"])).all()


@pytest.mark.parametrize("name", ["cornell32", "soup800"])
def test_intersect_prim_ray_brute_force_equals_the_traversal(helpers, name):
    """The closest hit over ALL triangles with the header's test = what the grid traversal of the reference found."""
    g = Golden(name)
    rays = np.ascontiguousarray(g.rays[:3000])
    hits = np.empty(len(rays), g.hits["hits_cell_ids"].dtype)
    helpers.brute_force(p(g.tris), len(g.tris), p(rays), len(rays), p(hits))
    want = g.hits["hits_cell_ids"][:3000]
    hit = want["id"] >= 0
    assert np.array_equal(hits["id"] >= 0, hit)
    close = np.abs(hits["t"][hit] - want["t"][hit]) <= 1e-5 * np.maximum(1.0, np.abs(want["t"][hit]))
    assert close.all()
    # same primitive except where two triangles are hit at (nearly) the same distance
    assert (hits["id"][hit] == want["id"][hit]).mean() > 0.98


def test_foreach_ref_and_intersect_prim_cell(helpers):
    g = Golden("soup800")
    info, entries, cells, refs = g.stage["merge"]
    checksum = C.c_longlong(0)
    words = helpers.sum_refs(p(cells), info["num_cells"], p(refs), C.byref(checksum))
    assert words == info["num_refs"] and checksum.value == int(refs.astype(np.int64).sum())
    info_c, _, small, refs_c = g.stage["compress"]
    words = helpers.sum_small_refs(p(small), info_c["num_cells"], p(refs_c), C.byref(checksum))
    assert words == info_c["num_refs"] and checksum.value == int(refs_c[refs_c >= 0].astype(np.int64).sum())
    bmin = np.array(info["bbox_min"], np.float32); bmax = np.array(info["bbox_max"], np.float32)
    csize = ((bmax - bmin) / (np.array(info["dims"]) << info["shift"])).astype(np.float32)
    assert helpers.refs_touch_their_cells(p(g.tris), p(cells), info["num_cells"], p(refs), p(bmin), p(csize)) == 0


def test_compute_range(helpers):
    out = np.zeros(7, np.int32)
    dims = np.array([4, 2, 8], np.int32)
    gmin, gmax = np.array([0, 0, 0], np.float32), np.array([4, 2, 8], np.float32)
    omin, omax = np.array([1.5, -3, 2.0], np.float32), np.array([2.5, 0.5, 100], np.float32)
    helpers.range_of(p(dims), p(gmin), p(gmax), p(omin), p(omax), p(out))
    assert list(out) == [1, 0, 2, 2, 0, 7, 2 * 1 * 6]
