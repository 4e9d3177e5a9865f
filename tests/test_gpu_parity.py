"""GPU tier (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against (1) the golden fixtures produced by the reference, (2) the CPU oracle on seeded
inputs, (3) the reference rebuilt for sm_100a when oracle/_ref is present, and (4)
size-independent properties at BASELINE.json's full sizes. Integer/index results are
compared bit-exactly; hit distances bit-exactly against the reference and within 1e-5
(north_star tolerance) against the CPU oracle."""
import numpy as np
import pytest

from hagrid_b200 import HIT_PRIM_ID, HIT_STEPS, RAY_DTYPE, Scene, scenes
from hagrid_b200.api import GridInfo
from util import FIXTURES, STAGES, Golden, grid_diff, t_close

pytestmark = pytest.mark.gpu

VARIANTS = (0, 1, 2, 3, 4)    # per-thread / persistent / per-thread re-tiled / automatic / tile-pulling resident warps


def run_stage(sc, stage, g):
    if stage == "build": sc.build_grid(g.top_density, g.snd_density)
    elif stage == "merge": sc.merge_grid(g.alpha)
    elif stage == "flatten": sc.flatten_grid()
    elif stage == "expand": sc.expand_grid(g.expansion)
    elif stage == "compress": assert sc.compress_grid()


def info_from_dict(d):
    gi = GridInfo()
    gi.bbox_min[:] = d["bbox_min"]; gi.bbox_max[:] = d["bbox_max"]; gi.dims[:] = d["dims"]
    gi.shift, gi.num_cells, gi.num_entries, gi.num_refs = d["shift"], d["num_cells"], d["num_entries"], d["num_refs"]
    gi.compressed = d["compressed"]; gi.num_offsets = len(d["offsets"])
    gi.offsets[:len(d["offsets"])] = d["offsets"]
    return gi


def upload(sc, stage_tuple):
    info, e, c, r = stage_tuple
    sc.upload(info_from_dict(info), e, c, r)


def dump(sc):
    gi, e, c, r = sc.download()
    return gi.as_dict(), (e, c, r)


@pytest.fixture(scope="module", params=FIXTURES)
def golden(request):
    return Golden(request.param)


def test_native_library_is_loaded(lib):
    assert lib.impl == "hagrid_b200" and lib.path.name == "libhagrid_b200.so"


def test_pipeline_matches_reference_golden(lib, golden):
    sc = Scene(golden.tris, lib=lib)
    for stage in STAGES:
        run_stage(sc, stage, golden)
        info, arrays = dump(sc)
        assert grid_diff(info, arrays, golden.stage[stage]) == [], stage
    sc.close()


@pytest.mark.parametrize("stage", STAGES[1:])
def test_each_stage_on_reference_input(lib, golden, stage):
    sc = Scene(golden.tris, lib=lib)
    upload(sc, golden.stage[STAGES[STAGES.index(stage) - 1]])
    run_stage(sc, stage, golden)
    info, arrays = dump(sc)
    assert grid_diff(info, arrays, golden.stage[stage]) == []
    sc.close()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("cells", ["cell", "small"])
def test_traversal_bit_exact_on_reference_grid(lib, golden, cells, variant):
    sc = Scene(golden.tris, lib=lib)
    upload(sc, golden.stage["expand" if cells == "cell" else "compress"])
    sc.setup_traversal()
    lib.set_option("traverse_variant", variant)
    try:
        for mode, key in ((HIT_STEPS, "steps"), (HIT_PRIM_ID, "ids")):
            got, want = sc.trace(golden.rays, mode), golden.hits[f"hits_{cells}_{key}"]
            assert np.array_equal(got["id"], want["id"])
            assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
            assert (got["u"] == 0).all() and (got["v"] == 0).all()
    finally:
        lib.set_option("traverse_variant", 3)
        sc.close()


def test_pipeline_against_cpu_oracle(lib):
    """Seeded input outside the fixtures: GPU construction vs the CPU restatement, then the
    oracle traces the GPU-built grid."""
    from oracle import oracle
    tris = scenes.small_mixed(2500, seed=21)
    sc = Scene(tris, lib=lib)
    cpu = oracle.Grid.build(tris, 0.12, 2.4)
    sc.build_grid(0.12, 2.4)
    for stage in ("merge", "flatten", "expand"):
        getattr(cpu, stage)()
        getattr(sc, stage + "_grid")()
    info, arrays = dump(sc)
    assert grid_diff(info, arrays, (cpu.info(),) + cpu.arrays()) == []
    sc.setup_traversal()
    rays = scenes.random_rays(tris, 20000, seed=4, tmax=10.0)
    gpu_ids, gpu_steps = sc.trace(rays, HIT_PRIM_ID), sc.trace(rays, HIT_STEPS)
    cpu_ids, cpu_steps = cpu.traverse(tris, rays, 1, threads=4), cpu.traverse(tris, rays, 0, threads=4)
    assert (gpu_ids["id"] == cpu_ids["id"]).mean() >= 0.999          # near-ties may flip without MUFU.RCP
    assert (gpu_steps["id"] == cpu_steps["id"]).mean() >= 0.999
    same = gpu_ids["id"] == cpu_ids["id"]
    assert t_close(gpu_ids["t"][same], cpu_ids["t"][same]).all()
    sc.close()


def test_against_reference_build(lib, ref_lib):
    """Differential test against cg-saarland/hagrid itself (rebuilt for sm_100a) on a fresh seeded
    scene: every construction stage byte-identical, all hits bit-identical, raster and random rays."""
    tris = scenes.small_mixed(20000, seed=33)
    a, b = Scene(tris, lib=ref_lib), Scene(tris, lib=lib)
    a.build_grid(0.15, 3.0); b.build_grid(0.15, 3.0)
    for stage in ("merge", "flatten", "expand"):
        getattr(a, stage + "_grid")(); getattr(b, stage + "_grid")()
        ia, aa = dump(a); ib, ab = dump(b)
        assert grid_diff(ib, ab, (ia,) + aa) == [], stage
    lo, hi = scenes.scene_bbox(tris)
    raster = scenes.primary_rays(lo - (hi - lo) * 0.6, 0.5 * (lo + hi), (0, 1, 0), 50.0, 256, 128, 50.0)
    rays = np.concatenate([raster, scenes.random_rays(tris, 200000, seed=8)])
    for compressed in (False, True):
        if compressed:
            assert a.compress_grid() and b.compress_grid()
        a.setup_traversal(); b.setup_traversal()
        for buf in (raster, rays):
            for mode in (HIT_STEPS, HIT_PRIM_ID):
                want, got = a.trace(buf, mode), b.trace(buf, mode)
                assert np.array_equal(got["id"], want["id"])
                assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    a.close(); b.close()


# ----------------------------------------------------------------------------- full-size properties
@pytest.fixture(scope="module")
def sponza(lib):
    tris = scenes.sponza262k()
    sc = Scene(tris, lib=lib)
    ms = sc.build_all(0.15, 3.0, 0.995, 3, compress=False, warmup=0, iters=1)
    sc.setup_traversal()
    yield tris, sc, float(ms[0])
    sc.close()


def test_c2_grid_invariants(sponza):
    tris, sc, _ = sponza
    gi, entries, cells, refs = sc.download()
    vd = np.array(gi.dims[:]) << gi.shift
    assert (cells["min"] >= 0).all() and (cells["max"] <= vd).all() and (cells["min"] < cells["max"]).all()
    assert int((cells["end"] - cells["begin"]).sum()) == gi.num_refs          # ranges tile the reference array
    assert refs.min() >= 0 and refs.max() < tris.shape[0]
    leaf = (entries & 3) == 0
    assert np.array_equal(np.unique(entries[leaf] >> 2), np.arange(gi.num_cells))   # every cell owns a voxel
    assert ((entries[~leaf] >> 2) < gi.num_entries).all()
    assert np.unique(refs).shape[0] == np.unique(refs[refs >= 0]).shape[0]


def test_c2_primary_rays_all_variants_agree_and_closest_hit_properties(lib, sponza):
    tris, sc, _ = sponza
    rays = scenes.default_view(tris)              # 1920 x 1080, the BASELINE.json C2 ray buffer
    try:
        results = {}
        for v in VARIANTS:
            lib.set_option("traverse_variant", v)
            results[v] = sc.trace(rays, HIT_PRIM_ID)
        base = results[0]
        for v in VARIANTS[1:]:
            assert np.array_equal(results[v]["id"], base["id"])
            assert np.array_equal(results[v]["t"].view(np.uint32), base["t"].view(np.uint32))
        lib.set_option("traverse_variant", 3)
        assert np.array_equal(sc.trace(rays, HIT_PRIM_ID)["id"], base["id"])          # idempotent
        hit = base["id"] >= 0
        assert hit.mean() > 0.9                                                       # camera is inside the atrium
        # nothing is closer than the reported hit: stopping just short of it must find nothing
        short = rays.copy()
        short["tmax"][hit] = base["t"][hit] * np.float32(0.999)
        res = sc.trace(short, HIT_PRIM_ID)
        assert (res["id"][hit] == -1).all() and np.array_equal(res["t"][hit], short["tmax"][hit])
        # and looking further must not change it
        far = rays.copy()
        far["tmax"] = rays["tmax"] * np.float32(4.0)
        res = sc.trace(far, HIT_PRIM_ID)
        assert np.array_equal(res["id"][hit], base["id"][hit]) and np.array_equal(res["t"][hit], base["t"][hit])
        # the hit point lies on the reported triangle's plane (relative to scene size)
        tr = tris[base["id"][hit]]
        p = rays["org"][hit].astype(np.float64) + rays["dir"][hit].astype(np.float64) * base["t"][hit][:, None].astype(np.float64)
        n = np.stack([tr["nx"], tr["ny"], tr["nz"]], 1).astype(np.float64)
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        assert np.abs(((p - tr["v0"]) * n).sum(1)).max() < 1e-3
    finally:
        lib.set_option("traverse_variant", 3)


def test_c3_compressed_grid_gives_same_hits(lib, sponza):
    tris, sc, _ = sponza
    rays = scenes.random_rays(tris, 1 << 20)
    want = sc.trace(rays, HIT_PRIM_ID)
    want_steps = sc.trace(rays, HIT_STEPS)
    sc2 = Scene(tris, lib=lib)
    sc2.build_all(0.15, 3.0, 0.995, 3, compress=True)
    assert sc2.info().compressed == 1
    sc2.setup_traversal()
    got = sc2.trace(rays, HIT_PRIM_ID)
    got_steps = sc2.trace(rays, HIT_STEPS)
    sc.setup_traversal()
    assert np.array_equal(got["id"], want["id"]) and np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    assert (got_steps["id"] >= want_steps["id"]).all()       # + one sentinel read per non-empty cell
    sc2.close()


def _bit_equal(got, want):
    return np.array_equal(got["id"], want["id"]) and np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))


@pytest.fixture(scope="module")
def sponza_reference(ref_lib):
    """The reference's own grid of the bench scene (C2 flags) and its hits for the bench's ray buffers."""
    tris = scenes.sponza262k()
    sc = Scene(tris, lib=ref_lib)
    sc.build_all(0.15, 3.0, 0.995, 3, compress=False)
    sc.setup_traversal()
    plain = dump(sc)
    views = {"default": scenes.default_view(tris), "long": scenes.default_view(tris, along_long_axis=True)}
    want = {k: (sc.trace(v, HIT_PRIM_ID), sc.trace(v, HIT_STEPS)) for k, v in views.items()}
    random = scenes.random_rays(tris, 1 << 22)
    want_random_plain = sc.trace(random, HIT_PRIM_ID)
    assert sc.compress_grid()
    sc.setup_traversal()
    small = dump(sc)
    want_random = (sc.trace(random, HIT_PRIM_ID), sc.trace(random, HIT_STEPS))
    sc.close()
    return {"tris": tris, "plain": plain, "small": small, "views": views, "want": want, "random": random,
            "want_random": want_random, "want_random_plain": want_random_plain}


def test_c2_headline_buffers_match_the_reference_in_every_variant(lib, sponza, sponza_reference):
    """BASELINE.json C2, the exact buffers bench.py times: grid byte-identical to the reference's, prim ids, t and step
    counts bit-identical for the default view and the long-axis view, in all five kernel selections."""
    tris, sc, _ = sponza
    R = sponza_reference
    info, arrays = dump(sc)
    assert grid_diff(info, arrays, (R["plain"][0],) + R["plain"][1]) == []
    try:
        for name, rays in R["views"].items():
            want_ids, want_steps = R["want"][name]
            assert (want_ids["id"] >= 0).mean() > 0.9
            for v in VARIANTS:
                lib.set_option("traverse_variant", v)
                assert _bit_equal(sc.trace(rays, HIT_PRIM_ID), want_ids), (name, v)
                assert _bit_equal(sc.trace(rays, HIT_STEPS), want_steps), (name, v)
    finally:
        lib.set_option("traverse_variant", 3)


def test_c2_host_buffer_frames_match_the_reference(lib, sponza, sponza_reference):
    """The e2e call of bench.py (hgb_traverse_grid_host) with page-locked and pageable buffers, several chunk sizes."""
    import torch
    tris, sc, _ = sponza
    rays = sponza_reference["views"]["default"]
    want = sponza_reference["want"]["default"][0]
    n = rays.shape[0]
    pinned_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).pin_memory()
    pinned_hits = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    try:
        for chunk in (0, 64 << 10, 1 << 20, 1 << 30):
            lib.set_option("host_frame_chunk_rays", chunk)
            pinned_hits.zero_()
            lib.check(lib.dll.hgb_traverse_grid_host(sc._h, pinned_rays.data_ptr(), pinned_hits.data_ptr(), n, HIT_PRIM_ID), "frame")
            got = pinned_hits.numpy().view(want.dtype).reshape(-1)
            assert _bit_equal(got, want), chunk
            assert _bit_equal(sc.traverse_host(rays, HIT_PRIM_ID), want), ("pageable", chunk)
    finally:
        lib.set_option("host_frame_chunk_rays", 0)


def test_c3_headline_buffer_matches_the_reference_in_every_variant(lib, sponza, sponza_reference):
    """BASELINE.json C3: 4 194 304 random rays on the compressed grid (and on the plain one), all kernel selections."""
    tris, sc, _ = sponza
    R = sponza_reference
    rays = R["random"]
    sc2 = Scene(tris, lib=lib)
    sc2.build_all(0.15, 3.0, 0.995, 3, compress=True)
    info, arrays = dump(sc2)
    assert grid_diff(info, arrays, (R["small"][0],) + R["small"][1]) == []
    sc2.setup_traversal()
    try:
        for v in VARIANTS:
            lib.set_option("traverse_variant", v)
            assert _bit_equal(sc2.trace(rays, HIT_PRIM_ID), R["want_random"][0]), v
            assert _bit_equal(sc2.trace(rays, HIT_STEPS), R["want_random"][1]), v
            assert _bit_equal(sc.trace(rays, HIT_PRIM_ID), R["want_random_plain"]), v     # no new setup call: per-scene state
    finally:
        lib.set_option("traverse_variant", 3)
    sc2.close()


def _same_grid_as_reference(lib, ref_lib, tris, td, sd, compress=False):
    a, b = Scene(tris, keep_alive=True, lib=ref_lib), Scene(tris, keep_alive=True, lib=lib)
    a.build_all(td, sd, 0.995, 3, compress); b.build_all(td, sd, 0.995, 3, compress)
    ia, aa = dump(a); ib, ab = dump(b)
    assert grid_diff(ib, ab, (ia,) + aa) == []
    return a, b, ia


def test_c4_two_million_triangle_hairball_matches_reference(lib, ref_lib):
    """BASELINE.json C4 at full size: the complete construction pipeline, byte for byte."""
    a, b, info = _same_grid_as_reference(lib, ref_lib, scenes.hairball(), 0.12, 2.4)
    assert info["num_refs"] > 10_000_000 and info["num_cells"] > 1_000_000
    a.close(); b.close()


def test_c5_san_miguel_scale_primary_and_bounce_rays_sharded_over_8_ranks(lib, ref_lib):
    """BASELINE.json C5 at full size (7.8 M triangles, 1920x1080 primary + one bounce): same grid and same
    hits as the reference; tracing the 8 rank shards of hagrid_b200.sharding separately gives the unsharded result."""
    from hagrid_b200 import sharding
    tris = scenes.sanmiguel7p8m()
    a, b, info = _same_grid_as_reference(lib, ref_lib, tris, 0.15, 3.0)
    assert 0 < info["num_refs"] < 2**31 and info["num_entries"] < 2**30      # int32 counts, Entry.begin is 30 bits
    primary = scenes.default_view(tris)
    b.setup_traversal()
    first = b.trace(primary, HIT_PRIM_ID)
    bounce = scenes.bounce_rays(tris, primary, first["id"], first["t"])
    second = b.trace(bounce, HIT_PRIM_ID)
    a.setup_traversal()
    for rays, got in ((primary, first), (bounce, second)):
        want = a.trace(rays, HIT_PRIM_ID)
        assert np.array_equal(got["id"], want["id"]) and np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    b.setup_traversal()
    world = 8
    for rays, whole in ((primary, first), (bounce, second)):
        parts = []
        for rank in range(world):
            lo, hi = sharding.shard_bounds(rays.shape[0], rank, world, sharding.raster_granule(1920))
            parts.append(b.trace(rays[lo:hi], HIT_PRIM_ID))
        joined = np.concatenate(parts)
        assert np.array_equal(joined["id"], whole["id"]) and np.array_equal(joined["t"].view(np.uint32), whole["t"].view(np.uint32))
    hit = first["id"] >= 0
    assert hit.mean() > 0.5 and (second["t"][second["id"] >= 0] > 0).all()
    a.close(); b.close()


def test_build_is_deterministic(lib):
    tris = scenes.hairball(60000, seed=2)
    a, b = Scene(tris, lib=lib), Scene(tris, keep_alive=True, lib=lib)
    a.build_all(0.12, 2.4); b.build_all(0.12, 2.4, warmup=1, iters=2)
    ia, aa = dump(a); ib, ab = dump(b)
    assert grid_diff(ib, ab, (ia,) + aa) == []
    assert a.peak_bytes() > 0
    a.close(); b.close()


def test_merge_in_one_launch_and_pass_by_pass_give_the_same_grid(lib, ref_lib):
    """merge_grid runs all its rounds in one cooperative launch on grids of up to 768 K cells and one launch per kernel and
    pass above that: the same grid either way, and the reference's, on scenes of both sizes with either path forced;
    alpha values that end the loop after one round and after many."""
    cases = [(scenes.sponza262k(), 0.15, 3.0, 0.995), (scenes.hairball(60000, seed=5), 0.12, 2.4, 0.995),
             (scenes.hairball(60000, seed=5), 0.12, 2.4, 0.5), (scenes.hairball(3000, seed=9), 0.12, 2.4, 1.0)]
    try:
        for tris, td, sd, alpha in cases:
            r = Scene(tris, lib=ref_lib)
            r.build_grid(td, sd); r.merge_grid(alpha)
            want = dump(r)
            r.close()
            for limit in (0, 1 << 30):
                lib.set_option("merge_one_launch_max_cells", limit)
                m = Scene(tris, lib=lib)
                m.build_grid(td, sd); m.merge_grid(alpha)
                info, arrays = dump(m)
                assert grid_diff(info, arrays, (want[0],) + want[1]) == [], (tris.shape[0], alpha, limit)
                m.close()
    finally:
        lib.set_option("merge_one_launch_max_cells", -1)


# ----------------------------------------------------------------------------- edge cases
def test_ragged_and_empty_ray_buffers(lib):
    g = Golden("cornell32")
    sc = Scene(g.tris, lib=lib)
    sc.build_all(g.top_density, g.snd_density)
    sc.setup_traversal()
    full = sc.trace(g.rays, HIT_PRIM_ID)
    try:
        for v in VARIANTS:
            lib.set_option("traverse_variant", v)
            for n in (1, 31, 33, 127, 129, 1000):
                part = sc.trace(g.rays[:n], HIT_PRIM_ID)
                assert np.array_equal(part["id"], full["id"][:n]) and np.array_equal(part["t"], full["t"][:n])
            sc.traverse(0, 0, 0, HIT_PRIM_ID)        # zero rays: a no-op, not an error
    finally:
        lib.set_option("traverse_variant", 3)
    sc.close()


def test_rays_missing_the_grid_and_degenerate_rays(lib):
    g = Golden("cornell32")
    sc = Scene(g.tris, lib=lib)
    sc.build_all(g.top_density, g.snd_density)
    sc.setup_traversal()
    rays = np.zeros(64, dtype=RAY_DTYPE)
    rays["org"] = (5000.0, 5000.0, 5000.0); rays["dir"] = (1.0, 0.5, 0.25); rays["tmax"] = 123.0
    rays["dir"][32:] = 0.0
    for mode in (HIT_PRIM_ID, HIT_STEPS):
        res = sc.trace(rays, mode)
        assert (res["t"] == 123.0).all()
        assert (res["id"] == (-1 if mode == HIT_PRIM_ID else 0)).all()
    sc.close()


def test_awkward_rays_match_the_reference_bit_for_bit(lib, ref_lib):
    """Rays that stress the discrete decisions of the march (src/traverse.cu:42-90): axis-aligned and zero
    direction components (safe_rcp = +-inf), origins exactly on voxel and cell planes, origins outside the
    grid, tmin > 0, short tmax, negative tmin, denormal components (flushed to zero by FTZ), huge coordinates."""
    tris = scenes.small_mixed(20000, seed=41)
    a, b = Scene(tris, lib=ref_lib), Scene(tris, lib=lib)
    a.build_all(0.15, 3.0); b.build_all(0.15, 3.0)
    info = b.info().as_dict()
    lo, hi = np.array(info["bbox_min"], np.float32), np.array(info["bbox_max"], np.float32)
    vdims = np.array(info["dims"]) << info["shift"]
    rng = np.random.default_rng(99)
    n = 60000
    rays = scenes.random_rays(tris, n, seed=17, tmax=50.0)
    k = n // 10
    axes = rng.integers(0, 3, k)
    rays["dir"][:k] = 0.0                                              # axis-aligned: two zero components
    rays["dir"][np.arange(k), axes] = rng.choice([-1.0, 1.0], k)
    rays["dir"][k:2 * k, rng.integers(0, 3)] = 0.0                     # one zero component
    rays["dir"][2 * k:3 * k, 0] = np.float32(1e-41)                    # denormal: zero under FTZ
    rays["dir"][3 * k:4 * k, 1] = -0.0
    cell = (hi - lo) / vdims                                           # origins exactly on voxel planes
    rays["org"][4 * k:5 * k] = lo + cell * rng.integers(0, vdims, (k, 3)).astype(np.float32)
    rays["org"][5 * k:6 * k] = lo + (hi - lo) * rng.choice([-0.5, 0.0, 1.0, 1.5], (k, 3)).astype(np.float32)   # on / outside the box
    rays["tmin"][6 * k:7 * k] = rng.random(k).astype(np.float32) * 3.0
    rays["tmax"][6 * k:7 * k] = rays["tmin"][6 * k:7 * k] + rng.random(k).astype(np.float32)
    rays["tmin"][7 * k:8 * k] = -5.0                                   # hits behind the origin become eligible
    rays["org"][8 * k:9 * k] *= np.float32(1e6)                        # far away, mostly missing the grid
    rays["tmax"][9 * k:] = np.float32(1e-3)
    for compress in (False, True):
        if compress:
            assert a.compress_grid() and b.compress_grid()
        a.setup_traversal()
        want = {m: a.trace(rays, m) for m in (HIT_PRIM_ID, HIT_STEPS)}
        b.setup_traversal()
        try:
            for v in VARIANTS:
                lib.set_option("traverse_variant", v)
                for m in (HIT_PRIM_ID, HIT_STEPS):
                    got = b.trace(rays, m)
                    assert np.array_equal(got["id"], want[m]["id"]), (compress, v, m)
                    assert np.array_equal(got["t"].view(np.uint32), want[m]["t"].view(np.uint32)), (compress, v, m)
        finally:
            lib.set_option("traverse_variant", 3)
    a.close(); b.close()


def test_awkward_scenes_build_like_the_reference(lib, ref_lib):
    """Degenerate and duplicated triangles, axis-aligned sheets, a wide range of triangle sizes: every
    construction stage byte-identical to the reference."""
    rng = np.random.default_rng(5)
    base = scenes.small_mixed(6000, seed=8)
    v0, v1, v2 = scenes.tri_vertices(base)
    v0, v1, v2 = v0.copy(), v1.copy(), v2.copy()
    v1[:300] = v0[:300]                                    # zero-area: two coincident corners
    v2[300:600] = v0[300:600] + (v1[300:600] - v0[300:600]) * np.float32(2.0)     # zero-area: collinear
    v0[600:900, 1] = v1[600:900, 1] = v2[600:900, 1] = np.float32(0.5)            # axis-aligned sheet
    scale = np.float32(10.0) ** rng.uniform(-3, 0, 300).astype(np.float32)        # tiny triangles
    c = (v0[900:1200] + v1[900:1200] + v2[900:1200]) / np.float32(3)
    for v in (v0, v1, v2):
        v[900:1200] = c + (v[900:1200] - c) * scale[:, None]
    tris = np.concatenate([scenes.make_tris(v0, v1, v2), base[:500]])             # + exact duplicates
    a, b = Scene(tris, lib=ref_lib), Scene(tris, lib=lib)
    a.build_grid(0.12, 2.4); b.build_grid(0.12, 2.4)
    for stage in ("build", "merge", "flatten", "expand", "compress"):
        if stage != "build":
            getattr(a, stage + "_grid")(); getattr(b, stage + "_grid")()
        ia, aa = dump(a); ib, ab = dump(b)
        assert grid_diff(ib, ab, (ia,) + aa) == [], stage
    a.close(); b.close()


def test_single_and_degenerate_triangles(lib):
    tris = scenes.make_tris([[0, 0, 0], [0, 0, 0]], [[1, 0, 0], [0, 0, 0]], [[0, 1, 0], [0, 0, 0]])   # second one is a point
    from oracle import oracle
    for t in (tris[:1], tris):
        sc = Scene(t, lib=lib)
        sc.build_all(0.12, 2.4)
        cpu = oracle.Grid.build(t, 0.12, 2.4)       # flat scene box: both sides apply the same thickness rule
        cpu.merge(); cpu.flatten(); cpu.expand()
        info, arrays = dump(sc)
        assert grid_diff(info, arrays, (cpu.info(),) + cpu.arrays()) == []
        sc.setup_traversal()
        rays = np.zeros(2, dtype=RAY_DTYPE)
        rays["org"] = [(0.25, 0.25, 1.0), (2.0, 2.0, 1.0)]; rays["dir"] = (0, 0, -1); rays["tmax"] = 10.0
        res = sc.trace(rays, HIT_PRIM_ID)
        assert res["id"][0] == 0 and abs(res["t"][0] - 1.0) < 1e-6 and res["id"][1] == -1
        sc.close()


def test_optional_stages_can_be_skipped(lib):
    """alpha <= 0 disables merging (src/merge.cu:357), iters == 0 disables expansion (src/expand.cu:201)."""
    g = Golden("soup800")
    sc = Scene(g.tris, lib=lib)
    sc.build_grid(g.top_density, g.snd_density)
    before = dump(sc)
    sc.merge_grid(0.0)
    after = dump(sc)
    assert grid_diff(after[0], after[1], (before[0],) + before[1]) == []
    sc.flatten_grid()
    before = dump(sc)
    sc.expand_grid(0)
    after = dump(sc)
    assert grid_diff(after[0], after[1], (before[0],) + before[1]) == []
    sc.setup_traversal()
    want = Golden("soup800").hits["hits_cell_ids"]
    got = sc.trace(g.rays, HIT_PRIM_ID)      # an unmerged, unexpanded grid still finds the same closest hits
    assert np.array_equal(got["id"], want["id"])
    sc.close()


def test_host_buffer_entry_point(lib):
    g = Golden("soup800")
    sc = Scene(g.tris, lib=lib)
    upload(sc, g.stage["expand"])
    sc.setup_traversal()
    got = sc.traverse_host(g.rays, HIT_PRIM_ID)
    assert np.array_equal(got["id"], g.hits["hits_cell_ids"]["id"])
    assert np.array_equal(got["t"].view(np.uint32), g.hits["hits_cell_ids"]["t"].view(np.uint32))
    sc.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_host_frames_are_pipelined_without_changing_hits(lib, sponza, pinned):
    """hgb_traverse_grid_host cuts a frame into chunks over several streams: raster frames (chunks are
    4-row tile bands), incoherent frames, ragged sizes and sizes below one chunk give the hits of one
    plain device launch, from pageable and from page-locked host buffers."""
    import torch
    tris, sc, _ = sponza
    sc.setup_traversal()          # the traversal constants are per process, like the reference's (src/traverse.cu:7-12)
    lo, hi = scenes.scene_bbox(tris)
    eye = 0.5 * (lo + hi)
    frames = [scenes.primary_rays(eye, eye + np.array([0.3, 0.0, 1.0], np.float32), (0, 1, 0), 60.0, 1280, 720, 1e4),
              scenes.random_rays(tris, 700001, seed=12),
              scenes.primary_rays(eye, eye + np.array([0.0, 0.1, 1.0], np.float32), (0, 1, 0), 60.0, 64, 8, 1e4),
              scenes.random_rays(tris, 77, seed=13)]
    for rays in frames:
        n = rays.shape[0]
        want = sc.trace(rays, HIT_PRIM_ID)
        for mode in (HIT_PRIM_ID, HIT_STEPS):
            if pinned:
                h_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).pin_memory()
                h_hits = torch.zeros((n, 4), dtype=torch.float32).pin_memory()
                lib.check(lib.dll.hgb_traverse_grid_host(sc._h, h_rays.data_ptr(), h_hits.data_ptr(), n, mode), "host frame")
                got = h_hits.numpy().view(want.dtype).reshape(-1)
            else:
                got = sc.traverse_host(rays, mode)
            ref = want if mode == HIT_PRIM_ID else sc.trace(rays, HIT_STEPS)
            assert np.array_equal(got["id"], ref["id"]) and np.array_equal(got["t"].view(np.uint32), ref["t"].view(np.uint32))


def test_a_reused_ray_buffer_may_change_its_character(lib, sponza):
    """The library remembers what kind of buffer it saw at an address (raster of camera rays or incoherent) and
    re-classifies it from kernel feedback when the contents change. Whatever it believes, the hits are the same."""
    tris, sc, _ = sponza
    sc.setup_traversal()
    raster = scenes.default_view(tris, 640, 360)
    n = raster.shape[0]
    first = sc.trace(raster, HIT_PRIM_ID)
    waves = [raster, scenes.bounce_rays(tris, raster, first["id"], first["t"]), scenes.random_rays(tris, n, seed=3), raster]
    lib.set_option("traverse_variant", 0)
    want = [sc.trace(w, HIT_PRIM_ID) for w in waves]
    lib.set_option("traverse_variant", 3)
    d_rays, d_hits = sc.device_alloc(n * 32), sc.device_alloc(n * 16)
    try:
        for _ in range(2):
            for w, expect in zip(waves, want):
                sc.to_device(d_rays, w)
                for _ in range(12):                       # long enough for the feedback to arrive and the policy to flip
                    sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID)
                    got = sc.to_host(np.empty(n, dtype=expect.dtype), d_hits)
                    assert np.array_equal(got["id"], expect["id"]) and np.array_equal(got["t"].view(np.uint32), expect["t"].view(np.uint32))
    finally:
        sc.device_free(d_rays); sc.device_free(d_hits)


TILE_HISTORY_SETTINGS = [            # tile_order (cost classes, bits), tile_split (tiles), tile_split_log, tile_split_share (%)
    (8, 256, 2, 50),                 # the defaults
    (8, 1024, 5, 0),                 # the thousand most expensive tiles as 32 single rays each
    (4, 64, 1, 0),
    (12, 256, 3, 25),
    (1, 0, 2, 50),                   # two classes, nothing in parts
    (0, 0, 2, 50),                   # off: buffer order
]


def _set_tile_history(lib, order, split, split_log, share):
    lib.set_option("tile_order", order); lib.set_option("tile_split", split)
    lib.set_option("tile_split_log", split_log); lib.set_option("tile_split_share", share)


def test_tiles_handed_out_by_their_history_give_the_same_hits(lib, sponza, sponza_reference):
    """The tile kernel times its tiles and hands them out longest first on later launches of the same buffer, the most
    expensive ones in parts (TileHistory in ray_traverse.cu). Neither may change a hit: the bench's C2 buffers, launch
    after launch under several settings, prim ids, t and step counts bit-identical to the reference's every time --
    also when the rays in the buffer change under a history made from other rays, and for a raster whose height is no
    multiple of the tile height traced by the tile kernel although it is small."""
    tris, sc, _ = sponza
    sc.setup_traversal()
    R = sponza_reference
    n = R["views"]["default"].shape[0]
    d_rays, d_hits = sc.device_alloc(n * 32), sc.device_alloc(n * 16)
    try:
        for setting in TILE_HISTORY_SETTINGS:
            _set_tile_history(lib, *setting)
            for launch in range(6):
                name = ("default", "long")[launch % 3 == 2]          # every third launch: other rays, same buffer
                sc.to_device(d_rays, R["views"][name])
                for mode, want in zip((HIT_PRIM_ID, HIT_STEPS), R["want"][name]):
                    sc.traverse(d_rays, d_hits, n, mode)
                    got = sc.to_host(np.empty(n, dtype=want.dtype), d_hits)
                    assert _bit_equal(got, want), (setting, launch, name, mode)
        # a small ragged raster through the tile kernel
        small = scenes.default_view(tris, 648, 357)
        lib.set_option("traverse_variant", 0)
        want = sc.trace(small, HIT_PRIM_ID)
        lib.set_option("traverse_variant", 4)
        m = small.shape[0]
        sc.to_device(d_rays, small)
        for setting in TILE_HISTORY_SETTINGS[:4]:
            _set_tile_history(lib, *setting)
            for launch in range(5):
                sc.traverse(d_rays, d_hits, m, HIT_PRIM_ID)
                got = sc.to_host(np.empty(m, dtype=want.dtype), d_hits)
                assert _bit_equal(got, want), (setting, launch)
        # more buffers than the library remembers (4), of different sizes, taking turns: a history is never carried
        # over to another buffer
        _set_tile_history(lib, 8, 256, 2, 0)
        frames = [scenes.default_view(tris, 648, h) for h in (96, 357, 200, 64, 300, 128)]
        lib.set_option("traverse_variant", 0)
        wants = [sc.trace(f, HIT_PRIM_ID) for f in frames]
        lib.set_option("traverse_variant", 4)
        bufs = [sc.device_alloc(f.shape[0] * 32) for f in frames]
        try:
            for b, f in zip(bufs, frames):
                sc.to_device(b, f)
            for round_ in range(3):
                for b, f, want in zip(bufs, frames, wants):
                    for launch in range(3):
                        sc.traverse(b, d_hits, f.shape[0], HIT_PRIM_ID)
                        got = sc.to_host(np.empty(f.shape[0], dtype=want.dtype), d_hits)
                        assert _bit_equal(got, want), (round_, f.shape[0], launch)
        finally:
            for b in bufs:
                sc.device_free(b)
    finally:
        lib.set_option("traverse_variant", 3)
        _set_tile_history(lib, *TILE_HISTORY_SETTINGS[0])
        sc.device_free(d_rays); sc.device_free(d_hits)


def test_tile_costs_can_be_read_back(lib, sponza):
    """hgb_tile_costs: the per-tile times the tile kernel records for its ticket list -- nothing for a buffer that was
    never traced, one entry per 32 rays afterwards, expensive tiles where the rays are long."""
    import ctypes as C
    tris, sc, _ = sponza
    sc.setup_traversal()
    rays = scenes.default_view(tris)
    n = rays.shape[0]
    tiles = (n + 31) // 32
    d_rays, d_hits = sc.device_alloc(n * 32), sc.device_alloc(n * 16)
    cost = np.zeros(tiles, dtype=np.uint16)
    try:
        sc.to_device(d_rays, rays)
        assert lib.dll.hgb_tile_costs(C.c_void_p(d_hits), n, C.c_void_p(cost.ctypes.data), tiles) == 0        # not a ray buffer it knows
        for _ in range(3):
            sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID)
        assert lib.dll.hgb_tile_costs(C.c_void_p(d_rays), n, C.c_void_p(cost.ctypes.data), tiles) == tiles
        assert (cost > 0).all() and cost.max() > 3 * np.median(cost)
        assert lib.dll.hgb_tile_costs(C.c_void_p(d_rays), n, C.c_void_p(cost.ctypes.data), 100) == 100        # capacity respected
        assert lib.dll.hgb_tile_costs(None, n, C.c_void_p(cost.ctypes.data), tiles) == 0
    finally:
        sc.device_free(d_rays); sc.device_free(d_hits)


def test_tracing_a_grid_without_its_setup_is_an_error(lib):
    """hgb_setup_traversal must follow every change of a scene's grid (the reference's call order,
    src/main.cpp:536-549; its constants are per process, src/traverse.cu:7-12): the C ABI refuses to trace a
    scene whose setup is missing or stale."""
    from hagrid_b200 import HagridError
    g = Golden("cornell32")
    a, b = Scene(g.tris, lib=lib), Scene(scenes.small_mixed(500, seed=2), lib=lib)
    a.build_all(g.top_density, g.snd_density); b.build_all(0.12, 2.4)
    a.setup_traversal()
    assert a.trace(g.rays, HIT_PRIM_ID).shape[0] == g.rays.shape[0]
    with pytest.raises(HagridError):
        b.trace(g.rays, HIT_PRIM_ID)
    a.build_all(g.top_density, g.snd_density)          # rebuilt: the old constants no longer describe it
    with pytest.raises(HagridError):
        a.trace(g.rays, HIT_PRIM_ID)
    a.setup_traversal()
    assert np.array_equal(a.trace(g.rays, HIT_PRIM_ID)["id"], g.hits["hits_cell_ids"]["id"])
    a.close(); b.close()


def test_scenes_are_independent_of_each_other(lib):
    """Traversal state is per scene: two scenes, set up once each, traced in turn (and from two host threads)
    without another hgb_setup_traversal; on a box with two GPUs the second scene lives on the other device."""
    import threading
    g = Golden("cornell32")
    other = scenes.small_mixed(20000, seed=5)
    second_device = 1 if lib.device_count() > 1 else 0
    a, b = Scene(g.tris, lib=lib), Scene(other, device=second_device, lib=lib)
    a.build_all(g.top_density, g.snd_density); b.build_all(0.15, 3.0)
    a.setup_traversal(); b.setup_traversal()
    rays_b = scenes.random_rays(other, 300000, seed=3)
    want_a, want_b = a.trace(g.rays, HIT_PRIM_ID), b.trace(rays_b, HIT_PRIM_ID)
    assert np.array_equal(want_a["id"], g.hits["hits_cell_ids"]["id"])
    for _ in range(3):
        assert _bit_equal(a.trace(g.rays, HIT_PRIM_ID), want_a)
        assert _bit_equal(b.trace(rays_b, HIT_PRIM_ID), want_b)
    failures = []

    def worker(sc, rays, want):
        for _ in range(20):
            if not _bit_equal(sc.trace(rays, HIT_PRIM_ID), want):
                failures.append(sc)
    threads = [threading.Thread(target=worker, args=(a, g.rays, want_a)), threading.Thread(target=worker, args=(b, rays_b, want_b))]
    for t in threads: t.start()
    for t in threads: t.join()
    assert not failures
    a.close(); b.close()


def test_new_triangles_invalidate_the_setup(lib):
    """hgb_scene_set_tris replaces the array the grid's references index: tracing before a rebuild is an error."""
    from hagrid_b200 import HagridError
    g = Golden("cornell32")
    sc = Scene(g.tris, lib=lib)
    sc.build_all(g.top_density, g.snd_density); sc.setup_traversal()
    sc.set_tris(g.tris[:8])
    with pytest.raises(HagridError):
        sc.trace(g.rays, HIT_PRIM_ID)
    sc.set_tris(g.tris)
    sc.build_all(g.top_density, g.snd_density); sc.setup_traversal()
    assert np.array_equal(sc.trace(g.rays, HIT_PRIM_ID)["id"], g.hits["hits_cell_ids"]["id"])
    sc.close()


def test_grid_upload_rejects_bad_headers_and_keeps_the_old_grid(lib):
    from hagrid_b200 import HagridError
    g = Golden("cornell32")
    sc = Scene(g.tris, lib=lib)
    sc.build_all(g.top_density, g.snd_density); sc.setup_traversal()
    info, e, c, r = g.stage["expand"]
    for key, bad in (("num_cells", -1), ("num_entries", 0), ("num_refs", -5), ("shift", 99), ("dims", [0, 1, 1])):
        d = dict(info); d[key] = bad
        gi = info_from_dict(d)
        with pytest.raises(HagridError):
            lib.check(lib.dll.hgb_grid_upload(sc._h, gi, e.ctypes.data, c.ctypes.data, r.ctypes.data), "grid_upload")
    assert np.array_equal(sc.trace(g.rays, HIT_PRIM_ID)["id"], g.hits["hits_cell_ids"]["id"])     # untouched
    sc.close()


def test_destroying_a_keep_alive_scene_returns_its_memory(lib):
    """keep-alive slots used to stay allocated after hgb_scene_destroy."""
    import torch
    tris = scenes.hairball(200000, seed=3)
    free0 = torch.cuda.mem_get_info(0)[0]
    for _ in range(3):
        sc = Scene(tris, keep_alive=True, lib=lib)
        sc.build_all(0.12, 2.4, warmup=1, iters=2)
        assert sc.peak_bytes() > 50 << 20
        sc.close()
    free1 = torch.cuda.mem_get_info(0)[0]
    assert free0 - free1 < 32 << 20, (free0, free1)


def test_buffer_pool_reuse(lib):
    sc = Scene(scenes.cornell32(), keep_alive=True, lib=lib)
    a = sc.device_alloc(1 << 20)
    sc.device_free(a)
    b = sc.device_alloc(1 << 20)
    assert a == b                                # keep mode hands the retained slot back
    sc.device_free(b)
    assert sc.peak_bytes() >= 1 << 20
    sc.close()
