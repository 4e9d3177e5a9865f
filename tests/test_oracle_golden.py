"""CPU tier: the oracle (oracle/hagrid_oracle.c) is pinned against outputs of the reference
itself (golden fixtures made on a B200 from the reference rebuilt for sm_100a)."""
import numpy as np
import pytest

from oracle import oracle
from util import FIXTURES, STAGES, Golden, grid_diff, t_close


@pytest.fixture(scope="module", params=FIXTURES)
def golden(request):
    return Golden(request.param)


def _run(grid, stage, g):
    if stage == "merge": grid.merge(g.alpha)
    elif stage == "flatten": grid.flatten()
    elif stage == "expand": grid.expand(g.expansion)
    elif stage == "compress": assert grid.compress()


def test_build_matches_reference(golden):
    grid = oracle.Grid.build(golden.tris, golden.top_density, golden.snd_density)
    assert grid_diff(grid.info(), grid.arrays(), golden.stage["build"]) == []


@pytest.mark.parametrize("stage", STAGES[1:])
def test_stage_on_reference_input_is_bit_exact(golden, stage):
    """merge (SAH in the device's fused shapes), flatten, expand (incl. the ping-pong quirk) and
    compress are integer/exact-float work: byte-identical to the reference given its input."""
    prev = STAGES[STAGES.index(stage) - 1]
    grid = oracle.Grid.from_arrays(*golden.stage[prev])
    _run(grid, stage, golden)
    assert grid_diff(grid.info(), grid.arrays(), golden.stage[stage]) == []


def test_full_pipeline_chained(golden):
    grid = oracle.Grid.build(golden.tris, golden.top_density, golden.snd_density)
    for stage in STAGES[1:]:
        _run(grid, stage, golden)
        assert grid_diff(grid.info(), grid.arrays(), golden.stage[stage]) == [], stage


@pytest.mark.parametrize("cells", ["cell", "small"])
def test_traversal_matches_reference(golden, cells):
    """Step counts are exact; prim ids are exact on these fixtures; t within 1e-5 (the CPU has
    no MUFU.RCP, so hit distances may differ from the device in the last bits)."""
    grid = oracle.Grid.from_arrays(*golden.stage["expand" if cells == "cell" else "compress"])
    steps = grid.traverse(golden.tris, golden.rays, mode=0, threads=2)
    ids = grid.traverse(golden.tris, golden.rays, mode=1, threads=1)
    want_steps, want_ids = golden.hits[f"hits_{cells}_steps"], golden.hits[f"hits_{cells}_ids"]
    assert np.array_equal(steps["id"], want_steps["id"])
    assert np.array_equal(ids["id"], want_ids["id"])
    assert t_close(ids["t"], want_ids["t"]).all()
    assert t_close(steps["t"], want_steps["t"]).all()
    assert (ids["u"] == 0).all() and (ids["v"] == 0).all()      # COMPUTE_UVS is never defined (src/prims.h:285-288)


def test_cell_and_small_cell_agree(golden):
    """Compression must not change what is hit; step counts grow by one per non-empty cell visited."""
    a, b = golden.hits["hits_cell_ids"], golden.hits["hits_small_ids"]
    assert np.array_equal(a["id"], b["id"]) and np.array_equal(a["t"].view(np.uint32), b["t"].view(np.uint32))
    assert (golden.hits["hits_small_steps"]["id"] >= golden.hits["hits_cell_steps"]["id"]).all()


def test_grid_invariants(golden):
    """Structural invariants of SURVEY.md A.3-A.6 on the reference's own output."""
    info, entries, cells, refs = golden.stage["expand"]
    vd = np.array(info["dims"]) << info["shift"]
    assert (cells["min"] >= 0).all() and (cells["max"] <= vd).all() and (cells["min"] < cells["max"]).all()
    assert (cells["begin"] <= cells["end"]).all() and cells["end"].max() <= info["num_refs"]
    assert refs.min() >= 0 and refs.max() < golden.tris.shape[0]
    leaf = (entries & 3) == 0
    assert (entries[leaf] >> 2).max() < info["num_cells"]
    assert ((entries[~leaf] >> 2) < info["num_entries"]).all()
    sinfo, _, scells, srefs = golden.stage["compress"]
    n = cells["end"] - cells["begin"]
    assert sinfo["num_refs"] == int((n[n > 0] + 1).sum())
    assert np.array_equal(scells["begin"] < 0, n == 0)
    assert (srefs == -1).sum() == int((n > 0).sum())


def test_flat_scene_box_is_given_a_thickness():
    """All triangles in one axis-aligned plane: the reference divides by a zero volume (src/grid.h:96-101).
    Oracle and product share one rule for this case (0.1 % of the widest extent); the hits must be right."""
    from hagrid_b200 import RAY_DTYPE, scenes
    tris = scenes.make_tris([[0, 0, 0], [0, 0, 0]], [[1, 0, 0], [0, 0, 0]], [[0, 1, 0], [0, 0, 0]])
    for t in (tris[:1], tris):
        grid = oracle.Grid.build(t, 0.12, 2.4)
        grid.merge(); grid.flatten(); grid.expand()
        info = grid.info()
        assert info["bbox_max"][2] > info["bbox_min"][2] and info["shift"] < 8
        rays = np.zeros(2, dtype=RAY_DTYPE)
        rays["org"] = [(0.25, 0.25, 1.0), (2.0, 2.0, 1.0)]; rays["dir"] = (0, 0, -1); rays["tmax"] = 10.0
        res = grid.traverse(t, rays, 1)
        assert res["id"][0] == 0 and abs(res["t"][0] - 1.0) < 1e-6 and res["id"][1] == -1
