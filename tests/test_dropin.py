"""Drop-in tier: the reference's unmodified front end (src/main.cpp + src/load_obj.cpp) compiled against
this repository's headers and linked with this repository's kernels (oracle/_ref/hagrid_dropin, built by
oracle/build_ref.sh) must behave like the reference's own executable (oracle/_ref/hagrid_ref) on the same
.obj and .rays files: same grid statistics line, same memory report, same intersection count."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from hagrid_b200 import scenes

ROOT = Path(__file__).resolve().parent.parent
REF_EXE = ROOT / "oracle" / "_ref" / "hagrid_ref"
NEW_EXE = ROOT / "oracle" / "_ref" / "hagrid_dropin"


def _need(*paths):
    for p in paths:
        if not p.exists():
            pytest.skip(f"{p} not built (oracle/build_ref.sh needs /root/reference)")


def test_front_end_links_against_this_library():
    """CPU tier: the drop-in executable exists, resolves every hagrid:: symbol from this repository's
    objects and answers the reference's usage text without a GPU."""
    _need(NEW_EXE)
    out = subprocess.run([str(NEW_EXE)], capture_output=True, text=True, timeout=60)
    assert "Usage: hagrid [options] file" in out.stdout
    syms = subprocess.run(["nm", "-C", "--defined-only", str(NEW_EXE)], capture_output=True, text=True, check=True).stdout
    for name in ("hagrid::build_grid(", "hagrid::merge_grid(", "hagrid::flatten_grid(", "hagrid::expand_grid(",
                 "hagrid::compress_grid(", "hagrid::setup_traversal(", "hagrid::traverse_grid(", "hagrid::profile(",
                 "hagrid::MemManager::alloc_slot("):
        assert name in syms, name
    assert "traverse_voting" in syms and "traverse_tiles" in syms          # our kernels, not the reference's `traverse<...>`


def _report(text):
    """The deterministic part of the front end's report (timings removed)."""
    m = re.search(r"\((\d+x\d+x\d+), (\d+) cells, (\d+) references\)", text)
    assert m, text
    keep = [l for l in text.splitlines() if l.split(":")[0] in ("Total memory", "Cells", "Entries", "References", "Triangles")]
    intr = re.search(r"(\d+) intersection\(s\)", text)
    return {"dims": m.group(1), "cells": int(m.group(2)), "refs": int(m.group(3)), "memory": keep,
            "intersections": int(intr.group(1)) if intr else None}


@pytest.mark.gpu
@pytest.mark.parametrize("compress", [False, True])
def test_front_end_reports_the_same_grid_and_hits(tmp_path, compress):
    _need(REF_EXE, NEW_EXE)
    tris = scenes.small_mixed(20000, seed=5)
    obj, rays_file = tmp_path / "scene.obj", tmp_path / "view.rays"
    scenes.write_obj(obj, tris)
    lo, hi = scenes.scene_bbox(tris)
    rays = np.concatenate([scenes.primary_rays(lo - (hi - lo) * 0.5, 0.5 * (lo + hi), (0, 1, 0), 55.0, 320, 200, 100.0),
                           scenes.random_rays(tris, 50000, seed=3)])
    scenes.write_rays(rays_file, rays)
    args = ["-td", "0.15", "-sd", "3.0", "-r", str(rays_file), "-n", "3", "-w", "1", "-tmax", "100", "-nb", "2", "-wb", "1"]
    if compress:
        args.append("--compress")
    reports = []
    for exe in (REF_EXE, NEW_EXE):
        res = subprocess.run([str(exe)] + args + [str(obj)], capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr
        reports.append(_report(res.stdout))
    assert reports[0] == reports[1]
    assert reports[1]["intersections"] == rays.shape[0]      # Hit.id = step count >= 0 for every ray (src/traverse.cu:93)


@pytest.mark.gpu
def test_front_end_keep_alive_rebuilds(tmp_path):
    """--keep-alive --build-iter N (the construction benchmark of BASELINE.json's C4) rebuilds through the
    caller's MemManager: mem.free of the previous grid's arrays must find them in the pool."""
    _need(NEW_EXE)
    tris = scenes.hairball(30000, seed=9)
    obj, rays_file = tmp_path / "hair.obj", tmp_path / "r.rays"
    scenes.write_obj(obj, tris)
    scenes.write_rays(rays_file, scenes.random_rays(tris, 4096, seed=1))
    res = subprocess.run([str(NEW_EXE), "-k", "-nb", "4", "-wb", "2", "-r", str(rays_file), str(obj)],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert _report(res.stdout)["cells"] > 0
