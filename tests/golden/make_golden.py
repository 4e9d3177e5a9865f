"""Generates the golden fixtures in this directory from the REFERENCE itself.

Run on a GPU box (it needs oracle/_ref/libhagrid_ref.so = cg-saarland/hagrid rebuilt
for sm_100a by oracle/build_ref.sh, driven through the include/hagrid_b200.h ABI):

    gpurun -- 'python tests/golden/make_golden.py && cp tests/golden/*.npz gpurun_out/'

For each small scene the file holds the triangles, the build parameters, the grid
after every construction stage (build, merge, flatten, expand, compress) and, for
the expanded (Cell) and the compressed (SmallCell) grid, the reference's hits for
a fixed ray buffer in both Hit.id modes (steps = verbatim src/traverse.cu:93,
ids = the same kernel with that line removed).
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, HIT_STEPS, Library, Scene, scenes  # noqa: E402

HERE = Path(__file__).resolve().parent
ref = Library(ROOT / "oracle/_ref/libhagrid_ref.so")
assert ref.impl == "reference"


def edge_rays(tris, n=256, seed=3):
    """Axis-parallel rays, rays starting inside/outside, grazing the box, zero-length and reversed ranges."""
    rng = np.random.default_rng(seed)
    lo, hi = scenes.scene_bbox(tris)
    rays = scenes.random_rays(tris, n, seed=seed, tmax=float(np.linalg.norm(hi - lo)) * 2)
    d = rays["dir"]
    d[0:32, 1:] = 0.0; d[0:32, 0] = np.where(rng.random(32) < 0.5, 1.0, -1.0)          # +-x
    d[32:64, 0] = 0.0; d[32:64, 2] = 0.0; d[32:64, 1] = np.where(rng.random(32) < 0.5, 1.0, -1.0)
    d[64:96, :2] = 0.0; d[64:96, 2] = np.where(rng.random(32) < 0.5, 1.0, -1.0)
    rays["org"][96:128] = lo - (hi - lo) * rng.random((32, 3), dtype=np.float32)         # from outside
    rays["org"][128:144] = hi + (hi - lo)                                                # pointing away / missing
    rays["tmax"][144:160] = 1e-3                                                         # very short
    rays["tmin"][160:176] = 0.25 * float(np.linalg.norm(hi - lo))                        # late start
    rays["tmax"][176:184] = -1.0                                                         # empty interval
    rays["dir"][184:192] = 0.0                                                           # degenerate direction
    return rays


def stage_dump(sc):
    gi, e, c, r = sc.download()
    return gi.as_dict(), e, c, r


def make(name, tris, td, sd, rays):
    out = {"tris": tris, "rays": rays, "params": np.array([td, sd, 0.995, 3], dtype=np.float64)}
    sc = Scene(tris, lib=ref)
    for stage, fn in (("build", lambda: sc.build_grid(td, sd)), ("merge", lambda: sc.merge_grid(0.995)),
                      ("flatten", sc.flatten_grid), ("expand", lambda: sc.expand_grid(3)), ("compress", sc.compress_grid)):
        if stage == "compress":
            sc.setup_traversal()
            out["hits_cell_steps"] = sc.trace(rays, HIT_STEPS)
            out["hits_cell_ids"] = sc.trace(rays, HIT_PRIM_ID)
        fn()
        info, e, c, r = stage_dump(sc)
        out[f"{stage}_info"] = np.frombuffer(json.dumps(info).encode(), dtype=np.uint8)
        out[f"{stage}_entries"], out[f"{stage}_cells"], out[f"{stage}_refs"] = e, c, r
    sc.setup_traversal()
    out["hits_small_steps"] = sc.trace(rays, HIT_STEPS)
    out["hits_small_ids"] = sc.trace(rays, HIT_PRIM_ID)
    np.savez_compressed(HERE / f"{name}.npz", **out)
    ids = out["hits_cell_ids"]["id"]
    print(name, json.loads(bytes(out["compress_info"]).decode()), "hit fraction", float((ids >= 0).mean()),
          "ids equal cell/small", bool((ids == out["hits_small_ids"]["id"]).all()))
    sc.close()


corn = scenes.cornell32()
make("cornell32", corn, 0.12, 2.4,
     np.concatenate([scenes.cornell_view(48, 48), scenes.random_rays(corn, 1024, tmax=2000.0), edge_rays(corn)]))
soup = scenes.small_mixed(800, seed=5)
make("soup800", soup, 0.12, 2.4,
     np.concatenate([scenes.random_rays(soup, 3072, tmax=20.0), edge_rays(soup)]))
strands = scenes.hairball(1500, seed=9)
make("strands1500", strands, 0.15, 3.0,
     np.concatenate([scenes.random_rays(strands, 2048, tmax=10.0), edge_rays(strands)]))
