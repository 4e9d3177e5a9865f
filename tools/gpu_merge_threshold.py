"""merge_grid in one cooperative launch against one launch per kernel and pass (run under gpurun): whole builds of scenes
of growing size, both paths forced in turn; prints the cell count merge_grid starts from, so that the line between the
two ("merge_one_launch_max_cells") can be read off."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
import numpy as np
from hagrid_b200 import Library, Scene, scenes
lib = Library()
cases = [("sponza262k", scenes.sponza262k(), 0.15, 3.0)]
cases += [(f"hairball{n // 1000}k", scenes.hairball(n, seed=3), 0.12, 2.4) for n in (20000, 60000, 150000, 300000, 600000)]
if "--big" in sys.argv:
    cases += [("hairball2m", scenes.hairball(), 0.12, 2.4), ("sanmiguel7p8m", scenes.sanmiguel7p8m(), 0.15, 3.0)]
for name, tris, td, sd in cases:
    ref = None
    pre = Scene(tris, lib=lib); pre.build_grid(td, sd); start_cells = pre.download()[0].num_cells; pre.close()
    for mode, val in (("per-launch", 0), ("one-launch", 1 << 30)):
        lib.set_option("merge_one_launch_max_cells", val)
        sc = Scene(tris, keep_alive=True, lib=lib)
        ms = sc.build_all(td, sd, 0.995, 3, False, warmup=3, iters=8)
        gi, e, c, r = sc.download()
        blob = (e.tobytes(), c.tobytes(), r.tobytes()); ref = ref or blob
        print(f"{name:14s} {mode:10s} build mean {ms.mean():.3f} min {ms.min():.3f} ms  cells before merge {start_cells} after {gi.num_cells}  identical {blob == ref}", flush=True)
        sc.close()
lib.set_option("merge_one_launch_max_cells", -1)
