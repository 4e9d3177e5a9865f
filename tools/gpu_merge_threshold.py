import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent; sys.path.insert(0, str(ROOT))
import numpy as np
from hagrid_b200 import Library, Scene, scenes
lib = Library()
for name, tris, td, sd in (("c2", scenes.sponza262k(), 0.15, 3.0), ("c4", scenes.hairball(), 0.12, 2.4), ("c5", scenes.sanmiguel7p8m(), 0.15, 3.0)):
    ref = None
    for mode, val in (("per-launch", 0), ("one-launch", 1 << 30)):
        lib.set_option("merge_one_launch_max_cells", val)
        sc = Scene(tris, keep_alive=True, lib=lib)
        ms = sc.build_all(td, sd, 0.995, 3, False, warmup=3, iters=6)
        gi, e, c, r = sc.download()
        blob = (e.tobytes(), c.tobytes(), r.tobytes()); ref = ref or blob
        print(f"{name} {mode:10s} build mean {ms.mean():.3f} min {ms.min():.3f} cells {gi.num_cells} identical {blob == ref}", flush=True)
        sc.close()
