"""Builds one scene a few times (run under ncu for a per-kernel launch list, or plain for timings).
usage: gpu_build_profile.py [hair2m|sponza|sanmiguel] [iters] [compress]"""
import json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import Library, Scene, scenes
which = sys.argv[1] if len(sys.argv) > 1 else "hair2m"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
compress = len(sys.argv) > 3 and sys.argv[3] == "compress"
tris, td, sd = {"hair2m": (scenes.hairball, 0.12, 2.4), "sponza": (scenes.sponza262k, 0.15, 3.0),
                "sanmiguel": (scenes.sanmiguel7p8m, 0.15, 3.0)}[which]
tris = tris()
lib = Library()
sc = Scene(tris, keep_alive=True, lib=lib)
stage_ms = {}
import time
for it in range(iters):
    for name, fn in (("build", lambda: sc.build_grid(td, sd)), ("merge", lambda: sc.merge_grid(0.995)), ("flatten", sc.flatten_grid),
                     ("expand", lambda: sc.expand_grid(3))) + ((("compress", sc.compress_grid),) if compress else ()):
        lib.synchronize(); t0 = time.perf_counter(); fn(); lib.synchronize()
        stage_ms.setdefault(name, []).append(round((time.perf_counter() - t0) * 1e3, 3))
print(json.dumps({"scene": which, "stage_wall_ms": stage_ms, "grid": {k: sc.info().as_dict()[k] for k in ("dims", "shift", "num_cells", "num_entries", "num_refs")}}))
