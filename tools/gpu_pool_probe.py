"""Allocator probe (run under gpurun): repeated keep-alive builds of a large scene; per-iteration time, pool size."""
import json, sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import Library, Scene, scenes
which = sys.argv[1] if len(sys.argv) > 1 else "sanmiguel"
keep = (sys.argv[2] if len(sys.argv) > 2 else "keep") == "keep"
tris = scenes.sanmiguel7p8m() if which == "sanmiguel" else scenes.hairball()
lib = Library()
sc = Scene(tris, keep_alive=keep, lib=lib)
rows = []
for i in range(12):
    ms = sc.build_all(0.15, 3.0, 0.995, 3, False, warmup=0, iters=1)
    free, total = torch.cuda.mem_get_info()
    rows.append((round(float(ms[0]), 1), round((total - free) / 2**30, 2), round(sc.peak_bytes() / 2**30, 2)))
print(json.dumps({"scene": which, "keep": keep, "ms, device GiB in use, MemManager peak GiB": rows}))
