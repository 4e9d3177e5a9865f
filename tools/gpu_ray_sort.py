"""Ray binning experiment (option "ray_sort"): incoherent buffers sorted by (direction octant, top-level entry cell) with this
library's radix sort inside the traversal call, then traced by the voting kernel through the permutation. C3 (4 M random rays,
compressed grid) and the second wave of C5, against the reference, hits compared. (gpurun)"""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import importlib
sys.argv = [sys.argv[0], "none"]
g = importlib.import_module("gpu_r02_traverse")
from hagrid_b200 import scenes, HIT_PRIM_ID
settings = {"unsorted": {"ray_sort": 0}, "binned": {"ray_sort": 1}}
tris = scenes.sponza262k()
sr, sm = g.scene_pair(tris, compress=True)
g.compare_buffer("c3_random_4M_compressed", sr, sm, scenes.random_rays(tris, 1 << 22), settings, 20)
sr.close(); sm.close()
tris = scenes.sanmiguel7p8m()
sr, sm = g.scene_pair(tris)
primary = scenes.default_view(tris)
first = sm.trace(primary, HIT_PRIM_ID)
bounce = scenes.bounce_rays(tris, primary, first["id"], first["t"])
g.compare_buffer("c5_bounce", sr, sm, bounce, settings, 10)
