"""Which kernel for incoherent buffers of which size: one thread per ray (variant 0) or the persistent voting warps
(variant 1)? Second wave of the C5 frame and random rays in the C2 scene, subsampled. (gpurun)"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import importlib
sys.argv = [sys.argv[0], "none"]
g = importlib.import_module("gpu_r02_traverse")
from hagrid_b200 import scenes, HIT_PRIM_ID
settings = {"per_thread": {"traverse_variant": 0}, "voting": {"traverse_variant": 1}}
tris = scenes.sanmiguel7p8m()
sr, sm = g.scene_pair(tris)
primary = scenes.default_view(tris)
first = sm.trace(primary, HIT_PRIM_ID)
bounce = scenes.bounce_rays(tris, primary, first["id"], first["t"])
for k in (16, 8, 4, 2, 1):
    g.compare_buffer(f"c5_bounce_1of{k}", sr, sm, np.ascontiguousarray(bounce[::k]), settings, 10)
sr.close(); sm.close()
tris = scenes.sponza262k()
sr, sm = g.scene_pair(tris, compress=True)
rays = scenes.random_rays(tris, 1 << 22)
for k in (32, 16, 8, 4, 1):
    g.compare_buffer(f"c3_random_1of{k}", sr, sm, np.ascontiguousarray(rays[::k]), settings, 10)
g.mine.set_option("traverse_variant", 3)
