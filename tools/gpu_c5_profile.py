"""C5 (7.8 M triangles) traversal under ncu: which limit binds the kernels once grid + triangles (about 1 GB) no longer
fit in L2. usage (under gpurun): ncu --metrics ... -k regex:traverse python tools/gpu_c5_profile.py"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Scene, scenes

tris = scenes.sanmiguel7p8m()
primary = scenes.default_view(tris)
sc = Scene(tris, keep_alive=True)
sc.build_all(0.15, 3.0)
sc.setup_traversal()
lo, hi = scenes.scene_bbox(tris)
diag = float(np.linalg.norm(hi - lo))
n = primary.shape[0]
d_rays, d_hits, d_second = sc.device_alloc(n * 32), sc.device_alloc(n * 16), sc.device_alloc(n * 32)
sc.to_device(d_rays, primary)
for _ in range(3):
    sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID)
sc.bounce_rays_device(d_rays, d_hits, n, 1e-3 * diag, diag, 7, d_second)
for _ in range(3):
    sc.traverse(d_second, d_hits, n, HIT_PRIM_ID)
sc.lib.synchronize()
gi = sc.info()
print({"cells": gi.num_cells, "refs": gi.num_refs, "entries": gi.num_entries})
