"""Prints the per-chunk timeline of one host-buffer frame (HGB_FRAME_TRACE=1), C2 primary rays (gpurun)."""
import os, sys
os.environ["HGB_FRAME_TRACE"] = "1"
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes
lib = Library()
tris = scenes.sponza262k()
sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(0.15, 3.0); sc.setup_traversal()
rays = scenes.default_view(tris); n = rays.shape[0]
h_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).pin_memory()
h_hits = torch.empty((n, 4), dtype=torch.float32).pin_memory()
for chunk in [int(a) for a in sys.argv[1:]] or [256]:
    lib.set_option("host_frame_chunk_rays", chunk * 1024)
    for rep in range(3):
        print(f"--- chunk {chunk}K rep {rep}", file=sys.stderr, flush=True)
        lib.check(lib.dll.hgb_traverse_grid_host(sc._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "f")
