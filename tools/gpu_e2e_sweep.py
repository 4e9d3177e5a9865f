"""Host-buffer frame (hgb_traverse_grid_host) schedule sweep on C2 primary rays (gpurun). Settings: 'name:key=val,...;...'"""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import importlib
spec = sys.argv[1]
sys.argv = [sys.argv[0], "none"]
g = importlib.import_module("gpu_r02_traverse")
from hagrid_b200 import scenes, HIT_PRIM_ID
tris = scenes.sponza262k()
sr, sm = g.scene_pair(tris)
primary = scenes.default_view(tris)
n = primary.shape[0]
want = sm.trace(primary, HIT_PRIM_ID)
h_rays = torch.from_numpy(primary.view(np.float32).reshape(n, 8)).pin_memory()
h_hits = torch.empty((n, 4), dtype=torch.float32).pin_memory()
r = g.timed(lambda: g.ref.check(g.ref.dll.hgb_traverse_grid_host(sr._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "f"), 20, 3)
print("reference", r)
for item in spec.split(";"):
    name, _, kv = item.partition(":")
    for k, v in (p.split("=") for p in kv.split(",") if p):
        g.mine.set_option(k, int(v))
    h_hits.zero_()
    t = g.timed(lambda: g.mine.check(g.mine.dll.hgb_traverse_grid_host(sm._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "f"), 30, 3)
    ok = bool(np.array_equal(h_hits.numpy().view(np.int32)[:, 0], want["id"]) and np.array_equal(h_hits.numpy()[:, 1], want["t"]))
    print(f"{name:34s} mean {t['ms_mean']:.4f} median {t['ms_median']:.4f} min {t['ms_min']:.4f} x{r['ms_mean'] / t['ms_mean']:.3f} {ok}", flush=True)
