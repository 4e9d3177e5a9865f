"""Would the incoherent kernel gain from being handed its rays most expensive first? (run under gpurun)
The second wave of the C5 frame and the C3 random rays, traced in buffer order and with their blocks of 32 / 256 / 2048 rays
sorted by the steps their rays take (descending and, for contrast, ascending): launch times only, the hits move with the rays."""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import importlib
sys.argv = [sys.argv[0], "none"]
g = importlib.import_module("gpu_r02_traverse")
from hagrid_b200 import scenes, HIT_PRIM_ID, HIT_STEPS

def probe(tag, sm, rays):
    steps = sm.trace(rays, HIT_STEPS)["id"].astype(np.int64)
    n = rays.shape[0]
    res = {"n": n, "steps_mean": float(steps.mean()), "steps_max": int(steps.max())}
    d_rays, d_hits = g.dev(rays)
    res["buffer_order"] = g.timed(lambda: sm.traverse(d_rays, d_hits, n, HIT_PRIM_ID), 10)
    for block in (32, 256, 2048):
        m = n // block * block
        cost = steps[:m].reshape(-1, block).sum(axis=1)
        for name, order in (("descending", np.argsort(-cost, kind="stable")), ("ascending", np.argsort(cost, kind="stable"))):
            idx = (order[:, None] * block + np.arange(block)[None, :]).reshape(-1)
            idx = np.concatenate([idx, np.arange(m, n)])
            d2, h2 = g.dev(np.ascontiguousarray(rays[idx]))
            res[f"block{block}_{name}"] = g.timed(lambda: sm.traverse(d2, h2, n, HIT_PRIM_ID), 10)
            del d2, h2
    print(tag, json.dumps(res), flush=True)

tris = scenes.sanmiguel7p8m()
sr, sm = g.scene_pair(tris)
primary = scenes.default_view(tris)
first = sm.trace(primary, HIT_PRIM_ID)
bounce = scenes.bounce_rays(tris, primary, first["id"], first["t"])
probe("c5_bounce", sm, bounce)
probe("c5_bounce_1of8", sm, np.ascontiguousarray(bounce[::8]))
sr.close(); sm.close()
tris = scenes.sponza262k()
sr, sm = g.scene_pair(tris, compress=True)
probe("c3_random", sm, scenes.random_rays(tris, 1 << 22))
sr.close(); sm.close()
