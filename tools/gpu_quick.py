"""Scratch experiment runner (gpurun): buffers of C2 / C5 under several option settings 'name:key=val,key=val;...'
usage: gpu_quick.py SPEC [c2,c5,c5b,c3]"""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import importlib
spec = sys.argv[1]
wanted = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c2"]
sys.argv = [sys.argv[0], "none"]
g = importlib.import_module("gpu_r02_traverse")
from hagrid_b200 import scenes, HIT_PRIM_ID, sharding
settings = {}
for item in spec.split(";"):
    name, _, kv = item.partition(":")
    settings[name] = {k: int(v) for k, v in (p.split("=") for p in kv.split(",") if p)}
if "c2" in wanted:
    tris = scenes.sponza262k()
    sr, sm = g.scene_pair(tris)
    g.compare_buffer("c2_primary", sr, sm, scenes.default_view(tris), settings, 50)
    g.compare_buffer("c2_long", sr, sm, scenes.default_view(tris, along_long_axis=True), settings, 30)
    if "shards" in wanted:
        primary = scenes.default_view(tris)
        for world in (2, 4, 8):
            idx = sharding.interleaved_bands(primary.shape[0], 0, world, sharding.raster_granule(1920))
            g.compare_buffer(f"c2_primary_1of{world}", sr, sm, np.ascontiguousarray(primary[idx]), settings, 30)
    sr.close(); sm.close()
if "c3" in wanted:
    tris = scenes.sponza262k()
    sr, sm = g.scene_pair(tris, compress=True)
    g.compare_buffer("c3_random", sr, sm, scenes.random_rays(tris, 1 << 22), settings, 15)
    sr.close(); sm.close()
if "c5" in wanted or "c5b" in wanted:
    tris = scenes.sanmiguel7p8m()
    sr, sm = g.scene_pair(tris)
    primary = scenes.default_view(tris)
    if "c5" in wanted:
        g.compare_buffer("c5_primary", sr, sm, primary, settings, 30)
        for world in (2, 4, 8):
            idx = sharding.interleaved_bands(primary.shape[0], 0, world, sharding.raster_granule(1920))
            g.compare_buffer(f"c5_primary_1of{world}", sr, sm, np.ascontiguousarray(primary[idx]), settings, 30)
    if "c5b" in wanted:
        first = sm.trace(primary, HIT_PRIM_ID)
        bounce = scenes.bounce_rays(tris, primary, first["id"], first["t"])
        g.compare_buffer("c5_bounce", sr, sm, bounce, settings, 10)
        g.compare_buffer("c5_bounce_1of8", sr, sm, np.ascontiguousarray(bounce[:: 8]), settings, 15)
