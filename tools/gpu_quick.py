"""Scratch experiment runner (gpurun): C2 primary + long view (+ C5 primary with 'c5'), variants given as 'name:key=val,key=val;...'"""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import importlib
spec = sys.argv[1]
scenes_wanted = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c2"]
sys.argv = [sys.argv[0], "none"]
g = importlib.import_module("gpu_r02_traverse")
from hagrid_b200 import scenes
settings = {}
for item in spec.split(";"):
    name, _, kv = item.partition(":")
    settings[name] = {k: int(v) for k, v in (p.split("=") for p in kv.split(",") if p)}
if "c2" in scenes_wanted:
    tris = scenes.sponza262k()
    sr, sm = g.scene_pair(tris)
    g.compare_buffer("c2_primary", sr, sm, scenes.default_view(tris), settings, 50)
    g.compare_buffer("c2_long", sr, sm, scenes.default_view(tris, along_long_axis=True), settings, 30)
    sr.close(); sm.close()
if "c5" in scenes_wanted:
    tris = scenes.sanmiguel7p8m()
    sr, sm = g.scene_pair(tris)
    g.compare_buffer("c5_primary", sr, sm, scenes.default_view(tris), settings, 30)
