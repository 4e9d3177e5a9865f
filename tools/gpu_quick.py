"""Scratch experiment runner (gpurun): C2 primary + long view, variants given as 'name:key=val,key=val;...'"""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import importlib
spec = sys.argv[1]
sys.argv = [sys.argv[0], "none"]
g = importlib.import_module("gpu_r02_traverse")
from hagrid_b200 import scenes
settings = {}
for item in spec.split(";"):
    name, _, kv = item.partition(":")
    settings[name] = {k: int(v) for k, v in (p.split("=") for p in kv.split(",") if p)}
which = sys.argv[2:] if False else None
tris = scenes.sponza262k()
sr, sm = g.scene_pair(tris)
g.compare_buffer("c2_primary", sr, sm, scenes.default_view(tris), settings, 50)
g.compare_buffer("c2_long", sr, sm, scenes.default_view(tris, along_long_axis=True), settings, 30)
