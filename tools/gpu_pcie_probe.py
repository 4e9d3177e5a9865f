"""PCIe probe (run under gpurun): H2D, D2H and concurrent H2D+D2H bandwidth from pinned memory, and the
timeline of hgb_traverse_grid_host chunks."""
import json, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
out = {}
n_in, n_out = 66355200, 33177600
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
def chunked(k=16):
    ci, co = n_in // k, n_out // k
    for i in range(k):
        with torch.cuda.stream(s1): d_in[i*ci:(i+1)*ci].copy_(h_in[i*ci:(i+1)*ci], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*co:(i+1)*co].copy_(d_out[i*co:(i+1)*co], non_blocking=True)
t = timed(h2d); out["h2d_ms"] = round(t*1e3, 3); out["h2d_gbs"] = round(n_in/t/1e9, 1)
t = timed(d2h); out["d2h_ms"] = round(t*1e3, 3); out["d2h_gbs"] = round(n_out/t/1e9, 1)
t = timed(both); out["both_ms"] = round(t*1e3, 3)
t = timed(chunked); out["both_chunked16_ms"] = round(t*1e3, 3)
print(json.dumps(out))
