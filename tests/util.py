"""Shared helpers for the test-suite: golden fixtures (tests/golden/*.npz, produced from the
reference by tests/golden/make_golden.py) and exact grid comparison."""
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"
STAGES = ("build", "merge", "flatten", "expand", "compress")
FIXTURES = ("cornell32", "soup800", "strands1500")


class Golden:
    def __init__(self, name):
        z = np.load(GOLDEN / f"{name}.npz")
        self.name = name
        self.tris = z["tris"]
        self.rays = z["rays"]
        self.top_density, self.snd_density, self.alpha, self.expansion = (float(z["params"][0]), float(z["params"][1]),
                                                                          float(z["params"][2]), int(z["params"][3]))
        self.stage = {s: (json.loads(bytes(z[f"{s}_info"]).decode()), z[f"{s}_entries"], z[f"{s}_cells"], z[f"{s}_refs"])
                      for s in STAGES}
        self.hits = {k: z[k] for k in ("hits_cell_steps", "hits_cell_ids", "hits_small_steps", "hits_small_ids")}


def grid_diff(got_info: dict, got_arrays, want) -> list:
    """Empty list when the grid (host fields + entries, cells, refs) is byte-identical to `want`."""
    info, e, c, r = want
    ge, gc, gr = got_arrays
    out = []
    for k in ("dims", "shift", "num_cells", "num_entries", "num_refs", "compressed", "offsets"):
        if got_info[k] != info[k]:
            out.append(f"{k}: {got_info[k]} != {info[k]}")
    for k in ("bbox_min", "bbox_max"):
        if np.asarray(got_info[k], np.float32).tobytes() != np.asarray(info[k], np.float32).tobytes():
            out.append(f"{k}: {got_info[k]} != {info[k]}")
    for name, a, b in (("entries", ge, e), ("cells", gc, c), ("refs", gr, r)):
        if a.shape != b.shape or a.dtype != b.dtype:
            out.append(f"{name}: shape/dtype {a.shape} {a.dtype} != {b.shape} {b.dtype}")
        elif a.tobytes() != b.tobytes():
            rows = (a.view(np.uint8).reshape(a.shape[0], -1) != b.view(np.uint8).reshape(b.shape[0], -1)).any(axis=1)
            out.append(f"{name}: {int(rows.sum())} of {a.shape[0]} rows differ (first {int(np.argmax(rows))})")
    return out


def t_close(t, t_ref, rel=1e-5):
    """|t - t_ref| <= rel * max(1, |t_ref|), the tolerance BASELINE.json's north_star states for hit distances."""
    t = np.asarray(t, np.float64); t_ref = np.asarray(t_ref, np.float64)
    same_inf = np.isinf(t) & np.isinf(t_ref) & (np.sign(t) == np.sign(t_ref))
    with np.errstate(invalid="ignore"):
        return same_inf | (np.abs(t - t_ref) <= rel * np.maximum(1.0, np.abs(t_ref)))
