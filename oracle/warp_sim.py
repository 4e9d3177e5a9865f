"""TEST/ANALYSIS INFRASTRUCTURE (uses the CPU oracle; nothing in the product imports it).
Offline SIMT model (CPU, oracle only): warp-level iteration counts of scheduling strategies for the
traversal loop, from the per-ray sequences of (cell visit, #references) the oracle records.
  if-while   : every iteration = one cell step for all live lanes, then a triangle loop of max(count) trips
  while-while: lanes step until they own a non-empty cell (or die), then one triangle loop
Costs are weighted with the SASS instruction counts of the two loop bodies."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle
from hagrid_b200 import scenes

CELL, TRI = 100, 45      # instructions per cell step / per triangle test (cuobjdump of ray_traverse.o)

def simulate(seqs):
    """seqs: (32, K) counts, -1 padded. Returns dict strategy -> (cell_iters, tri_iters)."""
    lens = (seqs >= 0).sum(1)
    # if-while
    K = lens.max()
    cell_if = K
    tri_if = sum(int(seqs[:, k][seqs[:, k] >= 0].max(initial=0)) for k in range(K))
    # while-while
    pos = np.zeros(32, int); cell_ww = tri_ww = 0
    while True:
        live = pos < lens
        if not live.any(): break
        # search phase: each live lane advances to its next non-empty cell (inclusive) or to the end
        adv = np.zeros(32, int); cnt = np.zeros(32, int)
        for l in np.nonzero(live)[0]:
            p = pos[l]; a = 0
            while p < lens[l]:
                a += 1
                if seqs[l, p] > 0: cnt[l] = seqs[l, p]; p += 1; break
                p += 1
            pos[l] = p; adv[l] = a
        cell_ww += adv.max(); tri_ww += cnt.max()
    return {"if-while": (cell_if, tri_if), "while-while": (cell_ww, tri_ww)}

def main():
    tris = scenes.sponza262k()
    g = oracle.Grid.build(tris, 0.15, 3.0); g.merge(0.995); g.flatten(); g.expand(3)
    rays = scenes.default_view(tris).reshape(1080, 1920)
    rng = np.random.default_rng(1)
    tot = {}; thread_cells = thread_tris = 0; nw = 0
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 300):
        y = int(rng.integers(0, 1080 // 4)) * 4; x = int(rng.integers(0, 1920 // 8)) * 8
        tile = np.ascontiguousarray(rays[y:y + 4, x:x + 8].reshape(-1))
        seqs = g.record(tris, tile, 128).astype(int)
        thread_cells += (seqs >= 0).sum(); thread_tris += seqs[seqs > 0].sum(); nw += 1
        for k, (c, t) in simulate(seqs).items():
            a = tot.setdefault(k, [0, 0]); a[0] += c; a[1] += t
    print(f"{nw} warps: per-thread cells {thread_cells / nw / 32:.2f} tris {thread_tris / nw / 32:.2f}")
    for k, (c, t) in tot.items():
        print(f"{k:12s} cell iters/warp {c / nw:6.2f} tri iters/warp {t / nw:6.2f}  cost {(c * CELL + t * TRI) / nw:8.0f}")

if __name__ == "__main__":
    main()
