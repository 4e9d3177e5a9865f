"""Where a rank's time goes in the sharded C5 frame (one GPU stands in for rank 0 of 1, 2, 4, 8): the two traversals, the
bounce kernel and the counters timed one by one (CUDA events, L2 flushed). (gpurun)"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes, sharding
lib = Library()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tris = scenes.sanmiguel7p8m()
sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(0.15, 3.0); sc.setup_traversal()
primary = scenes.default_view(tris)
lo, hi = scenes.scene_bbox(tris); diag = float(np.linalg.norm(hi - lo))
def timed(fn, iters=15):
    for _ in range(3):
        flush.zero_(); fn()
    torch.cuda.synchronize()
    a = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]; b = [torch.cuda.Event(enable_timing=True) for _ in range(iters)]
    for i in range(iters):
        flush.zero_(); a[i].record(); fn(); b[i].record()
    torch.cuda.synchronize()
    return float(np.mean([x.elapsed_time(y) for x, y in zip(a, b)]))
for world in (1, 2, 4, 8):
    for rank in ((0,) if world == 1 else (0, world - 1)):
        idx = sharding.interleaved_bands(primary.shape[0], rank, world, sharding.raster_granule(1920))
        mine = np.ascontiguousarray(primary[idx]); n = mine.shape[0]
        d_rays = torch.from_numpy(mine.view(np.float32).reshape(n, 8)).cuda()
        h1 = torch.empty((n, 4), dtype=torch.float32, device="cuda"); h2 = torch.empty_like(h1); bounce = torch.empty_like(d_rays)
        keys = torch.from_numpy(idx.astype(np.int32)).cuda(); counters = torch.zeros(2, dtype=torch.int64, device="cuda")
        t1 = timed(lambda: sc.traverse(d_rays, h1, n, HIT_PRIM_ID))
        tb = timed(lambda: sc.bounce_rays_keyed(d_rays, h1, n, 1e-3 * diag, diag, 7, keys, bounce))
        t2 = timed(lambda: sc.traverse(bounce, h2, n, HIT_PRIM_ID))
        tc = timed(lambda: sc.count_hits(h1, n, counters))
        tf = timed(lambda: (counters.zero_(), sc.trace_two_waves(d_rays, n, keys, 1e-3 * diag, diag, 7, h1, bounce, h2, counters)))
        steps = None
        print(f"world {world} rank {rank}: rays {n}  primary {t1:.4f}  bounce kernel {tb:.4f}  second wave {t2:.4f}  one count {tc:.4f}  whole frame call {tf:.4f} ms", flush=True)
