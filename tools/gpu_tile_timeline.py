"""Per-warp time line of one tile launch (diagnosis build with HGB_TILE_TRACE): when do warps start, when do they run dry,
how long is the tail. (gpurun)"""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes
lib = Library(sorted((ROOT / "hagrid_b200/_build/variants").glob("*TRACE*.so"))[0])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "sponza"
tris = scenes.sponza262k() if which == "sponza" else scenes.sanmiguel7p8m()
rays = scenes.default_view(tris); n = rays.shape[0]
sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(0.15, 3.0); sc.setup_traversal()
d_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda(); d_hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
for k, v in (a.split("=") for a in sys.argv[2:]):
    lib.set_option(k, int(v))
for rep in range(6):          # from the third launch on the tiles are handed out by the ticket list (unless tile_order=0)
    flush.zero_(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID); b.record(); torch.cuda.synchronize()
    buf = np.zeros(3 * 8192, dtype=np.int64)
    assert lib.dll.hgb_debug_tile_trace(C.c_void_p(buf.ctypes.data)) == 0
    t = buf.reshape(-1, 3)
    t = t[t[:, 1] > 0]                              # the warps of this launch (148 SMs x 12 blocks x 4)
    t0 = t[:, 0].min()
    start, end, cnt = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, t[:, 2]
    q = lambda x, p: float(np.percentile(x, p))
    print(f"rep {rep}: event {a.elapsed_time(b) * 1e3:.1f} us | warp starts p50 {q(start, 50):.1f} p99 {q(start, 99):.1f} max {start.max():.1f} | "
          f"warp ends p1 {q(end, 1):.1f} p10 {q(end, 10):.1f} p50 {q(end, 50):.1f} p90 {q(end, 90):.1f} p99 {q(end, 99):.1f} max {end.max():.1f} | "
          f"tiles/warp min {cnt.min()} mean {cnt.mean():.1f} max {cnt.max()}", flush=True)
    # active warps over time
    grid_t = np.linspace(0, end.max(), 21)
    act = [(int(((start <= x) & (end > x)).sum())) for x in grid_t]
    print("   active warps at 5% steps:", act, flush=True)
