"""C5 two-wave frame on one GPU: hgb_trace_two_waves with the frame cut into 1, 2, 3, 4, 8 chunks (gpurun)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import Library, Scene, scenes
lib = Library()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tris = scenes.sanmiguel7p8m() if (len(sys.argv) < 2 or sys.argv[1] == "c5") else scenes.sponza262k()
sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(0.15, 3.0); sc.setup_traversal()
rays = scenes.default_view(tris); n = rays.shape[0]
lo, hi = scenes.scene_bbox(tris); diag = float(np.linalg.norm(hi - lo))
d_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).cuda()
h1 = torch.empty((n, 4), dtype=torch.float32, device="cuda"); h2 = torch.empty_like(h1); bounce = torch.empty_like(d_rays)
counters = torch.zeros(2, dtype=torch.int64, device="cuda")
want = None
for chunks in (1, 2, 3, 4, 8):
    lib.set_option("two_wave_chunks", chunks)
    def step():
        counters.zero_(); sc.trace_two_waves(d_rays, n, None, 1e-3 * diag, diag, 7, h1, bounce, h2, counters)
    for _ in range(4):
        flush.zero_(); step()
    torch.cuda.synchronize()
    a = [torch.cuda.Event(enable_timing=True) for _ in range(20)]; b = [torch.cuda.Event(enable_timing=True) for _ in range(20)]
    for i in range(20):
        flush.zero_(); a[i].record(); step(); b[i].record()
    torch.cuda.synchronize()
    ms = np.array([x.elapsed_time(y) for x, y in zip(a, b)])
    got = (h1.cpu().numpy().tobytes(), h2.cpu().numpy().tobytes())
    want = want or got
    print(f"chunks {chunks}: {ms.mean():.4f} ms  min {ms.min():.4f}  same {got == want}", flush=True)
