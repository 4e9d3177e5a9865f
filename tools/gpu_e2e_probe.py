"""Host-buffer frame timings (run under gpurun): staged with different chunk counts vs zero-copy."""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_PRIM_ID, Library, Scene, scenes
lib = Library()
tris = scenes.sponza262k()
sc = Scene(tris, keep_alive=True, lib=lib); sc.build_all(0.15, 3.0); sc.setup_traversal()
kinds = {"primary": scenes.default_view(tris), "random4M": scenes.random_rays(tris, 1 << 22)}
out = {}
for kind, rays in kinds.items():
    n = rays.shape[0]
    want = sc.trace(rays, HIT_PRIM_ID)
    h_rays = torch.from_numpy(rays.view(np.float32).reshape(n, 8)).pin_memory()
    h_hits = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    def run(label, reps=20):
        for _ in range(3): lib.check(lib.dll.hgb_traverse_grid_host(sc._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "f")
        ms = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); lib.check(lib.dll.hgb_traverse_grid_host(sc._h, h_rays.data_ptr(), h_hits.data_ptr(), n, HIT_PRIM_ID), "f"); b.record(); b.synchronize()
            ms.append(a.elapsed_time(b))
        ok = bool(np.array_equal(h_hits.numpy().view(np.int32)[:, 0], want["id"]))
        out[f"{kind}_{label}"] = {"ms_median": round(float(np.median(ms)), 4), "ms_min": round(min(ms), 4), "mrays_s": round(n / np.median(ms) / 1e3, 1), "ok": ok}
    for chunk in (192, 256, 320, 384, 448, 512, 768):
        lib.set_option("host_frame_chunk_rays", chunk * 1024); run(f"chunk{chunk}K")
print(json.dumps(out))
