"""Two ray waves per frame on one GPU (run under gpurun): primary rays, one diffuse bounce, second traversal.

Device procedure (this library): the hits of the first wave stay in HBM, hgb_generate_bounce_rays writes the second
wave next to them, the second traversal reads it -- three launches, no PCIe traffic.
Host procedure (what a front end in the style of src/main.cpp:598-613 has to do with the reference): download
the hits, make the second wave on the CPU, upload it, trace again; timed without the CPU generation (a lower bound). Timed with the host clock around
synchronised regions (the C ABI exposes event timing for traversal only); second-wave hits are compared."""
import json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hagrid_b200 import HIT_DTYPE, HIT_PRIM_ID, RAY_DTYPE, Library, Scene, scenes

ITERS = 20
tris = scenes.sponza262k()
primary = scenes.default_view(tris)
n = primary.shape[0]
lo, hi = scenes.scene_bbox(tris)
diag = float(np.linalg.norm(hi - lo))
out = {"scene": "sponza262k", "rays_per_wave": int(n)}


def wall(lib, fn, iters=ITERS):
    fn(); lib.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    lib.synchronize()
    return (time.perf_counter() - t0) * 1e3 / iters


lib = Library()
sc = Scene(tris, keep_alive=True, lib=lib)
sc.build_all(0.15, 3.0)
sc.setup_traversal()
d_rays, d_hits, d_second, d_hits2 = sc.device_alloc(n * 32), sc.device_alloc(n * 16), sc.device_alloc(n * 32), sc.device_alloc(n * 16)
sc.to_device(d_rays, primary)


def device_frame():
    sc.traverse(d_rays, d_hits, n, HIT_PRIM_ID)
    sc.bounce_rays_device(d_rays, d_hits, n, 1e-3 * diag, diag, 7, d_second)
    sc.traverse(d_second, d_hits2, n, HIT_PRIM_ID)


out["device_two_wave_ms"] = round(wall(lib, device_frame), 4)
bounce_ms = wall(lib, lambda: sc.bounce_rays_device(d_rays, d_hits, n, 1e-3 * diag, diag, 7, d_second), 200)
first = sc.to_host(np.empty(n, HIT_DTYPE), d_hits)
moved = 48 * n + 32 * n + 12 * int((first["id"] >= 0).sum())          # ray + hit in, ray out, three normal words gathered
out["bounce_kernel"] = {"ms": round(bounce_ms, 4), "algorithmic_GB_s": round(moved / bounce_ms / 1e6, 1),
                        "hit_fraction": round(float((first["id"] >= 0).mean()), 4)}
second_dev = sc.to_host(np.empty(n, RAY_DTYPE), d_second)
hits_dev = sc.to_host(np.empty(n, HIT_DTYPE), d_hits2)
t0 = time.perf_counter()
second_cpu = scenes.bounce_rays_f32(tris, primary, first, 1e-3 * diag, diag, 7)
out["cpu_bounce_generation_numpy_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
out["second_wave_identical_to_host_restatement"] = bool(second_dev.tobytes() == second_cpu.tobytes())
for p in (d_rays, d_hits, d_second, d_hits2):
    sc.device_free(p)
sc.close()

ref = Library(ROOT / "oracle/_ref/libhagrid_ref.so")
rs = Scene(tris, keep_alive=True, lib=ref)
rs.build_all(0.15, 3.0)
rs.setup_traversal()
d_rays, d_hits, d_second = rs.device_alloc(n * 32), rs.device_alloc(n * 16), rs.device_alloc(n * 32)
rs.to_device(d_rays, primary)
host_hits = np.empty(n, HIT_DTYPE)


second_host = scenes.bounce_rays_f32(tris, primary, first, 1e-3 * diag, diag, 7)


def host_frame():
    # generation itself left out: download, upload and the two traversals bound the procedure from below for any
    # CPU generator (the numpy one above is timed separately)
    rs.traverse(d_rays, d_hits, n, HIT_PRIM_ID)
    rs.to_host(host_hits, d_hits)
    rs.to_device(d_second, second_host)
    rs.traverse(d_second, d_hits, n, HIT_PRIM_ID)


out["reference_host_two_wave_without_generation_ms"] = round(wall(ref, host_frame, 10), 3)
out["second_wave_hits_identical"] = bool(rs.to_host(np.empty(n, HIT_DTYPE), d_hits).tobytes() == hits_dev.tobytes())
rs.device_free(d_second); rs.device_free(d_rays); rs.device_free(d_hits); rs.close()
print(json.dumps(out))
