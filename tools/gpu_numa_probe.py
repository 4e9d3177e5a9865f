"""Does it matter on which CPUs (and so on which NUMA node) a single rank's page-locked frame buffers are first touched?
H2D 66 MB, D2H 33 MB and both at once, process unbound / bound to the CPUs NVML calls local to GPU 0 / bound to the others
(run under gpurun). One child process per mode: the binding has to precede the allocation."""
import json, os, subprocess, sys, time

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    mode = sys.argv[2]
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
    local = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
    allowed = os.sched_getaffinity(0)
    if mode == "local" and local & allowed:
        os.sched_setaffinity(0, local & allowed)
    elif mode == "remote" and allowed - local:
        os.sched_setaffinity(0, allowed - local)
    import torch
    n_in, n_out = 66355200, 33177600
    h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(1)
    d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def timed(fn, reps=20):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        return best
    def h2d():
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    def both():
        h2d(); d2h()
    res = {"mode": mode, "cpus_allowed": len(os.sched_getaffinity(0)), "cpus_local_to_gpu0": len(local), "cpus_total": os.cpu_count(),
           "h2d_ms": round(timed(h2d) * 1e3, 4), "d2h_ms": round(timed(d2h) * 1e3, 4), "both_ms": round(timed(both) * 1e3, 4)}
    res["h2d_gbs"] = round(n_in / res["h2d_ms"] / 1e6, 1)
    print(json.dumps(res))
    sys.exit(0)

for mode in ("unbound", "local", "remote", "unbound"):
    r = subprocess.run([sys.executable, __file__, "--child", mode], capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-400:], flush=True)
