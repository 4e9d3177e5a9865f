"""Per-instruction table (executed count, avg active threads, stall samples) of one kernel from
`ncu -i X.ncu-rep --page source --csv --print-source sass`.
usage: ncu_sass_table.py source.csv <kernel substring> [occurrence]"""
import csv, sys
path, want = sys.argv[1], sys.argv[2]
occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None and r:
        cur["rows"].append(r)
sel = [b for b in blocks if want in b["name"]][occ]
hdr = sel["rows"][0]; idx = {h: i for i, h in enumerate(hdr)}
tot_inst = sum(int(r[idx["Instructions Executed"]]) for r in sel["rows"][1:])
tot_samp = sum(int(r[idx["# Samples"]]) for r in sel["rows"][1:])
print(sel["name"][:100]); print("total warp-instr", tot_inst, "samples", tot_samp)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: 0 for h in stalls}
print(f"{'off':>5} {'inst':>10} {'%inst':>6} {'thr':>5} {'samp%':>6}  top stalls | sass")
base = int(sel["rows"][1][0], 16)
for r in sel["rows"][1:]:
    inst = int(r[idx["Instructions Executed"]]); samp = int(r[idx["# Samples"]])
    for h in stalls: agg[h] += int(r[idx[h]])
    top = sorted(((int(r[idx[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    tops = " ".join(f"{n}:{c}" for c, n in top if c)
    print(f"{int(r[0],16)-base:5x} {inst:10d} {100*inst/tot_inst:6.2f} {r[idx['Avg. Threads Executed']]:>5} {100*samp/max(tot_samp,1):6.2f}  {tops:28s} | {r[1].strip()}")
print("stall totals:", {k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
