"""Summarise the per-instruction page of an ncu report: executed warp instructions,
average active threads and stall samples per SASS instruction.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K --launch-count 1 > src.csv; python tools/ncu_source_summary.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
ia, ie, it, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[h + 1:]:
    if len(r) <= it or not r[ie].isdigit():
        continue
    data.append((r[0][-4:], r[ia].strip(), int(r[ie]), int(r[it]), int(r[iss])))
tot = sum(d[2] for d in data)
tots = sum(d[4] for d in data)
print("total warp-inst", tot, "thread-inst", sum(d[3] for d in data), "samples", tots)
for i, (a, s, e, t, sm) in enumerate(data):
    print(f"{i:3d} {a} {e/1e6:7.2f}M act {t/max(e,1):5.1f} smp {100.0*sm/max(tots,1):5.2f}%  {s[:64]}")
