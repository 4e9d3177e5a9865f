"""Turns an ncu report into the committed summaries under profiles/:
  <out>_metrics.csv   selected raw metrics, one column per captured launch
  <out>_sass_<k>.txt  per-instruction table (executed, active threads, stall samples) of launch k
usage: ncu_summarize.py report.ncu-rep profiles/r01_name [kernel substring for the sass tables]"""
import csv, io, re, subprocess, sys
from pathlib import Path

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]

def main():
    rep, out = sys.argv[1], sys.argv[2]
    want = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out + "_metrics.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{k}" for k in range(len(data))])
        w.writerow(["kernel", ""] + [d[idx["Kernel Name"]][:90] for d in data])
        for m in METRICS:
            if m in idx:
                w.writerow([m, units[idx[m]]] + [d[idx[m]] for d in data])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
    tmp = Path(out + "_source.tmp.csv"); tmp.write_text(src)
    names = [d[idx["Kernel Name"]] for d in data]
    seen = {}
    for k, name in enumerate(names):
        if want and want not in name: continue
        occ = seen.get(name, 0); seen[name] = occ + 1
        if occ: continue                      # one table per distinct kernel
        base = re.search(r"(\w+)\s*(<|\()", name.replace("void ", "").replace("<unnamed>::", ""))
        tab = subprocess.run([sys.executable, str(Path(__file__).with_name("ncu_sass_table.py")), str(tmp), base.group(1) if base else name[:40], "0"],
                             capture_output=True, text=True).stdout
        Path(f"{out}_sass_{k}.txt").write_text(tab)
    tmp.unlink()

if __name__ == "__main__":
    main()
