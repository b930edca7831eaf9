"""Compact per-kernel table from an .ncu-rep (read on the CPU box):  python tools/ncu_summary.py rep.ncu-rep > profiles/x.md"""
import csv, subprocess, sys
rep = sys.argv[1]
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
     ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe %"),
     ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
     ("launch__registers_per_thread", "regs"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
     ("launch__grid_size", "grid"), ("launch__block_size", "block")]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, units = r[0], r[1]
idx = {n: i for i, n in enumerate(h)}
print(f"ncu --set full --clock-control none, report `{rep.split('/')[-1]}` (per launch; cold-cache, serialised replays)\n")
print("| kernel | " + " | ".join(f"{lab} [{units[idx[m]]}]" if units[idx[m]] else lab for m, lab in M if m in idx) + " |")
print("|---|" + "---|" * sum(m in idx for m, _ in M))
for row in r[2:]:
    name = row[idx["Kernel Name"]].replace("bvhar::", "").replace("<unnamed>::", "").replace("void ", "").split("(")[0]
    vals = []
    for m, _ in M:
        if m not in idx: continue
        v = row[idx[m]]
        try: v = f"{float(v):.4g}"
        except ValueError: pass
        vals.append(v)
    print(f"| {name} | " + " | ".join(vals) + " |")
