#!/bin/bash
D=gpurun_out
# launches of one C3 sweep: bump, prior x2, normals, G2 (bulk), G1 (bulk), pbuild, chol, seq, G3 (bulk) ...: capture the bulk kernels of the 2nd sweep
ncu --set full --clock-control none --import-source on -k regex:'dgemm_tn_bulk' -s 3 -c 3 -o /tmp/s3_g python tools/ncu_target.py 1024 > $D/s3_ncu_g.log 2>&1; tail -1 $D/s3_ncu_g.log
python tools/ncu_summary.py /tmp/s3_g.ncu-rep
ncu -i /tmp/s3_g.ncu-rep --page details --csv | grep -E "No Eligible|Eligible Warps|Issued Warp|Achieved Occupancy|L1/TEX Hit|Duration" | cut -d, -f1,12-14
python - <<'PY'
import csv, subprocess
out = subprocess.run(["ncu", "-i", "/tmp/s3_g.ncu-rep", "--page", "source", "--print-source", "cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern = -1; lines = {}
for r in rows:
    if len(r) == 2 and r[0] == "Function Name": kern += 1; continue
    if len(r) > 6 and r[0].isdigit() and r[2] == "-":
        try: lines.setdefault(kern, []).append((int(r[4]), int(r[0]), r[1].strip()[:100]))
        except ValueError: pass
for kk, ls in lines.items():
    tot = sum(l[0] for l in ls) or 1
    print("kernel launch", kk, "samples", tot)
    for s_, ln, src in sorted(ls, reverse=True)[:10]: print(f"  {100*s_/tot:5.1f}%  :{ln:<4d} {src}")
PY
