"""Source-line hot spots of one kernel from an .ncu-rep:  python tools/ncu_hot.py rep.ncu-rep <kernel regex> [top]
Prints the share of warp-stall samples per CUDA source line (needs -lineinfo and --import-source on)."""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{pat}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; lines = []
first_kernel = None
for r in rows:
    if len(r) == 2 and r[0] == "Function Name":
        if first_kernel is None: first_kernel = r[1]
        active = r[1] == first_kernel
        continue
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0].isdigit() and r[2] == "-":  # a source line summary row
        try: lines.append((int(r[4]), cur, int(r[0]), r[1].strip(), int(r[7])))
        except ValueError: pass
tot = sum(l[0] for l in lines) or 1
print(first_kernel, "total samples", tot)
for s, f, ln, src, inst in sorted(lines, key=lambda x: -x[0])[:top]:
    print(f"{100 * s / tot:5.1f}%  {f}:{ln:<4d} inst={inst:<9d} {src[:110]}")
