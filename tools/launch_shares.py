"""profiles/r01_kernel_shares.md from the committed ncu launch list and bench line:
   python tools/launch_shares.py profiles/r01_bench_launches.csv profiles/r01_bench_final.json > profiles/r01_kernel_shares.md"""
import csv, json, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
SWEEP = ("dgemm_tn", "coef_", "impact_", "sv_", "prior_kernel", "bump_kernel", "ldlt_state", "record_kernel", "invert_diag")
def short(n):
    n = re.sub(r"\(.*", "", n).replace("bvhar::", "").replace("<unnamed>::", "")
    return n[:70]
agg, other = OrderedDict(), OrderedDict()
for r in rows:
    nm, ns = short(r[ki]), float(r[vi].replace(",", ""))
    tgt = agg if any(s in nm for s in SWEEP) else other
    c, t = tgt.get(nm, (0, 0.0)); tgt[nm] = (c + 1, t + ns)
tot = sum(t for _, t in agg.values())
bench = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
print("# r01: kernel shares of the bench step - ncu launch list vs in-bench CUDA events\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3` (profiles/r01_bench_launches.csv;")
print("per-launch times under ncu are cold-cache and serialised: the SHARES are what must agree with the event-timed pass of `bench.py`).\n")
print("Sweep kernels (everything else in the capture - setup kernels, the cuBLAS / DMMA-peak launches that measure the fp64 denominators - is listed below):\n")
print("| kernel | launches | ncu total [us] | share of sweep kernels |\n|---|---|---|---|")
for nm, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {nm} | {c} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |")
print("\nnot part of the sweep: " + "; ".join(f"{nm} x{c} ({t / 1e6:.1f} ms)" for nm, (c, t) in other.items()) + "\n")
km = bench["roofline"]["kernel_ms_per_step"]; ks = sum(km.values())
print(f"| bench.py event-timed kernel ({bench['steps']} steps, {bench['config']['chains_total']} chains) | ms / step | share |\n|---|---|---|")
for nm, v in sorted(km.items(), key=lambda kv: -kv[1]):
    print(f"| {nm} | {v:.3f} | {100 * v / ks:.1f}% |")
print(f"\nG1 (`gram_S_dmma`, the roofline kernel): {100 * km['gram_S_dmma'] / ks:.1f}% of the event-timed kernel sum; "
      f"graph-replayed step {bench['ms_per_step']:.3f} ms (branches overlap), roofline.share_of_step {bench['roofline']['share_of_step']:.3f}.")
