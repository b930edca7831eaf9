"""Per-kernel timing of the BASELINE.json configurations C1..C5 (bounded samples):  python tools/gpu_bench_cfg.py C4 [units]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bvhar_b200 import _spec as sp, datasets
from bvhar_b200._engine import CudaFit
from bvhar_b200.model import VarBayes, VharBayes

def model(cfg, chains):
    if cfg == "C1":
        return VarBayes(datasets.load_vix(), 5, chains, 40, 0, 1, sp.SsvsConfig(), sp.LdltConfig(), seed=1), None
    if cfg == "C2":
        return VharBayes(datasets.load_vix(), 5, 22, chains, 40, 0, 1, sp.HorseshoeConfig(), sp.SvConfig(), seed=1), None
    if cfg == "C3":
        return VharBayes(datasets.simulate_vhar_sv(1022, 20), 5, 22, chains, 40, 0, 1, sp.DlConfig(), sp.SvConfig(), seed=1), None
    if cfg == "C4":  # 4 chains per window, rolling windows of 404 rows
        nwin = max(1, chains // 4)
        y = datasets.simulate_var_sv(404 + nwin, 50, 4)
        m = VarBayes(y[:404], 4, 4, 40, 0, 1, sp.NgConfig(), sp.SvConfig(), seed=1)
        return m, (y, nwin)
    if cfg == "C5":
        return VharBayes(datasets.simulate_vhar_sv(1022, 100), 5, 22, chains, 40, 0, 1, sp.GdpConfig(), sp.LdltConfig(), seed=1), None
    raise SystemExit("cfg must be C1..C5")

cfg = sys.argv[1]
chains = int(sys.argv[2]) if len(sys.argv) > 2 else {"C1": 2, "C2": 64, "C3": 1024, "C4": 128, "C5": 256}[cfg]
K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
t0 = time.perf_counter()
m, roll = model(cfg, chains)
if roll is None:
    pk = m._pack(m.design_, m.response_, m._seeds(1, chains), record_mask=0)
    units = chains
else:
    y, nwin = roll
    x_tot, y_tot = m._design_of(y)
    n0 = m.response_.shape[0]
    pk = m._pack(x_tot, y_tot, m._seeds(nwin, 4), win_row0=list(range(nwin)), win_nrow=[n0] * nwin, record_mask=0)
    units = nwin * 4
f = CudaFit(pk)
print(f"{cfg}: units={units} k={pk.dim} d={pk.dim_design} n={int(pk.win_nrow[0])} setup {time.perf_counter() - t0:.2f}s", flush=True)
f.step(2, False)
f.step(K, False); ms = f.last_ms()
f.set_profiling(True); f.step(K, False); prof = f.profile()
print(f"graph: {ms / K:.3f} ms/sweep -> {units * K / (ms * 1e-3):.0f} unit-sweeps/s")
for nm, (cnt, tot) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"   {nm:14s} {tot / K:9.3f} ms/sweep")
