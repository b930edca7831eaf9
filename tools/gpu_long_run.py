"""Long-run sanity of the sampler: thousands of sweeps on small batches of C3 / C4 / C5 / C1-shaped problems; no numerical
flag may be raised and every state must stay finite (the parity tests only cover the first sweeps)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
src = open(os.path.join(ROOT, "tools", "gpu_bench_cfg.py")).read().split("cfg = sys.argv[1]")[0]
ns = {"__file__": os.path.join(ROOT, "tools", "gpu_bench_cfg.py")}
exec(compile(src, "gpu_bench_cfg_head", "exec"), ns)
from bvhar_b200._engine import CudaFit
for cfg, units, sweeps in (("C3", 64, 3000), ("C2", 16, 3000), ("C1", 8, 3000), ("C4", 8, 400), ("C5", 4, 150)):
    m, roll = ns["model"](cfg, units)
    if roll is None:
        pk = m._pack(m.design_, m.response_, m._seeds(1, units), record_mask=0)
        U = units
    else:
        y, nwin = roll
        x_tot, y_tot = m._design_of(y)
        n0 = m.response_.shape[0]
        pk = m._pack(x_tot, y_tot, m._seeds(nwin, 4), win_row0=list(range(nwin)), win_nrow=[n0] * nwin, record_mask=0)
        U = nwin * 4
    f = CudaFit(pk)
    t0 = time.perf_counter()
    f.step(sweeps, False)  # sync() raises on a numerical flag
    dt = time.perf_counter() - t0
    worst = 0.0
    for u in range(min(U, 4)):
        w, c = (u // 4, u % 4) if roll is not None else (0, u)
        for nm in ("coef_mat", "contem_coef", "prior_alpha_prec"):
            a = f.get_state(nm, w, c)
            assert np.all(np.isfinite(a)), (cfg, nm, u)
            worst = max(worst, float(np.abs(a).max()))
    print(f"{cfg}: {U} units x {sweeps} sweeps in {dt:.2f} s, all finite (largest |state| {worst:.3g})", flush=True)
    f.close()
