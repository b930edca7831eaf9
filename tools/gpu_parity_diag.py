"""Worst relative error per (stage, state) of the engine against the oracle at the BASELINE shapes (no assertions)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import baseline_configs as bc
import oracle_binding as ob
from bvhar_b200._engine import CudaFit
from test_gpu_baseline_parity import CASES, PRIOR_CFG, _names
from test_gpu_parity import relerr

sel = sys.argv[1:] or None
for cfg, prior, chains, windows in CASES:
    if sel and f"{cfg}-{prior}" not in sel and cfg not in sel:
        continue
    pk = bc.packed(cfg, chains=chains, windows=windows, iters=3, prior_cfg=PRIOR_CFG[prior]())
    o = ob.OracleFit(pk, faithful=False); g = CudaFit(pk)
    print(f"== {cfg} {prior}", flush=True)
    for it in (1, 2):
        for st in range(1, 10):
            o.run_stage(st, it); g.run_stage(st, it)
            for nm in _names(st, prior, bc.SV[cfg]):
                worst = 0.0
                for w in range(pk.n_windows):
                    for c in range(pk.n_chains):
                        try:
                            e = relerr(o.get_state(nm, w, c), g.get_state(nm, w, c))
                        except AssertionError as ex:
                            e = float("inf"); print("   !", nm, str(ex)[:80])
                        worst = max(worst, e)
                flag = " <-- " if worst >= 1e-10 else ""
                print(f"  sweep {it} stage {st} {nm:24s} {worst:.3e}{flag}", flush=True)
    o.close(); g.close()
