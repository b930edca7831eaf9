"""Diagnostic: stage-by-stage comparison of the CUDA engine against the CPU oracle (run on a GPU box)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_binding as ob
import problems as pr
from bvhar_b200._engine import CudaFit

STATES = {
    1: ["prior_alpha_prec"], 3: ["sqrt_sv"], 4: ["coef_mat", "sparse_coef"], 5: ["prior_chol_prec"],
    6: ["latent_innov"], 8: ["contem_coef", "chol_lower", "sparse_contem"], 9: [],
}
HYPER = {
    "SSVS": ["coef_dummy", "coef_slab", "coef_weight", "spike_scl", "contem_dummy", "contem_slab", "contem_weight", "contem_spike_scl"],
    "Horseshoe": ["local_lev", "latent_local", "shrink_fac", "group_lev", "latent_group", "global_lev", "contem_local_lev", "latent_contem_local", "contem_global_lev"],
    "HMN": ["own_lambda", "contem_lambda"],
    "NG": ["local_lev", "group_lev", "local_shape", "global_lev", "contem_fac", "contem_global_lev", "contem_shape"],
    "DL": ["local_lev", "latent_local", "group_lev", "global_lev", "dir_concen", "contem_local_lev", "latent_contem_local", "contem_global_lev", "contem_dir_concen"],
    "GDP": ["local_lev", "group_rate", "coef_gamma_shape", "coef_gamma_rate", "contem_fac", "contem_rate", "contem_gamma_shape", "contem_gamma_rate"],
    "Minnesota": [],
}

def relerr(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = np.maximum(np.abs(a), np.abs(b)) + 1e-300
    bad = ~(np.isfinite(a) & np.isfinite(b))
    e = np.abs(a - b) / np.maximum(den, 1e-12 * (1 + den.max()))
    e[bad & (np.isnan(a) == np.isnan(b))] = 0
    return float(np.nanmax(e)) if e.size else 0.0

def stage_check(prior, sv, verbose=True, **kw):
    pk, meta = pr.pack(prior=prior, sv=sv, **kw)
    o = ob.OracleFit(pk, faithful=False)
    g = CudaFit(pk)
    worst = 0.0
    for it in (1, 2):
        for st in range(1, 10):
            o.run_stage(st, it); g.run_stage(st, it)
            names = list(STATES.get(st, []))
            if st in (1, 5): names += HYPER[prior]
            if st == 9: names += (["lvol_draw", "lvol_sig", "lvol_init"] if sv else ["diag_vec"])
            if st == 3 and not sv: names = []
            for nm in names:
                for c in range(pk.n_chains):
                    try:
                        a = o.get_state(nm, 0, c); b = g.get_state(nm, 0, c)
                    except Exception as e:
                        print("   ", nm, "unavailable", e); continue
                    e = relerr(a, b)
                    worst = max(worst, e)
                    if verbose and (e > 1e-9 or not np.isfinite(e)):
                        i = int(np.nanargmax(np.abs(a - b))) if a.size else 0
                        print(f"  MISMATCH it={it} stage={st} {nm}[chain {c}] rel={e:.3e} idx={i} oracle={a[i] if a.size else None} gpu={b[i] if a.size else None}")
    print(f"{prior:10s} sv={sv} worst rel err {worst:.3e}")
    return worst

def sweep_check(prior, sv, n=3, **kw):
    pk, meta = pr.pack(prior=prior, sv=sv, iters=n, burn=0, **kw)
    o = ob.OracleFit(pk, faithful=False); g = CudaFit(pk)
    o.step(n, True); g.step(n, True)
    worst = 0
    for nm in g.record_names():
        for c in range(pk.n_chains):
            e = relerr(o.record(nm, 0, c), g.record(nm, 0, c))
            worst = max(worst, e)
            if e > 1e-7: print(f"  sweep MISMATCH {nm} chain {c}: {e:.3e}")
    print(f"{prior:10s} sv={sv} {n} sweeps worst rel err {worst:.3e}")
    return worst

if __name__ == "__main__":
    which = sys.argv[1:] or pr.PRIORS
    for prior in which:
        for sv in (True, False):
            try:
                stage_check(prior, sv)
                sweep_check(prior, sv)
            except Exception as e:
                import traceback; traceback.print_exc()
