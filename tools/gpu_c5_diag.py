"""C5 (k = 100, d = 301, LDLT, GDP, 256 chains): is a run reproducible?  Hash of the states of a few chains after every
sweep for several repetitions of the same fit; bench-like stepping (several sweeps per call, no sync in between)."""
import sys, os, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bvhar_b200 import _spec as sp, datasets
from bvhar_b200._engine import CudaFit
from bvhar_b200.model import VharBayes
chains = 256
m = VharBayes(datasets.simulate_vhar_sv(1022, 100), 5, 22, chains, 40, 0, 1, sp.GdpConfig(), sp.LdltConfig(), seed=1)
pk = m._pack(m.design_, m.response_, m._seeds(1, chains), record_mask=0)
NAMES = ("coef_mat", "prior_alpha_prec", "diag_vec", "contem_coef")
def snap(f, units):
    out = {}
    for u in units:
        for nm in NAMES:
            a = f.get_state(nm, 0, u)
            out[(u, nm)] = (hashlib.md5(a.tobytes()).hexdigest()[:8], float(np.nanmax(np.abs(a))), int(np.isnan(a).sum()))
    return out
units = (0, 100, 239, 240, 241, 255)
mode = sys.argv[1] if len(sys.argv) > 1 else "bench"
ref = None
for rep in range(4):
    f = CudaFit(pk)
    hist = []
    try:
        if mode == "bench":
            f.step(2, False); hist.append(snap(f, units))
            f.step(5, False); hist.append(snap(f, units))
            f.set_profiling(True); f.step(5, False); hist.append(snap(f, units))
        else:
            for s in range(12):
                f.step(1, False); hist.append(snap(f, units))
        print(f"rep {rep}: ok")
    except Exception as e:
        print(f"rep {rep}: FAILED after {len(hist)} phases: {str(e)[:90]}")
    if ref is None:
        ref = hist
    for ph, (a, b) in enumerate(zip(ref, hist)):
        for key in a:
            if a[key] != b[key]:
                print(f"   phase {ph} {key}: ref {a[key]} now {b[key]}")
    if hist:
        print("   last phase unit 240:", {nm: hist[-1][(240, nm)] for nm in NAMES})
    try:
        f.close()
    except Exception:
        pass
