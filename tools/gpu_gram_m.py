"""G1 time against chains per window (M = C k rows per window group) at the C4 shape: python tools/gpu_gram_m.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bvhar_b200 import _spec as sp, datasets
from bvhar_b200._engine import CudaFit
from bvhar_b200.model import VarBayes
nwin = 128
y = datasets.simulate_var_sv(404 + nwin, 50, 4)
for C in (1, 2, 3, 4, 5, 6, 8):
    m = VarBayes(y[:404], 4, C, 40, 0, 1, sp.NgConfig(), sp.SvConfig(), seed=1)
    x_tot, y_tot = m._design_of(y)
    n0 = m.response_.shape[0]
    pk = m._pack(x_tot, y_tot, m._seeds(nwin, C), win_row0=list(range(nwin)), win_nrow=[n0] * nwin, record_mask=0)
    f = CudaFit(pk)
    f.step(1, False); f.set_profiling(True); f.step(3, False); p = f.profile()
    ms = p["gram_S_dmma"][1] / p["gram_S_dmma"][0]
    gf = nwin * C * 50 * 400 * 201 * 202 / 1e9
    print(f"C={C} M={50 * C}: gram_S {ms:.3f} ms  {gf / ms:.2f} TFLOP/s  ({ms / nwin * 1e3:.1f} us per window)", flush=True)
    f.close()
