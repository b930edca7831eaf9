"""Driver for ncu on a BASELINE config sample: eager launches of a few sweeps.  python tools/ncu_cfg.py C4 128 [sweeps]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.argv = [sys.argv[0]] + sys.argv[1:]
cfg, units = sys.argv[1], int(sys.argv[2])
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
import importlib.util
spec = importlib.util.spec_from_file_location("cfgmod", os.path.join(ROOT, "tools", "gpu_bench_cfg.py"))
# reuse model() without running the script body
src = open(os.path.join(ROOT, "tools", "gpu_bench_cfg.py")).read().split("cfg = sys.argv[1]")[0]
ns = {"__file__": os.path.join(ROOT, "tools", "gpu_bench_cfg.py")}
exec(compile(src, "gpu_bench_cfg_head", "exec"), ns)
from bvhar_b200._engine import CudaFit
m, roll = ns["model"](cfg, units)
if roll is None:
    pk = m._pack(m.design_, m.response_, m._seeds(1, units), record_mask=0)
else:
    y, nwin = roll
    x_tot, y_tot = m._design_of(y)
    n0 = m.response_.shape[0]
    pk = m._pack(x_tot, y_tot, m._seeds(nwin, 4), win_row0=list(range(nwin)), win_nrow=[n0] * nwin, record_mask=0)
f = CudaFit(pk)
f.set_profiling(True)
f.step(K, False)
print("done", {k: round(v[1] / v[0], 3) for k, v in f.profile().items()})
