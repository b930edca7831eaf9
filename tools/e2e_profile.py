"""Where the end-to-end time of VharBayes.fit() goes (C3, 1024 chains, few sweeps)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from bvhar_b200._abi import REC_ALL
from bvhar_b200._engine import CudaFit
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K, W = 20, 3
t = [time.perf_counter()]
def lap(name):
    t.append(time.perf_counter()); print(f"{name:28s} {1e3 * (t[-1] - t[-2]):9.1f} ms", flush=True)
m = bench.build_model(chains, W + K, W, seed=2000, device=0, record_mask=REC_ALL); lap("build_model (untimed)")
pk = m._pack(m.design_, m.response_, m._seeds(1, chains)); lap("pack (host)")
f = CudaFit(pk); lap("fit_create (H2D, alloc)")
f.run(None); lap("fit_run")
names = f.record_names()
for nm in names:
    a = f.record_all(nm); lap(f"record_all {nm} {a.nbytes / 1e6:.0f} MB")
f.close(); lap("close")
t0 = time.perf_counter(); m.fit(); print("second full fit()", 1e3 * (time.perf_counter() - t0), "ms")
