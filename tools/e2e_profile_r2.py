"""Where the end-to-end time of VharBayes.fit() goes (C3, 1 024 chains, W + K sweeps, default record policy): five fits in a row."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from bvhar_b200._engine import CudaFit
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K, W = 20, 5
for rep in range(5):
    m = bench.build_model(chains, W + K, W, seed=2000 + rep, device=0)
    t = [time.perf_counter()]
    def lap(name):
        t.append(time.perf_counter()); print(f"   {name:34s} {1e3 * (t[-1] - t[-2]):8.1f} ms", flush=True)
    print(f"fit {rep}")
    pk = m._pack(m.design_, m.response_, m._seeds(1, chains)); lap("pack (host)")
    if rep == 4: os.environ["BVHAR_TIMING"] = "1"
    f = CudaFit(pk); lap("fit_create (H2D, alloc, setup)")
    f.run(None); lap("fit_run")
    tot = 0
    for nm in f.record_names():
        a = f.record_all(nm); tot += a.nbytes
    lap(f"record_all x{len(f.record_names())} {tot / 1e6:.0f} MB")
    f.record_mean("alpha_record"); f.record_mean("alpha_sparse_record"); f.record_mean("c_record"); f.record_mean("c_sparse_record"); lap("record_mean x4")
    f.close(); lap("close")
    print(f"   total {1e3 * (t[-1] - t[0]):.1f} ms")
    t0 = time.perf_counter(); m.fit(); print(f"   full fit() {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
