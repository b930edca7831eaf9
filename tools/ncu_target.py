"""Small driver for ncu: a few sweeps of the C3 workload at a reduced chain count."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bvhar_b200._engine import CudaFit
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 296
m = bench.build_model(chains, 16, 0, seed=1, device=0)
pk = m._pack(m.design_, m.response_, m._seeds(1, chains), record_mask=0)
f = CudaFit(pk)
f.set_profiling(True)   # eager launches (no graph) so that ncu sees plain kernel launches
f.step(3, False)
print("done", f.profile())
