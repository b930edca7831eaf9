"""Quick per-kernel timing of the C3 sweep (no e2e / cpu legs)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bvhar_b200._engine import CudaFit
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prior = sys.argv[2] if len(sys.argv) > 2 else "DL"
K = 10
m = bench.build_model(chains, 64, 0, seed=1, device=0, prior=prior)
pk = m._pack(m.design_, m.response_, m._seeds(1, chains), record_mask=0)
f = CudaFit(pk)
f.step(3, False)
f.step(K, False); ms = f.last_ms()
f.set_profiling(True); f.step(K, False); prof = f.profile()
print(f"chains={chains} prior={prior} graph: {ms / K:.3f} ms/sweep  -> {chains * K / (ms * 1e-3):.0f} chain-sweeps/s")
for nm, (cnt, tot) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"   {nm:14s} {tot / K:8.3f} ms/sweep")
print(f"   {'sum':14s} {sum(v[1] for v in prof.values()) / K:8.3f}")
