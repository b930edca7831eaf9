import sys, os, time, cProfile, pstats, gc
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bvhar_b200._abi import REC_ALL
m = bench.build_model(1024, 23, 3, seed=3000, device=0, record_mask=REC_ALL); m.fit(); del m; gc.collect()
m = bench.build_model(1024, 23, 3, seed=2000, device=0, record_mask=REC_ALL)
pr = cProfile.Profile(); t0 = time.perf_counter(); pr.enable(); m.fit(); pr.disable(); print("fit() total", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
