"""Warm end-to-end breakdown of VharBayes.fit() at the bench's C3 workload: python tools/e2e_profile2.py"""
import sys, os, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bvhar_b200._abi import REC_ALL
m = bench.build_model(1024, 23, 3, seed=2000, device=0, record_mask=REC_ALL)
m.fit(); m.param_ = None
os.environ["BVHAR_TIMING"] = "1"
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
m.fit()
pr.disable()
print("fit() total", time.perf_counter() - t0)
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
