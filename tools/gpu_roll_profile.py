"""Where the time of the rolling out-of-sample runner goes for a given number of windows:  python tools/gpu_roll_profile.py 128"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from bvhar_b200._engine import outforecast
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m_w, pk_w, y_w = bench.roll_problem(0, 8, 4, 6, 2, first_window=1)
outforecast(pk_w, y_w, 1, 4, m_w._seeds(4))
for iters in (200, 20):
    t0 = time.perf_counter()
    m, pk, y = bench.roll_problem(0, nw, 4, iters, iters // 2, first_window=1)
    t1 = time.perf_counter()
    os.environ["BVHAR_TIMING"] = "1"
    fc, _ = outforecast(pk, y, 1, 4, m._seeds(4))
    os.environ.pop("BVHAR_TIMING")
    t2 = time.perf_counter()
    print(f"windows {nw} iters {iters}: problem + pack {t1 - t0:.3f} s, outforecast {t2 - t1:.3f} s", flush=True)
