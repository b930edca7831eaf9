import sys, os, time
sys.path.insert(0, "/root/repo")
import bench
from bvhar_b200._abi import REC_ALL
from bvhar_b200._engine import CudaFit
m = bench.build_model(1024, 23, 3, seed=2000, device=0, record_mask=REC_ALL)
for i in range(2):
    t=time.perf_counter(); pk = m._pack(m.design_, m.response_, m._seeds(1, 1024)); print("pack", time.perf_counter()-t)
    t=time.perf_counter(); f = CudaFit(pk); print("create", time.perf_counter()-t)
    t=time.perf_counter(); f.close(); print("close", time.perf_counter()-t)
