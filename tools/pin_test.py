import torch, time, numpy as np
torch.cuda.init(); torch.zeros(1, device="cuda")
n = 3277 * 10**6 // 8
d = torch.zeros(n, dtype=torch.float64, device="cuda")
def T(f, name):
    torch.cuda.synchronize(); t=time.perf_counter(); r=f(); torch.cuda.synchronize(); print(f"{name:40s} {1e3*(time.perf_counter()-t):8.1f} ms", flush=True); return r
h = T(lambda: torch.empty(n, dtype=torch.float64, pin_memory=True), "cudaHostAlloc 3.3 GB")
T(lambda: h.copy_(d, non_blocking=True), "D2H pinned 3.3 GB")
T(lambda: h.copy_(d, non_blocking=True), "D2H pinned 3.3 GB again")
a = T(lambda: np.empty(n), "np.empty")
T(lambda: torch.from_numpy(a).copy_(d), "D2H pageable first touch")
T(lambda: torch.from_numpy(a).copy_(d), "D2H pageable again")
b = np.empty(n)
T(lambda: torch.cuda.cudart().cudaHostRegister(b.ctypes.data, b.nbytes, 0), "cudaHostRegister untouched 3.3 GB")
T(lambda: torch.from_numpy(b).copy_(d, non_blocking=True), "D2H registered")
T(lambda: torch.cuda.cudart().cudaHostUnregister(b.ctypes.data), "unregister")
h2 = T(lambda: torch.empty(n, dtype=torch.float64, pin_memory=True), "cudaHostAlloc 3.3 GB #2")
x = torch.empty(164*10**6//8, dtype=torch.float64).numpy()
T(lambda: torch.from_numpy(x).cuda(), "H2D pageable 164 MB")
