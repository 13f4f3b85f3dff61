"""xgpu_load_host_jr, pipelined, on BASELINE config 2: residual part by DMA vs stored straight into mapped pinned memory."""
import sys, os, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState
w = wl.inverter_array(50000, store_noise=0.0)
ss = SolverState(transientFlag=1, newtonIter=1)
ptr = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
eng = wl.build_engine(w)
n, nnz = eng.n, eng.nnz
hx = torch.tensor(w["x"], dtype=torch.float64).pin_memory()
out = {}
for mapped in (0, 1, 0, 1):
    eng.set_option("pipe_r_mapped", mapped)
    hr = torch.zeros(n, dtype=torch.float64).pin_memory(); hj = torch.zeros(nnz, dtype=torch.float64).pin_memory()
    def step():
        eng._chk(eng.lib.xgpu_load_host_jr(eng.h, ptr(hx), C.byref(ss), C.c_double(1e12), C.c_double(0.5), ptr(hr), ptr(hj)))
    for _ in range(10): step()
    ts = []
    for _ in range(60):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    out.setdefault(mapped, []).append((hr.clone(), hj.clone()))
    print("mapped %d: median %.1f us, min %.1f us" % (mapped, np.median(ts) * 1e6, np.min(ts) * 1e6), flush=True)
print("bitwise equal:", all(torch.equal(out[0][0][k], out[1][0][k]) for k in (0, 1)))
