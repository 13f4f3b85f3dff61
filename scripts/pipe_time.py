"""xgpu_load_host_jr on BASELINE config 2: one pass vs the two-part pipeline at several first-part shares."""
import sys, os, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState
w = wl.inverter_array(50000, store_noise=0.0)
ss = SolverState(transientFlag=1, newtonIter=1)
ptr = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
for percent in [int(a) for a in (sys.argv[1:] or ["50"])]:
    eng = wl.build_engine(w, options={"pipe_percent": percent})
    n, nnz = eng.n, eng.nnz
    hx = torch.tensor(w["x"], dtype=torch.float64).pin_memory()
    hr = torch.zeros(n, dtype=torch.float64).pin_memory(); hj = torch.zeros(nnz, dtype=torch.float64).pin_memory()
    def step():
        eng._chk(eng.lib.xgpu_load_host_jr(eng.h, ptr(hx), C.byref(ss), C.c_double(1e12), C.c_double(0.5), ptr(hr), ptr(hj)))
    res = {}
    for pipe in (0, 1):
        eng.set_option("pipeline_host", pipe)
        for _ in range(5): step()
        ts = []
        for _ in range(30):
            t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
        res[pipe] = np.median(ts) * 1e6
    print("percent %d: one pass %.1f us, pipelined %.1f us, windows %s" % (percent, res[0], res[1], eng.pipe_info()), flush=True)
    eng.close()
