"""xgpu_newton_step_host on BASELINE config 2 with the supply node inside the BTF blocks / declared as border."""
import sys, os, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState
w = wl.inverter_array(50000, store_noise=0.0)
ss = SolverState(transientFlag=1, newtonIter=1)
ptr = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
for border in (0, 1):
    eng = wl.build_engine(w)
    n = eng.n
    if border:
        eng.border_set(1)
    eng.set_state(0, w["store"]); eng.set_state(1, w["store"])
    hx = torch.tensor(w["x"], dtype=torch.float64).pin_memory(); hdx = torch.zeros(n, dtype=torch.float64).pin_memory()
    def step():
        eng._chk(eng.lib.xgpu_newton_step_host(eng.h, ptr(hx), C.byref(ss), C.c_double(1e12), C.c_double(0.5), None, ptr(hdx), None))
    for _ in range(4): step()
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    for _ in range(20): step()
    dt = (time.perf_counter() - t0) / 20
    print("border", border, "ms per Newton step %.4f" % (dt * 1e3), "launches per step", (eng.launch_count() - l0) / 20, "max|dx| %.3e" % float(hdx.abs().max()), flush=True)
    if border == 0: ref = hdx.clone()
    else: print("max |dx_border - dx_plain| %.3e" % float((hdx - ref).abs().max()))
    eng.close()
