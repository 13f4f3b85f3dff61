"""Small runs of the one-launch assembly (rail chunks + ticket finisher), the LU kernels and the transient driver with
Gear / DCOP: target of compute-sanitizer memcheck and racecheck."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState
ss = SolverState(transientFlag=1, newtonIter=1)
for n_inv in (3, 700, 2100):           # 2100 inverters: several chunks per rail destination
    w = wl.inverter_array(n_inv, store_noise=0.3)
    eng = wl.build_engine(w)
    outs = []
    for rep in range(2):               # same carried state (store, von) before each pass
        eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
        outs.append(eng.load_host(w["x"], ss))
    a, b = outs
    for k in a:
        assert np.array_equal(a[k], b[k]), k          # bitwise reproducible, ticket counters reset
    eng.close()
w = wl.ring_oscillator_array(3, 11)
eng = wl.build_engine(w)
for method in (7, 8):
    r = eng.tran_run(w["x"], 2e-10, 1e-12, [0, 1], method=method)
    assert r["rc"] == 0, r
eng.close()
print("sanitize run ok")
