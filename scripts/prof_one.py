"""One configuration of the BSIM4 evaluation kernel run a few times (target of ncu captures).
usage: prof_one.py n_inverters  (kernel variant through XYCE_B200_B4_* environment variables)"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState
n_inv = int(sys.argv[1])
w = wl.inverter_array(n_inv, store_noise=0.0)
eng = wl.build_engine(w)
ss = SolverState(transientFlag=1, newtonIter=1)
b = [eng.device_buffer(i) for i in range(11)]
eng.load_host(w["x"], ss)
for it in range(5):
    eng.update_state(b[0], b[9], b[10], b[7], b[8], ss)
eng.sync()
print("done")
