"""Run every compiled BSIM4 kernel variant once on a small array (target of compute-sanitizer memcheck)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState
w = wl.inverter_array(300, store_noise=0.3)
eng = wl.build_engine(w)
ss = SolverState(transientFlag=1, newtonIter=1)
for spec in (0, 1):
    for uni, ls in ((0, 0), (1, 0), (1, 1)):
        for t, mb in [(64, 4), (64, 6), (96, 4), (128, 2), (128, 3), (128, 4), (256, 1), (384, 1), (512, 1)]:
            for k, v in (("spec", spec), ("uniform", uni), ("lockstep", ls), ("threads", t), ("minblocks", mb)):
                eng.set_option("b4_" + k, v)
            out = eng.load_host(w["x"], ss)
            assert np.all(np.isfinite(out["f"]))
print("all variants ran")
