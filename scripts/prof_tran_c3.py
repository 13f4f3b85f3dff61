"""BASELINE config 3 (4 950 ring oscillators x 101 stages) .TRAN, short: the run ncu lists the launches of
(profiles/r02_launches_tran_c3_v3.csv).  usage: prof_tran_c3.py [tstop]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl  # noqa: E402

tstop = float(sys.argv[1]) if len(sys.argv) > 1 else 4e-11
w = wl.ring_oscillator_array(4950, 101)
eng = wl.build_engine(w)
eng.tran_run(w["x"], 1e-11, 1e-12, [0])          # warm-up: LU analysis, allocations
r = eng.tran_run(w["x"], tstop, 1e-12, [0])
print({k: r["stats"][k] for k in ("accepted", "rejected", "newton_iters", "residual_loads", "lu_refactors", "linear_solves")})
