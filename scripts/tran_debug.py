import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np
import oracle_ref
from b4_common import ref_circuit_from_workload
from xyce_b200 import workloads as wl
w = wl.ring_oscillator_array(3, 31)
probes = [0, 1, 15, w["vdd"], w["branch"]]
ref = ref_circuit_from_workload(oracle_ref.RefCircuit, w); ref.set_flags(transient=1)
want = ref.tran_run(w["x"], 1e-9, 1e-12, probes, w["linear"], w["sources"])
for arith in (0, 2):
    eng = wl.build_engine(w); eng.set_option("b4_arith", arith)
    got = eng.tran_run(w["x"], 1e-9, 1e-12, probes)
    a, b = got["steps"], want["steps"]
    m = min(len(a), len(b))
    bad = np.where((a[:m, 2] != b[:m, 2]) | (a[:m, 4] != b[:m, 4]))[0]
    print("arith", arith, "attempts", len(a), len(b), "first mismatch", bad[:3])
    if len(bad):
        i = bad[0]
        for k in range(max(0, i - 2), min(m, i + 3)):
            print(k, "gpu", a[k], "ref", b[k])
    eng.close()
