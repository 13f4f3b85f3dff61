"""How much of one BSIM4 evaluation does not depend on the bias point?  Runs the single-source evaluator on the
taint-tracking scalar (xb::TaintReal, tests/host_mirror libxb_host_taint.so) over the BASELINE config-2 operating points:
operations whose result depends only on the model card / bin / instance records could move into those records;
divisions of a bias-dependent value by a bias-independent divisor could become multiplications by a stored reciprocal."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from b4_common import HostMirror, HOST_SO  # noqa: E402
from xyce_b200 import workloads as wl  # noqa: E402

hm = HostMirror()
hm.lib = C.CDLL(HOST_SO.replace("libxb_host.so", "libxb_host_taint.so"))
w = wl.inverter_array(500)
tot = np.zeros(13)
n = w["n_inst"]
sto = w["store"].reshape(22, n)
for i in range(n):
    rec = dict(model_d=w["model_d"][w["model_idx"][i]], model_i=w["model_i"][w["model_idx"][i]],
               size_d=w["size_d"][w["size_idx"][i]], inst_d=w["inst_d"][i], inst_i=w["inst_i"][i])
    rec = {k: np.ascontiguousarray(v) for k, v in rec.items()}
    V = np.array([w["x"][g] if g >= 0 else 0.0 for g in w["lids"][i]])
    hm.eval(rec, dict(transient=1, newtonIter=1), V, np.ascontiguousarray(sto[:13, i]), True, w["von"][i])
    c = (C.c_ulonglong * 13)()
    hm.lib.xbh_taint_counts(c)
    tot += np.array(list(c), dtype=float)
tot /= n
names = ["add_sub", "mul", "div", "sqrt", "exp", "log"]
out = dict(bias_independent=dict(zip(names, tot[:6].tolist())), bias_dependent=dict(zip(names, tot[6:12].tolist())),
           divisions_of_a_bias_dependent_value_by_a_bias_independent_divisor=float(tot[12]))
print(json.dumps(out, indent=1))
