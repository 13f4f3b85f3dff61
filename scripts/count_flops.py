"""Measure the executed fp64 operation count of one BSIM4 evaluation (roofline unit U1) with the
operation-counting scalar type, over the BASELINE config-2 operating-point distribution."""
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
hm.lib = C.CDLL(HOST_SO.replace("libxb_host.so", "libxb_host_count.so"))
w = wl.inverter_array(1000)
tot = np.zeros(7)
n = w["n_inst"]
sto = w["store"].reshape(22, n)
for i in range(n):
    rec = dict(model_d=w["model_d"][w["model_idx"][i]], model_i=w["model_i"][w["model_idx"][i]],
               size_d=w["size_d"][w["size_idx"][i]], inst_d=w["inst_d"][i], inst_i=w["inst_i"][i])
    rec = {k: np.ascontiguousarray(v) for k, v in rec.items()}
    V = np.array([w["x"][g] if g >= 0 else 0.0 for g in w["lids"][i]])
    hm.eval(rec, dict(transient=1, newtonIter=1), V, np.ascontiguousarray(sto[:13, i]), True, w["von"][i])
    c = (C.c_ulonglong * 7)()
    hm.lib.xbh_op_counts(c)
    tot += np.array(list(c), dtype=float)
names = ["add_sub", "mul", "div", "sqrt", "exp", "log", "compare"]
per = dict(zip(names, (tot / n).tolist()))
per["flops_total"] = float(sum(per[k] for k in names[:6]))
per["note"] = "mean executed fp64 operations per BSIM4 evaluation, each + - * / sqrt exp log counted as 1; " \
              "2000 instances of the config-2 inverter array, x ~ U(0,VDD), pass-through limiters"
json.dump(per, open(os.path.join(ROOT, "xyce_b200", "data", "b4_flop_count.json"), "w"), indent=1)
print(per)
