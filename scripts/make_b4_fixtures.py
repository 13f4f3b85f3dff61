"""Generate xyce_b200/data/b4_inverter_records.npz by running the reference's own
Model::processParams4p82_ / Instance::processParams4p82_ / updateTemperature4p82_
(through oracle/_ref) on the benchmark model cards.  Needs /root/reference only to build
oracle/_ref; the output is committed."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
from oracle_ref import RefCircuit  # noqa: E402
from xyce_b200 import workloads as wl  # noqa: E402

c = RefCircuit(4)
c.add_model("nch", "NMOS", wl.NMOS_CARD)
c.add_model("pch", "PMOS", wl.PMOS_CARD)
c.add_instance("M:n", "nch", [1, 0, -1, -1], wl.NMOS_INST)
c.add_instance("M:p", "pch", [1, 0, 2, 2], wl.PMOS_INST)
c.finalize()
e = [c.export(0), c.export(1)]
out = {k: np.array([e[0][k], e[1][k]]) for k in ("model_d", "model_i", "size_d", "inst_d", "inst_i")}
for w, key in enumerate(("model_d", "model_i", "size_d", "inst_d", "inst_i")):
    out["names_" + key] = np.array(c.names(w))
np.savez(os.path.join(ROOT, "xyce_b200", "data", "b4_inverter_records.npz"), **out)
print({k: v.shape for k, v in out.items()})
