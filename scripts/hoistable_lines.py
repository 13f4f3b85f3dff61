"""Where are the bias-independent operations of one BSIM4 evaluation?  Companion of count_hoistable.py: the taint-tracking
scalar reports every operation whose operands do not depend on the bias point (and every division by a bias-independent
divisor) with its call stack; the stacks are resolved to source lines of the evaluator with addr2line.
Tool build: g++ -O0 -g -fno-inline -DXB_TAINT -DXB_TAINT_TRACE tests/host_mirror/b4_host.cpp -> argv[1]."""
import collections
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from b4_common import HostMirror  # noqa: E402
from xyce_b200 import workloads as wl  # noqa: E402

lib_path = sys.argv[1]
hm = HostMirror()
hm.lib = C.CDLL(lib_path)
w = wl.inverter_array(8)
n = w["n_inst"]
sto = w["store"].reshape(22, n)
for i in range(n):
    rec = dict(model_d=w["model_d"][w["model_idx"][i]], model_i=w["model_i"][w["model_idx"][i]],
               size_d=w["size_d"][w["size_idx"][i]], inst_d=w["inst_d"][i], inst_i=w["inst_i"][i])
    rec = {k: np.ascontiguousarray(v) for k, v in rec.items()}
    V = np.array([w["x"][g] if g >= 0 else 0.0 for g in w["lids"][i]])
    hm.eval(rec, dict(transient=1, newtonIter=1), V, np.ascontiguousarray(sto[:13, i]), True, w["von"][i])
hm.lib.xbh_taint_dump(b"/tmp/taint_stacks.txt")
lines = open("/tmp/taint_stacks.txt").read().split("\n")
base = int(lines[0].split()[1], 16)
events = []
addrs = set()
for ln in lines[1:]:
    if not ln.strip():
        continue
    t = ln.split()
    fr = [int(a, 16) - base for a in t[2:] if a not in ("(nil)", "0x0")]
    events.append((int(t[0]), int(t[1]), fr))
    addrs.update(fr)
addrs = sorted(addrs)
out = subprocess.run(["addr2line", "-e", lib_path] + [hex(a - 1) for a in addrs], capture_output=True, text=True).stdout.split("\n")
where = dict(zip(addrs, out))
KIND = ["add", "mul", "div", "sqrt", "exp", "log", "div_by_const"]
hist = collections.Counter()
for cnt, kind, fr in events:
    loc = "?"
    for a in fr:
        f = where.get(a, "?")
        if "bsim4_" in f or "xb_common" in f:
            loc = os.path.basename(f.split(" ")[0])
            break
    hist[(loc, KIND[kind])] += cnt / n
tot = collections.Counter()
for (loc, kind), c in sorted(hist.items()):
    print("%-34s %-12s %.2f" % (loc, kind, c))
    tot[kind] += c
print(dict(tot))
