"""Time BSIM4 kernel variants (arithmetic x block shape x record access x lock-step) on BASELINE config 2
and check each against the strict-parity variant.  usage: b4_variants.py [n_inverters] [out.json]"""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState

n_inv = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/b4_variants.json"
w = wl.inverter_array(n_inv, store_noise=0.0)
eng = wl.build_engine(w)
stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
ss = SolverState(transientFlag=1, newtonIter=1)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
SHAPES = [(128, 2), (128, 3), (128, 4), (128, 5), (256, 1), (256, 2), (384, 1), (512, 1)]
variants = [dict(arith=0, threads=128, minblocks=2, uniform=0, lockstep=0)]
for uni, ls in ((0, 0), (1, 0), (1, 1)):
    for t, mb in SHAPES:
        variants.append(dict(arith=2, threads=t, minblocks=mb, uniform=uni, lockstep=ls))
ref = None
res = []
for v in variants:
    for k in ("arith", "threads", "minblocks", "uniform", "lockstep"):
        eng.set_option("b4_" + k, v[k])
    eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
    out = eng.load_host(w["x"], ss)
    if ref is None:
        ref = out
    err = 0.0
    for k in ("f", "q", "dFdx", "dQdx"):
        sc = 1e-3 * np.max(np.abs(ref[k]))
        err = max(err, float(np.max(np.abs(out[k] - ref[k]) / np.maximum(np.abs(ref[k]), sc))))
    b = [eng.device_buffer(i) for i in range(11)]
    ts = []
    for it in range(14):
        flush.fill_(0.0)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        eng.update_state(b[0], b[9], b[10], b[7], b[8], ss)
        e1.record(stream)
        eng.load_vectors(b[1], b[2], b[3], b[4]); eng.load_matrices(b[5], b[6])
        e2.record(stream)
        torch.cuda.synchronize()
        ts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
    ts = np.array(ts[3:])
    r = dict(v, eval_ms=float(np.median(ts[:, 0])), eval_ms_min=float(np.min(ts[:, 0])), asm_ms=float(np.median(ts[:, 1])),
             max_rel_dev_vs_parity=err, evals_per_s_kernel=w["n_inst"] / (np.median(ts[:, 0]) * 1e-3))
    res.append(r); print(json.dumps(r), flush=True)
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)
