"""Time the mode-specialised BSIM4 kernel at high-occupancy block shapes (registers capped at 80-168 per thread)
for several group sizes, and check each shape against the strict-parity variant.
usage: b4_occupancy.py out.json n_inverters..."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState

out_path = sys.argv[1]
# (128, 5), (128, 6), (64, 11) need XB_B4_LAUNCH_SHAPES extended with those shapes (they were, for profiles/r01_b4_occupancy.json)
EXTRA = [(2, t, mb) for t, mb in ((128, 5), (128, 6), (64, 11))] if os.environ.get("XB_EXTRA_SHAPES") else []
res = []
for n_inv in [int(a) for a in sys.argv[2:]]:
    w = wl.inverter_array(n_inv, store_noise=0.0)
    eng = wl.build_engine(w)
    stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
    ss = SolverState(transientFlag=1, newtonIter=1)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    ref = None
    for (arith, t, mb) in [(0, 128, 2), (2, 128, 2), (2, 128, 3), (2, 128, 4)] + EXTRA:
        eng.set_option("b4_arith", arith); eng.set_option("b4_threads", t); eng.set_option("b4_minblocks", mb)
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
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            e0.record(stream)
            eng.update_state(b[0], b[9], b[10], b[7], b[8], ss)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = np.array(ts[3:])
        r = dict(instances=w["n_inst"], arith=arith, threads=t, minblocks=mb, warps_per_sm=t * mb // 32, eval_ms=float(np.median(ts)),
                 eval_ms_min=float(np.min(ts)), max_rel_dev_vs_parity=err, evals_per_s_kernel=w["n_inst"] / (np.median(ts) * 1e-3))
        res.append(r); print(json.dumps(r), flush=True)
    eng.close()
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)
