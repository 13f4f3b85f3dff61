"""Time the specialised BSIM4 evaluation kernel of one library build (XYCE_B200_LIB) over block shapes and sizes and
report its deviation from the strict-parity variant.  usage: b4_exp_time.py <tag> [shapes "128x3,128x4"] [sizes "50000,500000"]"""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState

tag = sys.argv[1]
shapes = [tuple(int(v) for v in s.split("x")) for s in (sys.argv[2] if len(sys.argv) > 2 else "128x3,128x4").split(",")]
sizes = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "50000,500000").split(",")]
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
ss = SolverState(transientFlag=1, newtonIter=1)
for n_inv in sizes:
    w = wl.inverter_array(n_inv, store_noise=0.0)
    eng = wl.build_engine(w)
    stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
    ref = None
    if n_inv <= 50000:
        eng.set_option("b4_arith", 0); eng.set_option("b4_spec", 0)
        eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
        ref = eng.load_host(w["x"], ss)
    eng.set_option("b4_arith", 2); eng.set_option("b4_spec", 1)
    for t, mb in shapes:
        try:
            eng.set_option("b4_threads", t); eng.set_option("b4_minblocks", mb)
            eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
            out = eng.load_host(w["x"], ss)
        except Exception as e:
            print(json.dumps(dict(tag=tag, n_inst=w["n_inst"], threads=t, minblocks=mb, error=str(e)[:80])), flush=True)
            continue
        err = None
        if ref is not None:
            err = 0.0
            for k in ("f", "q", "dFdx", "dQdx"):
                sc = 1e-3 * np.max(np.abs(ref[k]))
                err = max(err, float(np.max(np.abs(out[k] - ref[k]) / np.maximum(np.abs(ref[k]), sc))))
        b = [eng.device_buffer(i) for i in range(11)]
        ts = []
        for it in range(13):
            flush.fill_(0.0)
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            e0.record(stream)
            eng.update_state(b[0], b[9], b[10], b[7], b[8], ss)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = np.array(ts[3:])
        print(json.dumps(dict(tag=tag, n_inst=w["n_inst"], threads=t, minblocks=mb, eval_ms=round(float(np.median(ts)), 5),
                              eval_ms_min=round(float(np.min(ts)), 5), dev_vs_strict=err,
                              gevals_per_s=round(w["n_inst"] / np.median(ts) * 1e-6, 4))), flush=True)
    del eng
