import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState
for n_inv in (50000, 500000):
    w = wl.inverter_array(n_inv, store_noise=0.0)
    eng = wl.build_engine(w)
    stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
    ss = SolverState(transientFlag=1, newtonIter=1)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    b = [eng.device_buffer(i) for i in range(11)]
    ref = None
    for spec in (0, 1):
        eng.set_option("b4_spec", spec)
        eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
        out = eng.load_host(w["x"], ss)
        if ref is None: ref = out
        err = max(float(np.max(np.abs(out[k] - ref[k]) / np.maximum(np.abs(ref[k]), 1e-3 * np.max(np.abs(ref[k]))))) for k in ("f", "q", "dFdx", "dQdx"))
        for t, mb in [(128, 2), (128, 3), (128, 4), (256, 1), (384, 1)]:
            eng.set_option("b4_threads", t); eng.set_option("b4_minblocks", mb)
            ts = []
            for it in range(12):
                flush.fill_(0.0)
                e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
                e0.record(stream)
                eng.update_state(b[0], b[9], b[10], b[7], b[8], ss)
                e1.record(stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(2 * n_inv, "spec", spec, (t, mb), "eval_ms %.4f  evals/s %.3e  dev_vs_generic %.2e" % (np.median(ts[3:]), 2 * n_inv / np.median(ts[3:]) * 1e3, err), flush=True)
    eng.close()
