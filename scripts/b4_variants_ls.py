import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState
for n_inv in (50000, 100000, 250000, 500000):
    w = wl.inverter_array(n_inv, store_noise=0.0)
    eng = wl.build_engine(w)
    stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
    ss = SolverState(transientFlag=1, newtonIter=1)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    b = [eng.device_buffer(i) for i in range(11)]
    eng.load_host(w["x"], ss)      # puts the operating point into the context buffers (b[0] = x)
    for ls in (0, 1):
        for t, mb in [(64, 4), (64, 6), (96, 4), (128, 2), (128, 3), (128, 4), (256, 1), (384, 1), (512, 1)]:
            for k, v in (("arith", 2), ("threads", t), ("minblocks", mb), ("uniform", 1), ("lockstep", ls)):
                eng.set_option("b4_" + k, v)
            ts = []
            for it in range(12):
                flush.fill_(0.0)
                e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
                e0.record(stream)
                eng.update_state(b[0], b[9], b[10], b[7], b[8], ss)
                e1.record(stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(n_inv * 2, "lockstep", ls, (t, mb), "eval_ms %.4f  evals/s %.3e" % (np.median(ts[3:]), 2 * n_inv / np.median(ts[3:]) * 1e3), flush=True)
    eng.close()
