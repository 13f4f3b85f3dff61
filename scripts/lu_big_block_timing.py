"""Refactor + solve of one large strongly connected block (ring array with a resistive supply), plain stream launches
against CUDA-graph replay.  usage: lu_big_block_timing.py [n_rings] [stages]"""
import sys, os, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
import xyce_b200
from test_gpu_lu import coupled, ring_array_matrix
nr = int(sys.argv[1]) if len(sys.argv) > 1 else 500
st = int(sys.argv[2]) if len(sys.argv) > 2 else 101
A0 = coupled(ring_array_matrix(nr, st, seed=5)); A0.sort_indices()
for graphs in (0, 1):
    eng = xyce_b200.Engine(0)
    eng.set_option("lu_graphs", graphs)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)      # a capturable (non-default) stream
    eng.set_pattern(A0.indptr, A0.indices)
    v = torch.tensor(A0.data, dtype=torch.float64, device="cuda")
    rng = np.random.default_rng(6)
    xt = rng.normal(size=A0.shape[0]); b = A0 @ xt
    rhs = torch.tensor(b, dtype=torch.float64, device="cuda"); x = torch.zeros_like(rhs)
    t0 = time.perf_counter(); assert eng.lu_analyze(v.data_ptr()) == 0; t_an = time.perf_counter() - t0
    for _ in range(3):
        assert eng.lu_refactor(v.data_ptr()) == 0
        eng.lu_solve(v.data_ptr(), rhs.data_ptr(), x.data_ptr())
    torch.cuda.synchronize()
    l0 = eng.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tr = ts = 0.0
    n = 10
    t0 = time.perf_counter()
    for _ in range(n):
        ev[0].record(stream); assert eng.lu_refactor(v.data_ptr()) == 0
        ev[1].record(stream); eng.lu_solve(v.data_ptr(), rhs.data_ptr(), x.data_ptr()); ev[2].record(stream)
        torch.cuda.synchronize()
        tr += ev[0].elapsed_time(ev[1]); ts += ev[1].elapsed_time(ev[2])
    wall = (time.perf_counter() - t0) / n
    err = float(np.max(np.abs(x.cpu().numpy() - xt)) / np.max(np.abs(xt)))
    print(json.dumps(dict(unknowns=A0.shape[0], nnz=int(A0.nnz), graphs=graphs, host_analysis_s=t_an, refactor_ms=tr / n, solve_ms=ts / n,
                          wall_ms=1e3 * wall, launches_per_pair=(eng.launch_count() - l0) / n, lu=eng.lu_info(), rel_err=err)), flush=True)
    eng.close()
