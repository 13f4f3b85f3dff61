"""Multi-GPU Newton iteration on a ring-oscillator array (BASELINE config 4 shape): instance partition per
rank, local CUDA evaluation + assembly, shared-unknown reduction over NCCL, block-distributed KLU-pattern LU
(interior blocks per GPU, shared Schur system all-reduced).  Checks the distributed Newton update against the
single-GPU solve of the undistributed system and reports device-timed throughput (max over ranks).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_newton.py --rings 4950
"""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from xyce_b200 import partition as pt, workloads as wl
from xyce_b200.capi import Engine, SolverState


class TorchXP:
    """The few array calls partition.schur_solve needs, on CUDA float64 tensors."""
    def __init__(self, dev): self.dev = dev; self.linalg = torch.linalg
    def zeros(self, shape): return torch.zeros(shape, dtype=torch.float64, device=self.dev)
    def concatenate(self, a, axis=0): return torch.cat(list(a), dim=axis)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rings", type=int, default=400)
    ap.add_argument("--stages", type=int, default=101)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--graph-partition", type=int, default=0, help="1 = partition_instances (graph partition) instead of ring ranges")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    wg = wl.ring_oscillator_array(a.rings, a.stages)
    w = pt.partition_workload(wg, world, rank) if a.graph_partition else pt.partition_ring_array(wg, world, rank)
    eng = wl.build_engine(w, device=local)
    stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
    n, nnz, ni, ns = w["n_unknowns"], eng.nnz, w["n_interior"], w["n_shared"]
    f64 = dict(dtype=torch.float64, device=dev)
    x = torch.tensor(w["x"] + 0.01 * np.sin(2 * w["glob_of_local"]), **f64)     # perturbed start (function of the global index)
    vec = [torch.zeros(n, **f64) for _ in range(4)]
    mat = [torch.zeros(nnz, **f64) for _ in range(2)]
    J = torch.zeros(nnz, **f64)
    sto = [torch.zeros(w["n_store"], **f64) for _ in range(2)]
    sta = [torch.zeros(w["n_state"], **f64) for _ in range(2)]
    ss = SolverState(transientFlag=1, newtonIter=1)
    h = 1e-12
    sysm = pt.BlockArrowSystem(w["rowptr"], w["colind"], ni)
    # second context: LU of the interior block A_ii
    sol = Engine(local); sol.set_stream(stream.cuda_stream)
    sol.set_pattern(sysm.ii_rowptr, sysm.ii_colind)
    idx = {k: torch.tensor(getattr(sysm, k), dtype=torch.long, device=dev) for k in
           ("ii_src", "is_src", "is_row", "is_col", "si_src", "si_row", "si_col", "ss_src", "ss_row", "ss_col")}
    class SysT: pass
    st = SysT(); st.ni, st.ns = ni, ns
    for k, v in idx.items(): setattr(st, k, v)
    xp = TorchXP(dev)
    vii = torch.zeros(len(sysm.ii_src), **f64)
    analyzed = [False]
    # linear part of this rank (C/h + G), added on the host side of the plumbing
    L = w["linear"]
    def lin_vals():
        v = np.zeros(nnz)
        for p, sc in (("g", 1.0), ("c", 1.0 / h)):
            r, c, val = L[p + "_row"], L[p + "_col"], L[p + "_val"]
            for rr, cc, vv in zip(r, c, val):
                k = w["rowptr"][rr] + np.searchsorted(w["colind"][w["rowptr"][rr]:w["rowptr"][rr + 1]], cc)
                v[k] += sc * vv
        return torch.tensor(v, **f64)
    Jlin = lin_vals()

    def solve_interior(B):
        if not analyzed[0]:
            assert sol.lu_analyze(vii.data_ptr()) == 0; analyzed[0] = True
        else:
            assert sol.lu_refactor(vii.data_ptr()) == 0
        Bt = B.t().contiguous(); Yt = torch.zeros_like(Bt)
        for k in range(Bt.shape[0]):
            sol.lu_solve(vii.data_ptr(), Bt[k].data_ptr(), Yt[k].data_ptr())
        return Yt.t()

    def allreduce(t):
        dist.all_reduce(t); return t

    def newton_iteration():
        eng.update_state(x.data_ptr(), sta[0].data_ptr(), sta[1].data_ptr(), sto[0].data_ptr(), sto[1].data_ptr(), ss)
        eng.load_vectors(*[t.data_ptr() for t in vec])
        eng.load_matrices(mat[0].data_ptr(), mat[1].data_ptr())
        eng.jacobian_combine(1.0 / h, mat[1].data_ptr(), 1.0, mat[0].data_ptr(), J.data_ptr())
        J.add_(Jlin)
        # shared-unknown reduction of the residual pieces (F, Q rows of the replicated unknowns)
        shared = torch.stack([v[ni:] for v in vec])
        dist.all_reduce(shared)
        for v, s in zip(vec, shared): v[ni:] = s / world       # keep partial-sum convention for the solve
        rhs = -(vec[0] + vec[1] / h)
        vii.copy_(J[st.ii_src])
        return pt.schur_solve(xp, st, J, rhs, solve_interior, allreduce), rhs

    dx, rhs = newton_iteration()
    torch.cuda.synchronize(); dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(a.iters):
        dx, rhs = newton_iteration()
    ev1.record(stream); torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1) / a.iters], **f64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    err = None
    if a.check:
        # undistributed reference on every rank's own GPU (small sizes): same loads, one LU over everything
        e1 = wl.build_engine(wg, device=local); e1.set_stream(stream.cuda_stream)
        xg = torch.tensor(wg["x"] + 0.01 * np.sin(2 * np.arange(wg["n_unknowns"])), **f64)
        g_vec = [torch.zeros(wg["n_unknowns"], **f64) for _ in range(4)]
        g_mat = [torch.zeros(e1.nnz, **f64) for _ in range(2)]
        gJ = torch.zeros(e1.nnz, **f64)
        gsto = [torch.zeros(wg["n_store"], **f64) for _ in range(2)]; gsta = [torch.zeros(wg["n_state"], **f64) for _ in range(2)]
        e1.update_state(xg.data_ptr(), gsta[0].data_ptr(), gsta[1].data_ptr(), gsto[0].data_ptr(), gsto[1].data_ptr(), ss)
        e1.load_vectors(*[t.data_ptr() for t in g_vec]); e1.load_matrices(g_mat[0].data_ptr(), g_mat[1].data_ptr())
        e1.jacobian_combine(1.0 / h, g_mat[1].data_ptr(), 1.0, g_mat[0].data_ptr(), gJ.data_ptr())
        Lg = wg["linear"]; v = np.zeros(e1.nnz)
        for p, sc in (("g", 1.0), ("c", 1.0 / h)):
            for rr, cc, vv in zip(Lg[p + "_row"], Lg[p + "_col"], Lg[p + "_val"]):
                k = wg["rowptr"][rr] + np.searchsorted(wg["colind"][wg["rowptr"][rr]:wg["rowptr"][rr + 1]], cc)
                v[k] += sc * vv
        gJ.add_(torch.tensor(v, **f64))
        grhs = -(g_vec[0] + g_vec[1] / h)
        gdx = torch.zeros_like(grhs)
        assert e1.lu_analyze(gJ.data_ptr()) == 0
        e1.lu_solve(gJ.data_ptr(), grhs.data_ptr(), gdx.data_ptr()); torch.cuda.synchronize()
        want = gdx[torch.tensor(w["glob_of_local"], dtype=torch.long, device=dev)]
        err = float((dx - want).abs().max() / want.abs().max())
        dxn = float(want.abs().max())
        e1.close()
    if rank == 0 and a.check:
        print("debug: max|dx| single-GPU = %.6e, distributed = %.6e, max abs diff = %.3e" % (dxn, float(dx.abs().max()), float((dx - want).abs().max())))
    errs = torch.tensor([err if err is not None else 0.0], **f64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "mosfets_total": wg["n_inst"], "unknowns_total": wg["n_unknowns"],
                          "shared_unknowns": ns, "ms_per_newton_iteration": float(ms), "newton_iters_per_s": 1e3 / float(ms),
                          "mosfet_evals_per_s": wg["n_inst"] / (float(ms) * 1e-3),
                          "max_rel_err_vs_single_gpu": float(errs) if a.check else None}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
