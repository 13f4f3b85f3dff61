"""World-size-2 CPU (gloo) test of the multi-GPU decomposition logic in xyce_b200/partition.py: instance
partition, shared-unknown classification, and the block-distributed (Schur) solve with the shared system
all-reduced -- against the undistributed solution.  Interior blocks are factored by the library's host LU."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_lu_host import host_solve, ring_array_matrix
from xyce_b200 import partition as pt
from xyce_b200 import workloads as wl


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_rings, stages, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A = sp.csr_matrix(ring_array_matrix(n_rings, stages, seed=3)); A.sort_indices()
    n = A.shape[0]
    rng = np.random.default_rng(7)
    xt = rng.normal(size=n)
    b = A @ xt
    # ownership: ring r belongs to rank floor(r / rings_per_rank); vdd and the source branch are shared
    bnd = pt.split_ranges(n_rings, world)
    owner = np.full(n, -1)
    for r in range(world):
        owner[bnd[r] * stages:bnd[r + 1] * stages] = r
    glob, loc, ni = pt.local_numbering(owner, rank)
    # local system = rows/cols of this rank's unknowns; shared x shared entries and shared rhs split between ranks
    Al = sp.csr_matrix(A[glob][:, glob]); Al.sort_indices()
    vals = Al.data.copy()
    rows = np.repeat(np.arange(Al.shape[0]), np.diff(Al.indptr))
    ss = (rows >= ni) & (Al.indices >= ni)
    vals[ss] = vals[ss] / world                     # every rank contributes a share of the shared block
    rhs = b[glob].copy()
    rhs[ni:] /= world
    sysm = pt.BlockArrowSystem(Al.indptr, Al.indices, ni)
    Aii = sp.csr_matrix((vals[sysm.ii_src], sysm.ii_colind, sysm.ii_rowptr), shape=(ni, ni))

    def solve_interior(B):
        Y = np.zeros_like(B)
        for k in range(B.shape[1]):
            rc, x, _ = host_solve(Aii, np.ascontiguousarray(B[:, k]))
            assert rc == 0
            Y[:, k] = x
        return Y

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    x = pt.schur_solve(np, sysm, vals, rhs, solve_interior, allreduce)
    err = float(np.max(np.abs(x - xt[glob])) / np.max(np.abs(xt)))
    out_q.put((rank, err, ni, len(glob) - ni))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rings,stages", [(4, 7), (9, 31)])
def test_block_distributed_solve_world2(n_rings, stages):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rings, stages, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, ni, ns in res:
        assert err < 1e-10, (rank, err)
        assert ns == 2                                 # supply node + source branch are the shared unknowns
    assert sum(r[2] for r in res) == n_rings * stages


def test_ring_partition_is_consistent():
    w = wl.ring_oscillator_array(6, 11)
    parts = [pt.partition_ring_array(w, 3, r) for r in range(3)]
    assert sum(p["n_inst"] for p in parts) == w["n_inst"]
    assert sum(p["n_interior"] for p in parts) + parts[0]["n_shared"] == w["n_unknowns"]
    for p in parts:
        assert p["n_shared"] == 2 and p["lids"].max() < p["n_unknowns"]
        # interior unknowns of different ranks are disjoint
    allint = np.concatenate([p["glob_of_local"][:p["n_interior"]] for p in parts])
    assert len(np.unique(allint)) == len(allint)
    # the supply source's linear G stamps (shared unknowns only) are kept by rank 0 alone: the border reduction counts
    # them once; the source VALUE sits on a shared row and is replicated on every rank (B is not reduced)
    assert all(len(p["sources"]["row"]) == 1 for p in parts)
    assert len(parts[0]["linear"]["g_row"]) == 2 and all(len(p["linear"]["g_row"]) == 0 for p in parts[1:])


def test_linear_device_across_the_cut_makes_both_unknowns_shared():
    """A resistor between two rings owned by different ranks couples their interiors: both of its nodes must become
    shared (the classifier used to look at the BSIM4 nodes only and silently kept a foreign index)."""
    w = wl.ring_oscillator_array(4, 11)
    a, b = 3, 2 * 11 + 5                                   # node of ring 0 and node of ring 2
    L = w["linear"]
    for k, v in (("g_row", [a, a, b, b]), ("g_col", [a, b, a, b])):
        L[k] = np.concatenate([L[k], np.array(v, dtype=np.int32)])
    L["g_val"] = np.concatenate([L["g_val"], [1e-3, -1e-3, -1e-3, 1e-3]])
    parts = [pt.partition_ring_array(w, 2, r) for r in range(2)]
    for p in parts:
        assert p["n_shared"] == 4 and p["owner"][a] == -1 and p["owner"][b] == -1
        assert np.all(p["linear"]["g_row"] >= 0) and np.all(p["linear"]["g_col"] >= 0)
    # every stamp entry is kept exactly once
    assert sum(len(p["linear"]["g_row"]) for p in parts) == len(L["g_row"])


def test_graph_partition_keeps_rings_whole_and_balances():
    """partition_instances on the ring-oscillator array: the supply rail is recognised as a global net, every ring is a
    connected component and lands on one rank, loads are balanced, and only the rail (and the unknowns no instance
    touches) end up shared"""
    from xyce_b200 import workloads as wl
    w = wl.ring_oscillator_array(13, 11)
    inst_nodes = np.asarray(w["lids"])                                   # [n_inst, 12], -1 = ground
    for world in (2, 3, 4):
        owner = pt.partition_instances(inst_nodes, w["n_unknowns"], world)
        assert owner.min() == 0 and owner.max() == world - 1
        ring = inst_nodes[:, 0] // 11                                    # drain node = stage output, 11 stages per ring
        for r in range(13):
            assert len(set(owner[ring == r].tolist())) == 1              # a ring is never split
        counts = np.bincount(owner, minlength=world)
        assert counts.max() - counts.min() <= 2 * 11                     # within one ring (22 MOSFETs) of each other
        un = pt.classify_unknowns(inst_nodes, owner, w["n_unknowns"], world)
        shared = np.where(un == -1)[0]
        assert set(shared.tolist()) == {w["vdd"], w["branch"]}


def test_graph_partition_cuts_one_big_component_along_bfs_order():
    """a single chain of 2-terminal devices (one connected component heavier than any rank's share): graph growing cuts it
    into contiguous runs, so the number of shared unknowns is world - 1"""
    n = 1000
    inst_nodes = np.stack([np.arange(n), np.arange(1, n + 1)], axis=1)
    inst_nodes[-1, 1] = -1                                              # last device to ground
    for world in (2, 4, 8):
        owner = pt.partition_instances(inst_nodes, n, world)
        counts = np.bincount(owner, minlength=world)
        assert counts.max() <= 1.3 * n / world and counts.min() >= 0.7 * n / world
        un = pt.classify_unknowns(inst_nodes, owner, n, world)
        assert np.sum(un == -1) == world - 1
        assert np.sum(np.diff(owner) != 0) == world - 1                 # contiguous runs along the chain


def test_workload_partition_from_the_graph_partitioner():
    """partition_workload with the default (graph) partition on the ring array: same invariants as the ring-range split"""
    w = wl.ring_oscillator_array(7, 11)
    world = 3
    parts = [pt.partition_workload(w, world, r) for r in range(world)]
    assert sum(p["n_inst"] for p in parts) == w["n_inst"]
    assert sum(p["n_interior"] for p in parts) + parts[0]["n_shared"] == w["n_unknowns"]
    for p in parts:
        assert p["n_shared"] == 2 and p["lids"].max() < p["n_unknowns"] and p["n_inst"] % 22 == 0      # whole rings
    allint = np.concatenate([p["glob_of_local"][:p["n_interior"]] for p in parts])
    assert len(np.unique(allint)) == len(allint)
    assert all(len(p["sources"]["row"]) == 1 for p in parts)          # sources on shared rows are replicated
    # per-ring load capacitors stay with their ring's rank: C stamps of all parts together = the original ones
    assert sum(len(p["linear"]["c_row"]) for p in parts) == len(w["linear"]["c_row"])
