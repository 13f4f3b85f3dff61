"""Bordered block-diagonal solve (xgpu_border_*) on one GPU against the plain KLU-pattern LU and SuperLU, and the
NCCL plumbing with a one-rank communicator.  The multi-rank path itself is exercised by tests/test_gpu_multi.py
(needs >= 2 GPUs) and scripts/multi_gpu_tran.py."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

import xyce_b200
from test_lu_host import ring_array_matrix

pytestmark = pytest.mark.gpu


def arrow_matrix(n_cells, seed=0):
    """Inverter-array shape: cell j = unknowns (2j, 2j + 1), every cell coupled to ONE supply unknown (last)."""
    rng = np.random.default_rng(seed)
    n = 2 * n_cells + 1
    vdd = n - 1
    j = np.arange(n_cells)
    rows = np.concatenate([2 * j, 2 * j, 2 * j + 1, 2 * j + 1, 2 * j + 1, np.full(n_cells, vdd), np.full(n_cells, vdd), [vdd]])
    cols = np.concatenate([2 * j, 2 * j + 1, 2 * j, 2 * j + 1, np.full(n_cells, vdd), 2 * j, 2 * j + 1, [vdd]])
    vals = np.concatenate([2 + rng.random(n_cells), 0.1 * rng.random(n_cells), -1 - rng.random(n_cells), 3 + rng.random(n_cells),
                           -rng.random(n_cells), 0.01 * rng.random(n_cells), -rng.random(n_cells), [0.5 * n_cells]])
    A = sp.csr_matrix((vals, (rows, cols)), shape=(n, n)); A.sort_indices()
    return A


def border_solve(A0, A1, b, n_border, with_comm=False):
    dev = torch.device("cuda", 0)
    eng = xyce_b200.Engine(0)
    eng.set_pattern(A0.indptr, A0.indices)
    eng.finalize()
    if with_comm:
        eng.comm_init(xyce_b200.Engine.comm_unique_id(), 0, 1)
    eng.border_set(n_border)
    v0 = torch.tensor(A0.data, dtype=torch.float64, device=dev); v1 = torch.tensor(A1.data, dtype=torch.float64, device=dev)
    rhs = torch.tensor(b, dtype=torch.float64, device=dev); x = torch.zeros_like(rhs)
    assert eng.border_analyze(v0.data_ptr()) == 0
    assert eng.border_solve(v1.data_ptr(), rhs.data_ptr(), x.data_ptr()) == 0
    eng.sync()
    info = eng.lu_info() if n_border < A0.shape[0] else {}
    out = x.cpu().numpy()
    eng.close()
    return out, info


@pytest.mark.parametrize("with_comm", [False, True])
def test_ring_array_with_supply_and_branch_as_border(with_comm):
    A0 = sp.csr_matrix(ring_array_matrix(60, 31, seed=1)); A0.sort_indices()
    A1 = A0.copy(); rng = np.random.default_rng(2)
    A1.data = A1.data * rng.uniform(0.8, 1.25, A1.nnz)
    xt = rng.normal(size=A0.shape[0]); b = A1 @ xt
    x, info = border_solve(A0, A1, b, 2, with_comm)
    assert info["n"] == A0.shape[0] - 2 and info["blocks"] == 60 and info["largest_block"] == 31      # only the rings are factored
    assert np.max(np.abs(x - xt)) / np.max(np.abs(xt)) < 1e-10
    xs = spla.splu(sp.csc_matrix(A1)).solve(b)
    assert np.max(np.abs(x - xs)) / np.max(np.abs(xs)) < 1e-10


@pytest.mark.parametrize("n_cells", [5, 3000, 50000])
def test_arrow_matrix_becomes_equal_two_by_two_blocks(n_cells):
    """The supply unknown as border: what is ONE strongly connected block of 2 n + 1 rows for the plain LU falls apart
    into n equal 2 x 2 blocks (a single batched group); the long supply row is reduced in fixed-shape chunks."""
    A0 = arrow_matrix(n_cells, seed=1)
    A1 = A0.copy(); rng = np.random.default_rng(2)
    A1.data = A1.data * rng.uniform(0.8, 1.25, A1.nnz)
    xt = rng.normal(size=A0.shape[0]); b = A1 @ xt
    x, info = border_solve(A0, A1, b, 1)
    assert info["blocks"] == n_cells and info["largest_block"] == 2
    assert np.max(np.abs(x - xt)) / np.max(np.abs(xt)) < 1e-9
    assert np.max(np.abs(A1 @ x - b)) / np.max(np.abs(b)) < 1e-11


def test_border_of_zero_unknowns_is_the_plain_solve_and_a_pure_border_system_works():
    A0 = sp.csr_matrix(ring_array_matrix(20, 11, seed=3)); A0.sort_indices()
    rng = np.random.default_rng(4)
    xt = rng.normal(size=A0.shape[0]); b = A0 @ xt
    x, _ = border_solve(A0, A0, b, 0)
    assert np.max(np.abs(x - xt)) / np.max(np.abs(xt)) < 1e-10
    D = sp.csr_matrix(rng.normal(size=(7, 7)) + 5 * np.eye(7)); D.sort_indices()      # every unknown in the border
    xt = rng.normal(size=7); b = D @ xt
    x, _ = border_solve(D, D, b, 7)
    assert np.max(np.abs(x - xt)) / np.max(np.abs(xt)) < 1e-12


def test_shared_reduce_with_one_rank_is_the_identity():
    A0 = arrow_matrix(100)
    eng = xyce_b200.Engine(0)
    eng.set_pattern(A0.indptr, A0.indices); eng.finalize()
    eng.comm_init(xyce_b200.Engine.comm_unique_id(), 0, 1)
    eng.border_set(1)
    assert eng.border_info() == (200, 1, 201)
    dev = torch.device("cuda", 0)
    v = [torch.arange(201, dtype=torch.float64, device=dev) * (k + 1) for k in range(4)]
    w = [t.clone() for t in v]
    eng.shared_reduce(*[t.data_ptr() for t in v]); eng.sync()
    for a, b in zip(v, w):
        assert torch.equal(a, b)
    eng.close()
