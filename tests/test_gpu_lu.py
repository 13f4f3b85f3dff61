"""GPU sparse LU: refactor (fixed pattern + pivot order) and block-level triangular solves against
SciPy SuperLU and the reference tree's Kundert Sparse (oracle/_ref), tolerance 1e-10 relative
on the solution (LU is backward stable; orderings differ between the three solvers)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import xyce_b200
from test_lu_host import ksparse_solve, ring_array_matrix
import oracle_ref

pytestmark = pytest.mark.gpu


def gpu_factor_solve(A0, A1, b):
    """analyze on A0 (host pivoting), refactor with the values of A1 on the GPU, solve A1 x = b."""
    import torch
    A0 = sp.csr_matrix(A0); A0.sort_indices()
    A1 = sp.csr_matrix(A1); A1.sort_indices()
    assert np.array_equal(A0.indptr, A1.indptr) and np.array_equal(A0.indices, A1.indices)
    eng = xyce_b200.Engine(0)
    eng.set_pattern(A0.indptr, A0.indices)
    dev = torch.device("cuda", 0)
    v0 = torch.tensor(A0.data, dtype=torch.float64, device=dev)
    v1 = torch.tensor(A1.data, dtype=torch.float64, device=dev)
    rhs = torch.tensor(b, dtype=torch.float64, device=dev)
    x = torch.zeros_like(rhs)
    assert eng.lu_analyze(v0.data_ptr()) == 0
    assert eng.lu_refactor(v1.data_ptr()) == 0
    eng.lu_solve(v1.data_ptr(), rhs.data_ptr(), x.data_ptr())
    eng.sync()
    info = eng.lu_info()
    out = x.cpu().numpy()
    eng.close()
    return out, info


@pytest.mark.parametrize("n_rings,stages", [(1, 5), (9, 11), (300, 101), (40, 700)])
def test_ring_arrays(n_rings, stages):
    A0 = ring_array_matrix(n_rings, stages, seed=1)
    A1 = A0.copy()
    rng = np.random.default_rng(2)
    A1.data = A1.data * rng.uniform(0.8, 1.25, A1.nnz)       # new Newton iterate: same pattern, new values
    xt = rng.normal(size=A0.shape[0])
    b = A1 @ xt
    x, info = gpu_factor_solve(A0, A1, b)
    assert info["blocks"] == n_rings + 2 and info["largest_block"] == stages
    assert np.max(np.abs(x - xt)) / np.max(np.abs(xt)) < 1e-10
    xs = spla.splu(sp.csc_matrix(A1)).solve(b)
    assert np.max(np.abs(x - xs)) / np.max(np.abs(xs)) < 1e-10


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_single_block_vs_ksparse(seed):
    rng = np.random.default_rng(seed)
    n = 400
    A0 = sp.csr_matrix(sp.random(n, n, density=0.01, random_state=seed, format="csr") + sp.diags(rng.uniform(1, 2, n)))
    A0.sort_indices()
    A1 = A0.copy(); A1.data = A1.data * rng.uniform(0.9, 1.1, A1.nnz)
    b = rng.normal(size=n)
    x, _ = gpu_factor_solve(A0, A1, b)
    rk, xk = ksparse_solve(A1, b)
    assert rk == 0
    assert np.max(np.abs(x - xk)) / np.max(np.abs(xk)) < 1e-9
    assert np.max(np.abs(A1 @ x - b)) / np.max(np.abs(b)) < 1e-10


def test_singular_refactor_is_reported():
    import torch
    A = sp.csr_matrix(np.array([[2.0, 1.0], [1.0, 3.0]]))
    eng = xyce_b200.Engine(0)
    eng.set_pattern(A.indptr, A.indices)
    v = torch.tensor(A.data, dtype=torch.float64, device="cuda")
    assert eng.lu_analyze(v.data_ptr()) == 0
    z = torch.tensor([0.0, 1.0, 1.0, 0.0], dtype=torch.float64, device="cuda")   # zero pivots on the fixed order
    assert eng.lu_refactor(z.data_ptr()) == 2
    eng.close()


def coupled(A, seed=0):
    """make the ring array irreducible: a resistive supply (the branch row gets a diagonal entry), so that BTF
    finds ONE large block whose last column / row (the rail) is dense"""
    A = sp.lil_matrix(A)
    n = A.shape[0]
    A[n - 1, n - 1] = 0.7
    return sp.csr_matrix(A)


@pytest.mark.parametrize("n_rings,stages", [(12, 101), (60, 101), (3, 700)])
def test_large_single_block_column_level_schedule(n_rings, stages):
    """blocks of more than 512 rows take the column-level refactor and the row-level solves; with 60 x 101 the rail
    column has more than 4096 entries and is computed by the dense-column forward sweep"""
    A0 = coupled(ring_array_matrix(n_rings, stages, seed=3))
    A1 = A0.copy()
    rng = np.random.default_rng(4)
    A1.data = A1.data * rng.uniform(0.8, 1.25, A1.nnz)
    xt = rng.normal(size=A0.shape[0])
    b = A1 @ xt
    x, info = gpu_factor_solve(A0, A1, b)
    assert info["blocks"] == 1 and info["largest_block"] == A0.shape[0]
    assert np.max(np.abs(A1 @ x - b)) / np.max(np.abs(b)) < 1e-10
    xs = spla.splu(sp.csc_matrix(A1)).solve(b)
    assert np.max(np.abs(x - xs)) / np.max(np.abs(xs)) < 1e-9


def test_large_block_timing_smoke():
    """500 rings x 101 stages in one block (50 502 unknowns): refactor + solve must finish in milliseconds, not
    the seconds a single warp walking 50k columns would need"""
    import time, torch
    A0 = coupled(ring_array_matrix(500, 101, seed=5))
    A0.sort_indices()
    eng = xyce_b200.Engine(0)
    eng.set_pattern(A0.indptr, A0.indices)
    v = torch.tensor(A0.data, dtype=torch.float64, device="cuda")
    rng = np.random.default_rng(6)
    xt = rng.normal(size=A0.shape[0]); b = A0 @ xt
    rhs = torch.tensor(b, dtype=torch.float64, device="cuda"); x = torch.zeros_like(rhs)
    assert eng.lu_analyze(v.data_ptr()) == 0
    for _ in range(2):
        assert eng.lu_refactor(v.data_ptr()) == 0
        eng.lu_solve(v.data_ptr(), rhs.data_ptr(), x.data_ptr())
    eng.sync()
    t0 = time.perf_counter()
    for _ in range(5):
        assert eng.lu_refactor(v.data_ptr()) == 0
        eng.lu_solve(v.data_ptr(), rhs.data_ptr(), x.data_ptr())
    eng.sync()
    dt = (time.perf_counter() - t0) / 5
    print("large block refactor + solve: %.3f ms" % (1e3 * dt))
    assert np.max(np.abs(x.cpu().numpy() - xt)) / np.max(np.abs(xt)) < 1e-9
    assert dt < 0.05
    eng.close()
