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


def _factor_and_solve(A0, A1, b, **options):
    import torch
    A0 = sp.csr_matrix(A0); A0.sort_indices()
    A1 = sp.csr_matrix(A1); A1.sort_indices()
    eng = xyce_b200.Engine(0)
    for k, v in options.items():
        eng.set_option(k, v)
    eng.set_pattern(A0.indptr, A0.indices)
    dev = torch.device("cuda", 0)
    v0 = torch.tensor(A0.data, dtype=torch.float64, device=dev)
    v1 = torch.tensor(A1.data, dtype=torch.float64, device=dev)
    rhs = torch.tensor(b, dtype=torch.float64, device=dev)
    x = torch.zeros_like(rhs)
    assert eng.lu_analyze(v0.data_ptr()) == 0
    x0 = torch.zeros_like(rhs)
    eng.lu_solve(v0.data_ptr(), rhs.data_ptr(), x0.data_ptr())          # solve straight after the analysis (host factor values)
    rc = eng.lu_refactor(v1.data_ptr())
    eng.lu_solve(v1.data_ptr(), rhs.data_ptr(), x.data_ptr())
    eng.sync()
    f = eng.lu_export()
    out = (rc, x.cpu().numpy(), x0.cpu().numpy(), f)
    eng.close()
    return out


@pytest.mark.parametrize("n_rings,stages", [(16, 5), (333, 11), (300, 101), (4950, 101), (40, 300)])
def test_batched_groups_equal_the_warp_per_block_kernels_bitwise(n_rings, stages):
    """Equal-pattern blocks as one batched group (one thread per block, interleaved values, bundle programs) against
    the same plan on the warp-per-block kernels: same left-looking order per factor entry, so factors and solutions
    are bitwise identical -- and right (SuperLU)."""
    A0 = ring_array_matrix(n_rings, stages, seed=1)
    A1 = A0.copy()
    rng = np.random.default_rng(2)
    A1.data = A1.data * rng.uniform(0.8, 1.25, A1.nnz)
    xt = rng.normal(size=A0.shape[0])
    b = A1 @ xt
    rc_b, x_b, x0_b, f_b = _factor_and_solve(A0, A1, b, lu_batch=1)
    rc_w, x_w, x0_w, f_w = _factor_and_solve(A0, A1, b, lu_batch=0)
    assert rc_b == 0 and rc_w == 0
    assert np.array_equal(f_b["Lx"], f_w["Lx"]) and np.array_equal(f_b["Ux"], f_w["Ux"])
    assert np.array_equal(x_b, x_w) and np.array_equal(x0_b, x0_w)
    assert np.max(np.abs(x_b - xt)) / np.max(np.abs(xt)) < 1e-10
    x0s = spla.splu(sp.csc_matrix(A0)).solve(b)
    assert np.max(np.abs(x0_b - x0s)) / np.max(np.abs(x0s)) < 1e-10


@pytest.mark.parametrize("batch", [0, 1])
@pytest.mark.parametrize("n_rings,stages", [(40, 11), (3, 700)])
def test_pivot_monitor_reports_a_sequence_that_klu_would_no_longer_choose(n_rings, stages, batch):
    """Values drift so that one diagonal pivot falls below 0.001 x its column's largest candidate without becoming zero:
    a refactorization on the old pivot sequence (KLU_REPIVOT=0) would go on silently; the monitor returns code 3 so the
    caller re-analyses (what the reference's default KLU_REPIVOT=1 does on every factorization, N_LAS_AmesosSolver.C:316-318).
    Covers the batched, staged (40 x 11 with lu_batch = 0) and large-block (700 rows) refactor kernels."""
    import torch
    A0 = sp.csr_matrix(ring_array_matrix(n_rings, stages, seed=1)); A0.sort_indices()
    A1 = A0.copy()
    eng = xyce_b200.Engine(0)
    eng.set_option("lu_batch", batch)
    eng.set_pattern(A0.indptr, A0.indices)
    dev = torch.device("cuda", 0)
    v0 = torch.tensor(A0.data, dtype=torch.float64, device=dev)
    assert eng.lu_analyze(v0.data_ptr()) == 0
    assert eng.lu_refactor(v0.data_ptr()) == 0
    # the first pivot of the second ring's block is a matrix entry itself (no updates reach the first column of a block)
    # and the column has candidates below it (the block is irreducible): shrink that entry
    f = eng.lu_export()
    blk = [b for b in range(len(f["block_ptr"]) - 1) if f["block_ptr"][b + 1] - f["block_ptr"][b] == stages][1]
    t = f["block_ptr"][blk]
    assert f["Lp"][t + 1] > f["Lp"][t] and f["Up"][t + 1] - f["Up"][t] == 1
    r, c = int(f["row_perm"][t]), int(f["col_perm"][t])
    d = A1.indptr[r] + np.searchsorted(A1.indices[A1.indptr[r]:A1.indptr[r + 1]], c)
    assert A1.indices[d] == c
    A1.data[d] *= 1e-6
    v1 = torch.tensor(A1.data, dtype=torch.float64, device=dev)
    assert eng.lu_refactor(v1.data_ptr()) == 3
    eng.set_option("lu_pivot_check", 0)
    assert eng.lu_refactor(v1.data_ptr()) == 0               # klu_refactor semantics proper: no test
    eng.set_option("lu_pivot_check", 1)
    assert eng.lu_analyze(v1.data_ptr()) == 0                # re-pivot ...
    assert eng.lu_refactor(v1.data_ptr()) == 0               # ... and the new sequence passes
    rng = np.random.default_rng(3)
    xt = rng.normal(size=A0.shape[0]); b = torch.tensor(A1 @ xt, dtype=torch.float64, device=dev); x = torch.zeros_like(b)
    eng.lu_solve(v1.data_ptr(), b.data_ptr(), x.data_ptr()); eng.sync()
    assert np.max(np.abs(x.cpu().numpy() - xt)) / np.max(np.abs(xt)) < 1e-9
    eng.set_option("lu_repivot", 1)                          # KLU_REPIVOT=1: every refactorization pivots on the host
    assert eng.lu_refactor(v0.data_ptr()) == 0
    eng.close()
