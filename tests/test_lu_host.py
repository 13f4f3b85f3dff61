"""Host side of the sparse LU (BTF + ordering + pivoting Gilbert-Peierls): solutions against the reference
tree's own Kundert Sparse 1.3 (oracle/_ref, compiled from /root/reference/.../ksparse) and against
SciPy's SuperLU; structure checks on ring-oscillator-array matrices."""
import ctypes as C
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import oracle_ref
import xyce_b200


def host_solve(A, b):
    lib = xyce_b200.load_library()
    A = sp.csr_matrix(A); A.sort_indices()
    rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    x, info = np.zeros(A.shape[0]), np.zeros(8)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    rc = lib.xgpu_lu_host_factor_solve(A.shape[0], ip(rp), ip(ci), dp(v), dp(np.ascontiguousarray(b)), dp(x), dp(info))
    return rc, x, info


def ksparse_solve(A, b):
    lib = C.CDLL(oracle_ref.REF_SO)
    A = sp.csr_matrix(A); A.sort_indices()
    rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
    x = np.zeros(A.shape[0])
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = lib.xref_ksparse_solve(A.shape[0], ip(rp), ip(ci), dp(v), dp(np.ascontiguousarray(b)), dp(x))
    return rc, x


def ring_array_matrix(n_rings, stages, seed=0):
    """Jacobian-like matrix of n_rings ring oscillators sharing one supply node pinned by a source branch."""
    rng = np.random.default_rng(seed)
    n = n_rings * stages + 2
    vdd, br = n - 2, n - 1
    rows, cols, vals = [], [], []
    def add(r, c, v):
        rows.append(r); cols.append(c); vals.append(v)
    for r in range(n_rings):
        for k in range(stages):
            out, inp = r * stages + k, r * stages + (k - 1) % stages
            add(out, out, 3 + rng.random()); add(out, inp, 1 + rng.random()); add(out, vdd, -rng.random())
            add(inp, out, 0.1 * rng.random()); add(inp, inp, 0.5)
            add(vdd, out, -rng.random()); add(vdd, inp, 0.01); add(inp, vdd, 0.02)
    add(vdd, vdd, 1.0); add(vdd, br, 1.0); add(br, vdd, 1.0)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))


@pytest.mark.parametrize("n_rings,stages", [(1, 5), (7, 11), (40, 101)])
def test_ring_array_btf_and_solution(n_rings, stages):
    A = ring_array_matrix(n_rings, stages)
    rng = np.random.default_rng(1)
    xt = rng.normal(size=A.shape[0])
    b = A @ xt
    rc, x, info = host_solve(A, b)
    assert rc == 0
    # the source branch row and the supply KCL row are singletons, every ring is one block
    assert info[1] == n_rings + 2 and info[2] == stages
    assert np.max(np.abs(x - xt)) / np.max(np.abs(xt)) < 1e-10
    xs = spla.splu(sp.csc_matrix(A)).solve(b)
    assert np.max(np.abs(x - xs)) / np.max(np.abs(xs)) < 1e-10


@pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_against_reference_ksparse(seed):
    rng = np.random.default_rng(seed)
    n = 60
    A = sp.random(n, n, density=0.08, random_state=seed, format="csr") + sp.diags(rng.uniform(1, 2, n))
    A = sp.csr_matrix(A)
    b = rng.normal(size=n)
    rc, x, _ = host_solve(A, b)
    rk, xk = ksparse_solve(A, b)
    assert rc == 0 and rk == 0
    assert np.max(np.abs(x - xk)) / np.max(np.abs(xk)) < 1e-9
    assert np.max(np.abs(A @ x - b)) / np.max(np.abs(b)) < 1e-10


def test_unsymmetric_needs_pivoting():
    # zero diagonal entries force the transversal + off-diagonal pivots
    A = sp.csr_matrix(np.array([[0, 2.0, 0, 0], [1.0, 0, 0, 3.0], [0, 0, 0, 4.0], [0, 1.0, 5.0, 1e-9]]))
    b = np.array([1.0, 2.0, 3.0, 4.0])
    rc, x, _ = host_solve(A, b)
    assert rc == 0
    assert np.allclose(A @ x, b, rtol=1e-12, atol=1e-12)


def test_structurally_singular_is_reported():
    A = sp.csr_matrix(np.array([[1.0, 1.0, 0], [1.0, 1.0, 0], [0, 0, 0.0]]))
    A = sp.csr_matrix(([1.0, 1.0, 1.0, 1.0], ([0, 0, 1, 1], [0, 1, 0, 1])), shape=(3, 3))
    rc, _, _ = host_solve(A, np.ones(3))
    assert rc == 1


def batch_selfcheck(A0, A1):
    lib = xyce_b200.load_library()
    A0 = sp.csr_matrix(A0); A0.sort_indices()
    A1 = sp.csr_matrix(A1); A1.sort_indices()
    rp, ci = A0.indptr.astype(np.int32), A0.indices.astype(np.int32)
    out = np.zeros(4)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    rc = lib.xgpu_lu_host_batch_selfcheck(A0.shape[0], ip(rp), ip(ci), dp(A0.data.astype(np.float64)), dp(A1.data.astype(np.float64)), dp(out))
    return rc, out


@pytest.mark.parametrize("n_rings,stages", [(16, 5), (40, 11), (64, 101), (20, 300)])
def test_equal_pattern_blocks_are_batched_and_their_programs_are_right(n_rings, stages):
    """Every ring of a ring array has the same symbolic pattern: one batched group holds them all, and its bundle
    programs (refactor and solve, executed here on the host exactly as a GPU lane does) reproduce a plain
    left-looking refactorization and the triangular solves of each block bit for bit: the bundles only regroup
    independent operations, every factor entry still sees its updates in left-looking order."""
    A0 = ring_array_matrix(n_rings, stages, seed=1)
    A1 = A0.copy()
    A1.data = A1.data * np.random.default_rng(2).uniform(0.8, 1.25, A1.nnz)
    rc, out = batch_selfcheck(A0, A1)
    assert rc == 0
    assert out[0] == 1 and out[1] == n_rings, out
    assert out[2] == 0.0 and out[3] == 0.0, out


def test_few_or_unequal_blocks_are_not_batched():
    A0 = ring_array_matrix(6, 11, seed=1)          # fewer than kBatchMinBlocks equal blocks
    rc, out = batch_selfcheck(A0, A0)
    assert rc == 0 and out[0] == 0 and out[1] == 0
    rng = np.random.default_rng(0)                 # one irreducible random block
    n = 300
    A = sp.csr_matrix(sp.random(n, n, density=0.02, random_state=1, format="csr") + sp.diags(rng.uniform(1, 2, n)))
    rc, out = batch_selfcheck(A, A)
    assert rc == 0 and out[0] == 0
