"""xgpu_lu_import: GPU refactor + solve on the symbolic result of an EXTERNAL sparse LU (what a KLU-enabled Xyce hands
over after klu_analyze / klu_factor: P, Q, block boundaries, L / U patterns) -- here SciPy's SuperLU plays that role --
and the round trip through xgpu_lu_export."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import xyce_b200
from test_gpu_lu import coupled, ring_array_matrix

pytestmark = pytest.mark.gpu


def _gpu_solve(eng, A, b):
    import torch
    v = torch.tensor(A.data, dtype=torch.float64, device="cuda")
    rhs = torch.tensor(b, dtype=torch.float64, device="cuda"); x = torch.zeros_like(rhs)
    assert eng.lu_refactor(v.data_ptr()) == 0
    eng.lu_solve(v.data_ptr(), rhs.data_ptr(), x.data_ptr())
    eng.sync()
    return x.cpu().numpy()


def _superlu_plan(A):
    """permutations and L / U patterns of SuperLU: L U = Pr A Pc with (Pr A)[perm_r[i], :] = A[i, :] and
    (A Pc)[:, perm_c[j]] = A[:, j]; L carries an explicit unit diagonal, U its pivot somewhere in the column"""
    lu = spla.splu(sp.csc_matrix(A), permc_spec="COLAMD", diag_pivot_thresh=0.001)
    n = A.shape[0]
    row_perm = np.argsort(lu.perm_r).astype(np.int32)          # position -> row of A
    col_perm = np.argsort(lu.perm_c).astype(np.int32)          # position -> column of A
    L, U = sp.csc_matrix(lu.L), sp.csc_matrix(lu.U)
    return lu, dict(row_perm=row_perm, col_perm=col_perm, block_ptr=np.array([0, n], dtype=np.int32),
                    Lp=L.indptr, Li=L.indices, Up=U.indptr, Ui=U.indices)


@pytest.mark.parametrize("n_rings,stages", [(3, 31), (4, 101), (12, 101)])      # 95, 406 (warp-per-block) and 1214 rows (large-block path)
def test_refactor_and_solve_on_an_imported_superlu_plan(n_rings, stages):
    A0 = coupled(ring_array_matrix(n_rings, stages, seed=3))
    A0.sort_indices()
    lu, plan = _superlu_plan(A0)
    eng = xyce_b200.Engine(0)
    eng.set_pattern(A0.indptr, A0.indices)
    eng.lu_import(**plan)
    info = eng.lu_info()
    assert info["blocks"] == 1 and info["nnz_L"] == lu.L.nnz - A0.shape[0] and info["nnz_U"] == lu.U.nnz
    rng = np.random.default_rng(4)
    b = rng.normal(size=A0.shape[0])
    x = _gpu_solve(eng, A0, b)
    assert np.max(np.abs(A0 @ x - b)) / np.max(np.abs(b)) < 1e-10
    assert np.max(np.abs(x - lu.solve(b))) / np.max(np.abs(x)) < 1e-9
    # the factor VALUES computed on the GPU on SuperLU's pattern and pivot order are SuperLU's own factor
    ex = eng.lu_export()
    Ug = sp.csc_matrix((ex["Ux"], ex["Ui"], ex["Up"]), shape=A0.shape)
    Lg = sp.csc_matrix((ex["Lx"], ex["Li"], ex["Lp"]), shape=A0.shape) + sp.identity(A0.shape[0], format="csc")
    assert abs(Ug - sp.csc_matrix(lu.U)).max() <= 1e-10 * abs(lu.U).max()
    assert abs(Lg - sp.csc_matrix(lu.L)).max() <= 1e-10 * abs(lu.L).max()
    # new values on the same plan (klu_refactor semantics)
    A1 = A0.copy(); A1.data = A1.data * rng.uniform(0.8, 1.25, A1.nnz)
    x1 = _gpu_solve(eng, A1, b)
    assert np.max(np.abs(A1 @ x1 - b)) / np.max(np.abs(b)) < 1e-10
    eng.close()


@pytest.mark.parametrize("n_rings,stages,couple", [(9, 11, False), (40, 101, False), (6, 101, True)])
def test_export_import_round_trip_is_bitwise(n_rings, stages, couple):
    """own analysis (BTF blocks, levels, off-diagonal entries) exported and imported into a fresh context"""
    A0 = ring_array_matrix(n_rings, stages, seed=1)
    if couple:
        A0 = coupled(A0)
    A0 = sp.csr_matrix(A0); A0.sort_indices()
    import torch
    e1 = xyce_b200.Engine(0)
    e1.set_pattern(A0.indptr, A0.indices)
    v = torch.tensor(A0.data, dtype=torch.float64, device="cuda")
    assert e1.lu_analyze(v.data_ptr()) == 0
    rng = np.random.default_rng(2)
    b = rng.normal(size=A0.shape[0])
    x1 = _gpu_solve(e1, A0, b)
    ex = e1.lu_export()
    e2 = xyce_b200.Engine(0)
    e2.set_pattern(A0.indptr, A0.indices)
    e2.lu_import(ex["row_perm"], ex["col_perm"], ex["block_ptr"], ex["Lp"], ex["Li"], ex["Up"], ex["Ui"])
    assert e2.lu_info() == e1.lu_info()
    x2 = _gpu_solve(e2, A0, b)
    assert np.array_equal(x1, x2)
    assert np.max(np.abs(A0 @ x2 - b)) / np.max(np.abs(b)) < 1e-10
    e1.close(); e2.close()


@pytest.mark.parametrize("n_rings,stages", [(4, 101), (12, 101)])
def test_imported_plan_with_row_scaling(n_rings, stages):
    """the external solver factored the row-scaled matrix diag(1 / Rs) A (KLU's default scale = 2: Rs = max |a_ij| of
    the row): the GPU refactorization of the UNSCALED values on that plan reproduces its factor, solves agree"""
    A0 = sp.csr_matrix(coupled(ring_array_matrix(n_rings, stages, seed=7))); A0.sort_indices()
    rng = np.random.default_rng(9)
    A0.data = A0.data * np.repeat(10.0 ** rng.uniform(-3, 3, A0.shape[0]), np.diff(A0.indptr))      # badly scaled rows
    Rs = np.asarray(abs(A0).max(axis=1).todense()).ravel()
    As = sp.diags(1.0 / Rs) @ A0
    lu, plan = _superlu_plan(sp.csr_matrix(As))
    eng = xyce_b200.Engine(0)
    eng.set_pattern(A0.indptr, A0.indices)
    eng.lu_import(row_scale=Rs[plan["row_perm"]], **plan)
    b = rng.normal(size=A0.shape[0])
    x = _gpu_solve(eng, A0, b)
    assert np.max(np.abs(As @ x - b / Rs)) / np.max(np.abs(b / Rs)) < 1e-10
    assert np.max(np.abs(x - lu.solve(b / Rs))) / np.max(np.abs(x)) < 1e-9
    ex = eng.lu_export()
    Ug = sp.csc_matrix((ex["Ux"], ex["Ui"], ex["Up"]), shape=A0.shape)
    assert abs(Ug - sp.csc_matrix(lu.U)).max() <= 1e-10 * abs(lu.U).max()
    eng.close()


def test_malformed_plans_are_rejected():
    A0 = sp.csr_matrix(coupled(ring_array_matrix(2, 11, seed=3))); A0.sort_indices()
    _, plan = _superlu_plan(A0)
    eng = xyce_b200.Engine(0)
    eng.set_pattern(A0.indptr, A0.indices)
    bad = dict(plan); bad["row_perm"] = plan["row_perm"].copy(); bad["row_perm"][0] = bad["row_perm"][1]
    with pytest.raises(RuntimeError, match="not a permutation"):
        eng.lu_import(**bad)
    bad = dict(plan); bad["Ui"] = plan["Ui"].copy(); bad["Up"] = plan["Up"].copy()
    # drop the last column's pivot
    bad["Up"][-1] -= 1
    with pytest.raises(RuntimeError, match="pivot|missing"):
        eng.lu_import(**bad)
    bad = dict(plan); bad["col_perm"] = np.roll(plan["col_perm"], 1)         # pattern no longer covers A
    with pytest.raises(RuntimeError):
        eng.lu_import(**bad)
    eng.lu_import(**plan)            # and a good plan still loads afterwards
    eng.close()
