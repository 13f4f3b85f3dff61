"""CUDA-graph replay of the LU launch sequences (large diagonal blocks: hundreds of launches per refactor / solve):
same results as plain stream launches, new values in the same buffers are picked up, a new analysis drops the graphs."""
import numpy as np
import pytest
import scipy.sparse as sp

import xyce_b200
from test_gpu_lu import coupled, ring_array_matrix

pytestmark = pytest.mark.gpu


def _solve_sequence(A0, scales, graphs):
    import torch
    eng = xyce_b200.Engine(0)
    eng.set_option("lu_graphs", graphs)
    eng.set_pattern(A0.indptr, A0.indices)
    v = torch.tensor(A0.data, dtype=torch.float64, device="cuda")
    rng = np.random.default_rng(8)
    b = rng.normal(size=A0.shape[0])
    rhs = torch.tensor(b, dtype=torch.float64, device="cuda"); x = torch.zeros_like(rhs)
    assert eng.lu_analyze(v.data_ptr()) == 0
    outs, launches = [], []
    for sc in scales:
        v.copy_(torch.tensor(A0.data * sc, dtype=torch.float64, device="cuda"))       # new values, same buffer
        l0 = eng.launch_count()
        assert eng.lu_refactor(v.data_ptr()) == 0
        eng.lu_solve(v.data_ptr(), rhs.data_ptr(), x.data_ptr())
        eng.sync()
        launches.append(eng.launch_count() - l0)
        outs.append(x.cpu().numpy().copy())
    # re-analysis (new pivot sequence) must drop the captured graphs
    assert eng.lu_analyze(v.data_ptr()) == 0
    assert eng.lu_refactor(v.data_ptr()) == 0
    eng.lu_solve(v.data_ptr(), rhs.data_ptr(), x.data_ptr())
    eng.sync()
    outs.append(x.cpu().numpy().copy())
    eng.close()
    return outs, launches, b


@pytest.mark.parametrize("n_rings,stages", [(12, 101), (60, 101)])
def test_graph_replay_matches_plain_launches(n_rings, stages):
    A0 = coupled(ring_array_matrix(n_rings, stages, seed=3))
    A0.sort_indices()
    rng = np.random.default_rng(5)
    scales = [rng.uniform(0.8, 1.25, A0.nnz) for _ in range(4)]
    plain, l_plain, b = _solve_sequence(A0, scales, 0)
    graph, l_graph, _ = _solve_sequence(A0, scales, 1)
    assert l_plain == l_graph and l_plain[0] > 24          # a sequence long enough to be captured
    for xp, xg in zip(plain, graph):
        assert np.array_equal(xp, xg)                       # same kernels, same order: bitwise equal
    for sc, xg in zip(scales, graph):
        A1 = sp.csr_matrix((A0.data * sc, A0.indices, A0.indptr), shape=A0.shape)
        assert np.max(np.abs(A1 @ xg - b)) / np.max(np.abs(b)) < 1e-10
