"""BASELINE-size runs checked through size-independent properties and sampled oracle parity:
config 2 (100 000 BSIM4 instances, load + stamp) and config 3 (999 900 MOSFETs, Jacobian + KLU-pattern LU)."""
import numpy as np
import pytest
import torch

import oracle_ref
import xyce_b200
from b4_common import rel_err, solver_state
from xyce_b200 import workloads as wl

pytestmark = pytest.mark.gpu
FLAGS = dict(transient=1, newtonIter=1)


def load(eng, w, x=None):
    eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
    return eng.load_host(w["x"] if x is None else x, solver_state(**FLAGS))


def test_c2_sampled_oracle_parity_and_determinism():
    n_inv = 50000
    w = wl.inverter_array(n_inv, store_noise=0.3)          # limiters active on part of the array
    eng = wl.build_engine(w)
    a = load(eng, w)
    b = load(eng, w)
    for k in a:                                             # atomic-free assembly: bitwise reproducible
        assert np.array_equal(a[k], b[k]), k
    # oracle on a sample: inverters [s0, s0 + m) as their own array, same node voltages and limiter history
    m, s0 = 1500, 31000
    ws = wl.inverter_array(m)
    ref = oracle_ref.RefCircuit(2 * m + 1)
    ref.add_model("nch", "NMOS", wl.NMOS_CARD); ref.add_model("pch", "PMOS", wl.PMOS_CARD)
    vdd_s = 2 * m
    for j in range(m):
        ref.add_instance("M:n%d" % j, "nch", [2 * j + 1, 2 * j, -1, -1], wl.NMOS_INST)
    for j in range(m):
        ref.add_instance("M:p%d" % j, "pch", [2 * j + 1, 2 * j, vdd_s, vdd_s], wl.PMOS_INST)
    ref.finalize()
    xs = np.concatenate([w["x"][2 * s0:2 * (s0 + m)], [w["x"][2 * n_inv]]])
    sto_big = w["store"].reshape(22, 2 * n_inv)
    idx = np.concatenate([np.arange(s0, s0 + m), n_inv + np.arange(s0, s0 + m)])     # NMOS then PMOS of the sample
    sto_s = np.zeros(ref.n_sto)
    sto_s.reshape(2 * m, 22)[:, :] = sto_big[:, idx].T                                 # reference layout: instance-major
    ref.set_flags(**FLAGS)
    ref.set_state(curr_sto=sto_s, next_sto=sto_s); ref.set_von(w["von"][idx])
    want = ref.load(xs)
    rows = slice(2 * s0, 2 * (s0 + m))                                                  # in / out rows of the sample
    for k in ("f", "q", "dFdxdVp", "dQdxdVp"):
        scale = 1e-3 * np.max(np.abs(want[k][:2 * m])) if np.any(want[k][:2 * m]) else 1e-300
        assert rel_err(a[k][rows], want[k][:2 * m], scale) < 1e-12, k
    # matrix rows of the sample: same (row, col) entries with the supply column remapped
    rp, ci = w["rowptr"], w["colind"]
    for k in ("dFdx", "dQdx"):
        got = {}
        for r in range(2 * s0, 2 * (s0 + m)):
            for p in range(rp[r], rp[r + 1]):
                c = ci[p]
                got[(r - 2 * s0, c - 2 * s0 if c < 2 * n_inv else vdd_s)] = a[k][p]
        ref_vals = want[k]
        scale = 1e-3 * np.max(np.abs(ref_vals))
        cnt = 0
        for r in range(2 * m):
            for p in range(ref.rowptr[r], ref.rowptr[r + 1]):
                g = got[(r, ref.colind[p])]
                assert abs(g - ref_vals[p]) <= 1e-12 * max(abs(ref_vals[p]), scale), (k, r)
                cnt += 1
        assert cnt == len(got)
    eng.close()


def test_c2_supply_rail_row_is_consistent_with_its_column():
    """Size-independent property at 100 000 instances: source and bulk of every PMOS sit on the supply node, and a
    MOSFET's currents do not change when all four terminals move together, so the column sums of its dF/dx stamp
    vanish.  Hence the supply-row diagonal -- one destination with 100 000 contributions, reduced by the chunked
    block tree -- must equal minus the sum of the supply-column entries of all other rows, which are assembled by
    the short-destination path.  Checks the long-destination reduction and 32-bit indexing at full size."""
    n_inv = 50000
    w = wl.inverter_array(n_inv)
    eng = wl.build_engine(w)
    a = load(eng, w)
    rp, ci = w["rowptr"], w["colind"]
    vdd = 2 * n_inv
    assert np.all(np.isfinite(a["dFdx"])) and np.all(np.isfinite(a["dQdx"]))
    for k in ("dFdx", "dQdx"):
        diag = a[k][rp[vdd] + np.searchsorted(ci[rp[vdd]:rp[vdd + 1]], vdd)]
        col = 0.0
        mag = abs(diag)
        for r in range(2 * n_inv):
            seg = ci[rp[r]:rp[r + 1]]
            p = np.searchsorted(seg, vdd)
            if p < len(seg) and seg[p] == vdd:
                col += a[k][rp[r] + p]; mag += abs(a[k][rp[r] + p])
        assert abs(diag + col) <= 1e-10 * mag, k
    eng.close()


def test_c3_jacobian_lu_residual_at_full_size():
    w = wl.ring_oscillator_array(4950, 101)
    eng = wl.build_engine(w)
    n, nnz = w["n_unknowns"], eng.nnz
    dev = dict(dtype=torch.float64, device="cuda")
    ss = solver_state(**FLAGS)
    x = torch.tensor(w["x"], **dev)
    sta = [torch.zeros(w["n_state"], **dev) for _ in range(2)]; sto = [torch.zeros(w["n_store"], **dev) for _ in range(2)]
    out = [torch.zeros(n, **dev) for _ in range(4)] + [torch.zeros(nnz, **dev) for _ in range(2)]
    eng.load_dae(x.data_ptr(), sta[0].data_ptr(), sta[1].data_ptr(), sto[0].data_ptr(), sto[1].data_ptr(), ss,
                 *[t.data_ptr() for t in out], accumulate=False)
    J = torch.zeros(nnz, **dev)
    h = 1e-12
    eng.jacobian_combine(1.0 / h, out[5].data_ptr(), 1.0, out[4].data_ptr(), J.data_ptr())
    # linear part (load capacitors / h, source branch) added through the pattern
    L = w["linear"]; rp, ci = w["rowptr"], w["colind"]
    Jh = J.cpu().numpy()
    for pre, sc in (("g", 1.0), ("c", 1.0 / h)):
        r, c, v = L[pre + "_row"], L[pre + "_col"], L[pre + "_val"]
        pos = np.array([rp[a] + np.searchsorted(ci[rp[a]:rp[a + 1]], b) for a, b in zip(r, c)])
        np.add.at(Jh, pos, sc * v)
    J = torch.tensor(Jh, **dev)
    rng = np.random.default_rng(0)
    xt = rng.normal(size=n)
    import scipy.sparse as sp
    A = sp.csr_matrix((Jh, ci, rp), shape=(n, n))
    b = torch.tensor(A @ xt, **dev); sol = torch.zeros(n, **dev)
    assert eng.lu_analyze(J.data_ptr()) == 0
    info = eng.lu_info()
    assert info["blocks"] == 4950 + 2 and info["largest_block"] == 101      # every ring its own BTF block
    assert eng.lu_refactor(J.data_ptr()) == 0
    eng.lu_solve(J.data_ptr(), b.data_ptr(), sol.data_ptr()); eng.sync()
    xs = sol.cpu().numpy()
    assert np.max(np.abs(A @ xs - A @ xt)) <= 1e-10 * np.max(np.abs(A @ xt))
    assert np.max(np.abs(xs - xt)) <= 1e-8 * np.max(np.abs(xt))
    eng.close()


@pytest.mark.parametrize("n_inv", [1, 2, 63, 64, 65, 129])
def test_ragged_and_tiny_groups(n_inv):
    """group sizes around the block boundaries (and a single inverter): same results as the oracle"""
    w = wl.inverter_array(n_inv, store_noise=0.2)
    eng = wl.build_engine(w)
    a = load(eng, w)
    ref = oracle_ref.RefCircuit(2 * n_inv + 1)
    ref.add_model("nch", "NMOS", wl.NMOS_CARD); ref.add_model("pch", "PMOS", wl.PMOS_CARD)
    for j in range(n_inv):
        ref.add_instance("M:n%d" % j, "nch", [2 * j + 1, 2 * j, -1, -1], wl.NMOS_INST)
    for j in range(n_inv):
        ref.add_instance("M:p%d" % j, "pch", [2 * j + 1, 2 * j, 2 * n_inv, 2 * n_inv], wl.PMOS_INST)
    ref.finalize()
    sto = np.zeros(ref.n_sto); sto.reshape(2 * n_inv, 22)[:, :] = w["store"].reshape(22, 2 * n_inv).T
    ref.set_flags(**FLAGS); ref.set_state(curr_sto=sto, next_sto=sto); ref.set_von(w["von"])
    want = ref.load(w["x"])
    assert np.array_equal(ref.rowptr, w["rowptr"]) and np.array_equal(ref.colind, w["colind"])
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(a[k], want[k], scale) < 1e-12, k
    eng.close()


def test_accumulate_keeps_the_plus_equals_contract():
    w = wl.inverter_array(300)
    eng = wl.build_engine(w)
    n, nnz = w["n_unknowns"], eng.nnz
    dev = dict(dtype=torch.float64, device="cuda")
    ss = solver_state(**FLAGS)
    x = torch.tensor(w["x"], **dev)
    sta = [torch.zeros(w["n_state"], **dev) for _ in range(2)]; sto = [torch.tensor(w["store"], **dev) for _ in range(2)]
    base = [torch.full((n,), 3.0, **dev) for _ in range(4)] + [torch.full((nnz,), -2.0, **dev) for _ in range(2)]
    zero = [torch.zeros(n, **dev) for _ in range(4)] + [torch.zeros(nnz, **dev) for _ in range(2)]
    st = (x.data_ptr(), sta[0].data_ptr(), sta[1].data_ptr(), sto[0].data_ptr(), sto[1].data_ptr(), ss)
    eng.load_dae(*st, *[t.data_ptr() for t in zero], accumulate=False)
    for t in sto: t.copy_(torch.tensor(w["store"], **dev))
    eng.b4_set_von(0, w["von"])
    eng.load_dae(*st, *[t.data_ptr() for t in base], accumulate=True)
    eng.sync()
    for z, b, c in zip(zero, base, [3.0] * 4 + [-2.0] * 2):
        assert torch.allclose(b, z + c, rtol=0, atol=1e-12 * float(z.abs().max() + 1))
    eng.close()


# ---- BASELINE config 3 at its stated size against the oracle -------------------------------------------------------
C3_RINGS, C3_STAGES = 4950, 101
C3_SAMPLE = [0, 1, 617, 1234, 2475, 3333, 4500, 4949]      # rings probed across the array


def _single_ring_oracle(w, r):
    """Ring r of the array as its own circuit around the reference's BSIM4 objects (same rotation of the initial condition)."""
    w1 = wl.ring_oscillator_array(1, C3_STAGES, shifts=[int(w["shift"][r])])
    from b4_common import ref_circuit_from_workload
    ref = ref_circuit_from_workload(oracle_ref.RefCircuit, w1)
    return w1, ref


def test_c3_sampled_oracle_parity_of_stamps_at_full_size():
    """999 900 MOSFETs: residual rows and Jacobian entries of 8 rings spread over the array equal the reference objects'
    on the same ring built alone (1e-12; per-entry scale = the largest entry of the ring's own stamp)."""
    w = wl.ring_oscillator_array(C3_RINGS, C3_STAGES)
    rng = np.random.default_rng(5)
    x = w["x"].copy()
    x[:C3_RINGS * C3_STAGES] += rng.normal(0.0, 0.05, C3_RINGS * C3_STAGES)      # off the rails: every region of the model
    eng = wl.build_engine(w)
    got = eng.load_host(x, solver_state(**FLAGS))
    rp, ci = w["rowptr"], w["colind"]
    S = C3_STAGES
    for r in C3_SAMPLE:
        w1, ref = _single_ring_oracle(w, r)
        ref.set_flags(**FLAGS)
        x1 = np.concatenate([x[r * S:(r + 1) * S], [x[w["vdd"]], x[w["branch"]]]])
        want = ref.load(x1)
        rows = slice(r * S, (r + 1) * S)
        for k in ("f", "q", "dFdxdVp", "dQdxdVp"):
            scale = 1e-3 * np.max(np.abs(want[k][:S])) if np.any(want[k][:S]) else 1e-300
            assert rel_err(got[k][rows], want[k][:S], scale) < 1e-12, (k, r)
        for k in ("dFdx", "dQdx"):
            scale = 1e-3 * np.max(np.abs(want[k]))
            cnt = 0
            for lr in range(S):
                gr = r * S + lr
                have = {}
                for p in range(rp[gr], rp[gr + 1]):
                    c = ci[p]
                    have[c - r * S if c < C3_RINGS * S else S + (c - C3_RINGS * S)] = got[k][p]
                for p in range(ref.rowptr[lr], ref.rowptr[lr + 1]):
                    g = have[ref.colind[p]]
                    assert abs(g - want[k][p]) <= 1e-12 * max(abs(want[k][p]), scale), (k, r, lr)
                    cnt += 1
                assert len(have) == ref.rowptr[lr + 1] - ref.rowptr[lr]
            assert cnt > 3 * S
    eng.close()


def test_c3_tran_every_sampled_ring_follows_the_single_ring_oracle():
    """The 4 950-ring .TRAN (200 ps) against the oracle at size.  The rings interact only through the ideal supply, so
    ring r of the array must follow the reference flow (reference BSIM4 objects + Kundert Sparse under the same driver)
    of that ring alone WHEN BOTH TAKE THE SAME TIME STEPS: the array's accepted steps (size and order) are replayed on
    the oracle side (the array's own step selection depends on norms over all 499 952 unknowns, which a 103-unknown
    circuit cannot reproduce).  Checked: Newton iterations per step identical, all 101 node waveforms of 8 rings within
    RELTOL / ABSTOL at every time point."""
    w = wl.ring_oscillator_array(C3_RINGS, C3_STAGES)
    S = C3_STAGES
    probes = np.concatenate([r * S + np.arange(S) for r in C3_SAMPLE] + [[w["vdd"], w["branch"]]]).astype(np.int32)
    eng = wl.build_engine(w)
    got = eng.tran_run(w["x"], 2e-10, 1e-12, probes)
    eng.close()
    assert got["rc"] == 0, got.get("error")
    acc = got["steps"][got["steps"][:, 4] > 0]
    assert len(acc) == got["stats"]["accepted"] and len(acc) >= 30
    h, order, iters = acc[:, 1], acc[:, 3].astype(np.int32), acc[:, 2]
    worst = 0.0
    for j, r in enumerate(C3_SAMPLE):
        w1, ref = _single_ring_oracle(w, r)
        ref.set_flags(transient=1)
        want = ref.tran_run(w1["x"], 2e-10, 1e-12, np.arange(S), w1["linear"], w1["sources"], replay=(h, order))
        assert want["rc"] == 0
        assert len(want["t"]) == len(got["t"]) and np.allclose(want["t"], got["t"], rtol=1e-12, atol=0)
        assert np.array_equal(want["steps"][:, 2], iters), (r, want["steps"][:, 2], iters)      # Newton iterations per step
        gw = got["wave"][:, j * S:(j + 1) * S]
        tol = 1e-3 * np.maximum(np.abs(want["wave"]), np.abs(gw)) + 1e-6
        assert np.all(np.abs(gw - want["wave"]) <= tol), (r, float(np.max(np.abs(gw - want["wave"]))))
        worst = max(worst, float(np.max(np.abs(gw - want["wave"]))))
        assert np.ptp(want["wave"]) > 0.8          # the switching front moves through the ring
    print("C3 .TRAN: %d accepted steps, %d Newton iterations, worst |dv| vs single-ring oracle %.3e V"
          % (len(acc), int(iters.sum()), worst))


@pytest.mark.parametrize("zero_copy", [0, 1])
def test_load_host_jr_is_the_combination_of_the_six_arrays(zero_copy):
    """xgpu_load_host_jr (J = qs dQdx + fs dFdx, r = -(qs Q + fs F) + qs dQdxdVp + fs dFdxdVp) against the same
    combination of the xgpu_load_host outputs, element by element in the same operation order (bitwise); pinned,
    mapped buffers exercise the zero-copy stores."""
    import ctypes as C
    w = wl.inverter_array(3000, store_noise=0.3)
    eng = wl.build_engine(w)
    six = load(eng, w)
    qs, fs = 1.0 / 3e-12, 0.5
    eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
    eng.set_option("zero_copy_out", zero_copy)
    ss = solver_state(**FLAGS)
    hx = torch.tensor(w["x"], dtype=torch.float64).pin_memory()
    hr = torch.zeros(eng.n, dtype=torch.float64).pin_memory(); hj = torch.zeros(eng.nnz, dtype=torch.float64).pin_memory()
    p = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
    rc = eng.lib.xgpu_load_host_jr(eng.h, p(hx), C.byref(ss), C.c_double(qs), C.c_double(fs), p(hr), p(hj))
    assert rc == 0
    want_j = qs * six["dQdx"] + fs * six["dFdx"]
    want_r = -(qs * six["q"] + fs * six["f"]) + (qs * six["dQdxdVp"] + fs * six["dFdxdVp"])
    assert np.allclose(hj.numpy(), want_j, rtol=1e-15, atol=0) and np.any(want_j)
    assert np.allclose(hr.numpy(), want_r, rtol=0, atol=1e-15 * np.max(np.abs(qs * six["q"]))) and np.any(six["dQdxdVp"])
    eng.close()


@pytest.mark.parametrize("workload", ["inverters", "rings"])
def test_newton_step_host_solves_the_system_the_loads_return(workload):
    """xgpu_newton_step_host (x in, dx out, everything between on the device) against SciPy's solution of J dx = r with
    J and r as xgpu_load_host_jr returns them plus the linear-device stamps (rings: load capacitors + the supply branch)
    and the caller's history term; second call = refactorization on the first call's pivot sequence."""
    import scipy.sparse as sp, scipy.sparse.linalg as spl
    if workload == "inverters":
        w = wl.inverter_array(2000, store_noise=0.3)
    else:
        w = wl.ring_oscillator_array(40, 31)
    eng = wl.build_engine(w)
    n, nnz = eng.n, eng.nnz
    rowptr, colind = w["rowptr"], w["colind"]
    qs, fs = 1.0 / 2e-12, 1.0
    ss = solver_state(**FLAGS)
    rng = np.random.default_rng(5)
    hist = rng.normal(0, 1e-6, n)
    xs = [w["x"], w["x"] + rng.normal(0, 0.02, n)]
    for it, x in enumerate(xs):
        eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
        r, J = eng.load_host_jr(x, ss, qs, fs)
        A = sp.csr_matrix((J, colind, rowptr), shape=(n, n)).tolil()
        if "linear" in w:
            L = w["linear"]
            G = sp.coo_matrix((L["g_val"], (L["g_row"], L["g_col"])), shape=(n, n)).tocsr()
            Cm = sp.coo_matrix((L["c_val"], (L["c_row"], L["c_col"])), shape=(n, n)).tocsr()
            A = A + qs * Cm + fs * G
            r = r - (qs * (Cm @ x) + fs * (G @ x))
        r = r - hist
        want = spl.spsolve(sp.csc_matrix(A), r)
        eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
        rhs = np.zeros(n)
        got = eng.newton_step_host(x, ss, qs, fs, hist=hist, rhs=rhs)
        assert np.allclose(rhs, r, rtol=1e-12, atol=1e-12 * np.max(np.abs(r)))
        scale = np.max(np.abs(want))
        assert np.max(np.abs(got - want)) <= 1e-9 * scale, (workload, it, float(np.max(np.abs(got - want))), scale)
    eng.close()


@pytest.mark.parametrize("workload", ["inverters", "rings"])
@pytest.mark.parametrize("percent", [50, 57])
def test_pipelined_host_path_is_bitwise_the_one_pass_path(workload, percent):
    """xgpu_load_host_jr in two parts (first part of every (model, bin) run evaluated, its destinations assembled,
    combined and copied on a second stream while the second part is evaluated) against the one-pass path: J, r, store,
    state, carried limiter thresholds bit for bit."""
    w = wl.inverter_array(3000, store_noise=0.3) if workload == "inverters" else wl.ring_oscillator_array(40, 31)
    qs, fs = 1.0 / 3e-12, 0.5
    ss = solver_state(**FLAGS)
    rng = np.random.default_rng(9)
    x = w["x"] + rng.normal(0, 0.05, len(w["x"]))
    out = {}
    for pipe in (0, 1):
        import xyce_b200
        eng = wl.build_engine(w) if pipe == 0 else None
        if pipe == 1:
            # the share of the first part must be chosen before xgpu_finalize: build by hand like wl.build_engine does
            eng = wl.build_engine(w, options={"pipe_percent": percent})
        eng.set_option("pipeline_host", pipe)
        eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
        r, J = eng.load_host_jr(x, ss, qs, fs)
        out[pipe] = (r, J, eng.get_state(0), eng.get_state(2), eng.b4_get_von(0, w["n_inst"]), eng.pipe_info())
        eng.close()
    assert out[0][5][0] == out[1][5][0] == 1                 # both circuits are numbered device by device: they pipeline
    assert out[1][5][1] > 0.3 * len(x) and out[1][5][2] > 0.3 * len(out[0][1])
    for a, b in zip(out[0][:5], out[1][:5]):
        assert np.array_equal(a, b)
    assert np.any(out[0][1]) and np.any(out[0][0])
