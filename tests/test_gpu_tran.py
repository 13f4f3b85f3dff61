"""End-to-end .TRAN on the GPU (device eval + assembly + KLU-pattern LU + Newton/OneStep driver) against the
same control flow around the REFERENCE device code and Kundert Sparse (oracle/_ref):
Newton iteration counts per step identical, waveforms within Xyce's RELTOL/ABSTOL (1e-3 / 1e-6)."""
import numpy as np
import pytest

import oracle_ref
from b4_common import ref_circuit_from_workload
from xyce_b200 import workloads as wl

pytestmark = pytest.mark.gpu


def run_both(n_rings, stages, tstop, tstep=1e-12):
    w = wl.ring_oscillator_array(n_rings, stages)
    probes = [0, 1, stages // 2, w["vdd"], w["branch"]]
    ref = ref_circuit_from_workload(oracle_ref.RefCircuit, w)
    ref.set_flags(transient=1)
    want = ref.tran_run(w["x"], tstop, tstep, probes, w["linear"], w["sources"])
    eng = wl.build_engine(w)
    got = eng.tran_run(w["x"], tstop, tstep, probes)
    eng.close()
    return w, want, got


@pytest.mark.parametrize("n_rings,stages,tstop", [(1, 11, 2e-9), (1, 101, 2e-9), (3, 31, 1e-9)])
def test_ring_oscillator_tran_matches_reference_flow(n_rings, stages, tstop):
    w, want, got = run_both(n_rings, stages, tstop)
    assert want["rc"] == 0 and got["rc"] == 0, got.get("error")
    # identical step sequence and Newton iteration counts
    assert got["stats"]["accepted"] == want["stats"]["accepted"]
    assert got["stats"]["rejected"] == want["stats"]["rejected"]
    assert np.array_equal(got["steps"][:, 2], want["steps"][:, 2])          # Newton iterations per attempt
    # accept / reject pattern per attempt (the positive return codes 1 = "norm too small" and 2 = "normal
    # convergence" may swap when ||RHS||_2 sits at machine epsilon: the two-stage device reduction and the
    # sequential host sum differ in the last bits)
    assert np.array_equal(np.sign(got["steps"][:, 4]), np.sign(want["steps"][:, 4]))
    assert np.array_equal(got["steps"][:, 4] == -100, want["steps"][:, 4] == -100)
    assert np.allclose(got["t"], want["t"], rtol=1e-9, atol=0)
    # waveforms within RELTOL/ABSTOL
    tol = 1e-3 * np.maximum(np.abs(want["wave"]), np.abs(got["wave"])) + 1e-6
    assert np.all(np.abs(got["wave"] - want["wave"]) <= tol)
    # it really oscillates
    if stages <= 31:          # the switching front passes at least one of the probed ring nodes inside tstop
        swing = [want["wave"][:, p].max() - want["wave"][:, p].min() for p in range(3)]
        assert max(swing) > 0.8
    assert got["stats"]["lu_analyses"] >= 1 and got["stats"]["lu_refactors"] > 0


def test_mixed_device_netlist_tran_matches_reference_flow():
    """BASELINE config 5 shape: diode clipper + Gummel-Poon common-emitter stage + MOSFET level 1 inverter + R, C, V
    in one netlist; .TRAN on the GPU against the same driver around the reference's Diode / BJT / MOSFET1 objects
    and Kundert Sparse.  (The ADMS-shaped rlc plugin has no reference object in the tree; it is checked against its
    discrete equivalent and the analytic solution in test_gpu_devices.py.)"""
    import xyce_b200
    from dev_common import BJT_CARDS, DIODE_CARDS, MOS1_CARDS
    IN, A, VCC, B, C, E, D, BR_IN, BR_CC = range(9)
    ref = oracle_ref.RefCircuit(9)
    dp = dict(DIODE_CARDS["rs_bv"]); dp.pop("LEVEL", None)
    qt, qp = BJT_CARDS["basic"]; mt, mp = MOS1_CARDS["basic"]
    ref.add_dev_model("d", "dmod", "D", 1, dp)
    ref.add_dev_model("q", "qmod", qt, 1, dict(qp, RB=20.0, RC=5.0, RE=0.5))
    ref.add_dev_model("m1", "mmod", mt, 1, dict(mp, RD=10.0, RS=10.0))
    ref.add_dev_instance("d", "D:1", "dmod", [A, -1], dict(AREA=1.0))
    ref.add_dev_instance("d", "D:2", "dmod", [-1, A], dict(AREA=1.0))
    ref.add_dev_instance("q", "Q:1", "qmod", [C, B, E, -1], dict(AREA=1.0))
    ref.add_dev_instance("m1", "M:1", "mmod", [D, A, -1, -1], dict(L=2e-6, W=2e-5, AD=2e-11, AS=2e-11, PD=2e-5, PS=2e-5))
    g, c = [], []
    def res(a, b, r):
        gg = 1.0 / r
        for (i, j, v) in ((a, a, gg), (a, b, -gg), (b, a, -gg), (b, b, gg)):
            if i >= 0 and j >= 0: g.append((i, j, v))
    def cap(a, b, v):
        for (i, j, s) in ((a, a, v), (a, b, -v), (b, a, -v), (b, b, v)):
            if i >= 0 and j >= 0: c.append((i, j, s))
    def vsrc(node, br):
        g.append((node, br, 1.0)); g.append((br, node, 1.0))
    vsrc(IN, BR_IN); vsrc(VCC, BR_CC)
    res(IN, A, 1e3); cap(IN, B, 1e-10); res(VCC, B, 47e3); res(B, -1, 10e3); res(VCC, C, 2.2e3); res(E, -1, 470.0)
    res(VCC, D, 10e3); cap(D, -1, 1e-12); cap(C, -1, 2e-12)
    lin = dict(g_row=np.array([t[0] for t in g], dtype=np.int32), g_col=np.array([t[1] for t in g], dtype=np.int32),
               g_val=np.array([t[2] for t in g]), c_row=np.array([t[0] for t in c], dtype=np.int32),
               c_col=np.array([t[1] for t in c], dtype=np.int32), c_val=np.array([t[2] for t in c]))
    src = dict(row=np.array([BR_IN, BR_CC], dtype=np.int32), scale=np.ones(2), type=np.array([2, 0], dtype=np.int32),
               params=np.array([[0.0, 2.0, 1e6, 0, 0, 0, 0], [5.0, 0, 0, 0, 0, 0, 0]]))
    ref.add_pattern_entries(np.concatenate([lin["g_row"], lin["c_row"]]), np.concatenate([lin["g_col"], lin["c_col"]]))
    ref.finalize()
    x0 = np.zeros(ref.n); x0[VCC] = 5.0
    probes = [IN, A, B, C, D, BR_CC]
    ref.set_flags(transient=1)
    want = ref.tran_run(x0, 2e-6, 1e-9, probes, lin, src)
    eng = xyce_b200.Engine(0)
    eng.set_pattern(ref.rowptr, ref.colind)
    eng.set_sizes(ref.n_sta, ref.n_sto)
    ex = [ref.diode_export(i) for i in (0, 1)]
    eng.add_simple_group(1, np.array([e["rec"] for e in ex]), [e["flags"] for e in ex], np.array([e["lids"] for e in ex]),
                         [e["sto0"] for e in ex], 1, [e["sta0"] for e in ex], 1)
    for tid, key, idx in ((3, "q", 2), (2, "m1", 3)):
        e = ref.dev_export(idx, key)
        eng.add_simple_group(tid, np.array([e["rec"]]), [e["flags"]], np.array([e["lids"]]), [e["sto0"]], 1, [e["sta0"]], 1)
    eng.set_linear(lin["g_row"], lin["g_col"], lin["g_val"], lin["c_row"], lin["c_col"], lin["c_val"])
    eng.set_sources(src["row"], src["scale"], src["type"], src["params"])
    eng.finalize()
    got = eng.tran_run(x0, 2e-6, 1e-9, probes)
    eng.close()
    assert want["rc"] == 0 and got["rc"] == 0, got.get("error")
    # identical step sequence and Newton iteration counts
    assert got["stats"]["accepted"] == want["stats"]["accepted"] and got["stats"]["rejected"] == want["stats"]["rejected"]
    assert np.array_equal(got["steps"][:, 2], want["steps"][:, 2])
    assert np.allclose(got["t"], want["t"], rtol=1e-9, atol=0)
    tol = 1e-3 * np.maximum(np.abs(want["wave"]), np.abs(got["wave"])) + 1e-6
    assert np.all(np.abs(got["wave"] - want["wave"]) <= tol)
    # the stages do something: the clipper limits node A, the collector and the drain swing
    wv = want["wave"]
    assert np.max(np.abs(wv[:, 1])) < 1.2 and np.ptp(wv[:, 3]) > 0.5 and np.ptp(wv[:, 4]) > 1.0
