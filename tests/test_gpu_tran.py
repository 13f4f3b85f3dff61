"""End-to-end .TRAN on the GPU (device eval + assembly + KLU-pattern LU + Newton/OneStep driver) against the
same control flow around the REFERENCE device code and Kundert Sparse (oracle/_ref):
Newton iteration counts per step identical, waveforms within Xyce's RELTOL/ABSTOL (1e-3 / 1e-6)."""
import numpy as np
import pytest

import oracle_ref
from b4_common import ref_circuit_from_workload
from dev_common import mixed_engine, mixed_netlist
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
    ref, lin, src, x0, probes = mixed_netlist()
    ref.set_flags(transient=1)
    want = ref.tran_run(x0, 2e-6, 1e-9, probes, lin, src)
    eng = mixed_engine(ref, lin, src)
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


def _same_flow(got, want):
    assert want["rc"] == 0 and got["rc"] == 0, got.get("error")
    assert got["stats"]["accepted"] == want["stats"]["accepted"] and got["stats"]["rejected"] == want["stats"]["rejected"]
    assert np.array_equal(got["steps"][:, 2], want["steps"][:, 2])
    assert np.array_equal(got["steps"][:, 3], want["steps"][:, 3])          # integration order per attempt
    assert np.allclose(got["t"], want["t"], rtol=1e-9, atol=0)
    tol = 1e-3 * np.maximum(np.abs(want["wave"]), np.abs(got["wave"])) + 1e-6
    assert np.all(np.abs(got["wave"] - want["wave"]) <= tol)


def test_gear_ring_oscillator_matches_reference_flow():
    """.OPTIONS TIMEINT METHOD=GEAR (Gear12: BDF order 1-2) on the GPU against the same control flow around the
    reference BSIM4 objects + Kundert Sparse."""
    w = wl.ring_oscillator_array(2, 11)
    probes = [0, 1, 5, w["vdd"], w["branch"]]
    ref = ref_circuit_from_workload(oracle_ref.RefCircuit, w)
    ref.set_flags(transient=1)
    want = ref.tran_run(w["x"], 1.5e-9, 1e-12, probes, w["linear"], w["sources"], method=8)
    eng = wl.build_engine(w)
    got = eng.tran_run(w["x"], 1.5e-9, 1e-12, probes, method=8)
    eng.close()
    _same_flow(got, want)
    assert set(got["steps"][:, 3]) == {1.0, 2.0}
    assert want["wave"][:, 0].max() - want["wave"][:, 0].min() > 0.8


@pytest.mark.parametrize("method", [7, 8])
def test_dcop_then_tran_mixed_netlist_matches_reference_flow(method):
    """DC operating point (NoTimeIntegration, DampedNewton DC_OP defaults, initJct / initFix flags) followed by
    .TRAN, diode + BJT + MOSFET level 1 netlist: identical Newton iteration counts in the DCOP and in every step,
    DC solution at 1e-9, waveforms within RELTOL / ABSTOL."""
    ref, lin, src, x0, _ = mixed_netlist()
    probes = list(range(9))
    ref.set_flags(transient=1)
    want = ref.tran_run(x0, 1e-6, 1e-9, probes, lin, src, dcop=1, method=method)
    eng = mixed_engine(ref, lin, src)
    got = eng.tran_run(x0, 1e-6, 1e-9, probes, dcop=1, method=method)
    eng.close()
    _same_flow(got, want)
    assert got["stats"]["dcop_newton_iters"] == want["stats"]["dcop_newton_iters"] >= 2
    assert got["stats"]["dcop_status"] == want["stats"]["dcop_status"] > 0
    assert np.allclose(got["wave"][0], want["wave"][0], rtol=1e-9, atol=1e-12)
