"""CPU-only checks of the Newton/OneStep control flow (tran_driver.h) running on the reference device code:
the 11-stage BSIM4 ring oscillates with a stable period and the step controller accepts/rejects sensibly."""
import numpy as np
import pytest

import oracle_ref
from b4_common import ref_circuit_from_workload
from xyce_b200 import workloads as wl

pytestmark = pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")


def test_ring_oscillator_period_is_stable():
    w = wl.ring_oscillator_array(1, 11)
    ref = ref_circuit_from_workload(oracle_ref.RefCircuit, w)
    ref.set_flags(transient=1)
    r = ref.tran_run(w["x"], 4e-9, 1e-12, [0], w["linear"], w["sources"])
    assert r["rc"] == 0
    t, v = r["t"], r["wave"][:, 0]
    up = [t[i] for i in range(1, len(t)) if v[i - 1] < 0.5 <= v[i]]
    assert len(up) >= 3
    periods = np.diff(up)
    assert np.all(np.abs(periods - periods[-1]) < 0.05 * periods[-1])
    s = r["stats"]
    assert s["accepted"] > 10 * s["rejected"] / 2 and s["newton_iters"] / s["attempts"] < 4
    # supply node pinned by the source, branch current small and negative (current flows out of the source)


def _period(r):
    t, v = r["t"], r["wave"][:, 0]
    up = [t[i] for i in range(1, len(t)) if v[i - 1] < 0.5 <= v[i]]
    return np.diff(up)


def test_gear_and_trapezoid_agree_on_the_ring_period():
    """Gear12 (BDF 1-2, N_TIA_Gear12.C) and OneStep (trapezoid) control flows on the same ring: same physics,
    so the oscillation period agrees within the integration tolerance; both orders get used."""
    out = {}
    for method in (7, 8):
        w = wl.ring_oscillator_array(1, 11)
        ref = ref_circuit_from_workload(oracle_ref.RefCircuit, w)
        ref.set_flags(transient=1)
        out[method] = ref.tran_run(w["x"], 4e-9, 1e-12, [0], w["linear"], w["sources"], method=method)
        assert out[method]["rc"] == 0
    pt, pg = _period(out[7]), _period(out[8])
    assert len(pg) >= 2 and abs(pg[-1] - pt[-1]) < 0.03 * pt[-1]
    assert set(out[8]["steps"][:, 3]) == {1.0, 2.0}
    assert not np.array_equal(out[7]["t"][:50], out[8]["t"][:50])          # it really is a different integrator


def test_dc_operating_point_of_the_mixed_netlist():
    """NoTimeIntegration + DampedNewton (DC_OP defaults) from a zero start: bias point of the common-emitter stage
    is consistent (KCL through the collector / emitter resistors), the MOSFET with its gate at 0 V is off, and the
    transient that follows starts from rest."""
    from dev_common import mixed_netlist
    ref, lin, src, x0, probes = mixed_netlist()
    ref.set_flags(transient=1)
    r = ref.tran_run(x0, 2e-7, 1e-9, list(range(9)), lin, src, dcop=1)
    assert r["rc"] == 0 and r["stats"]["dcop_status"] > 0 and 2 <= r["stats"]["dcop_newton_iters"] <= 30
    IN, A, VCC, B, C, E, D, BR_IN, BR_CC = range(9)
    x = r["wave"][0]
    ic, ie, ib = (x[VCC] - x[C]) / 2.2e3, x[E] / 470.0, (x[VCC] - x[B]) / 47e3 - x[B] / 10e3
    assert abs(x[VCC] - 5.0) < 1e-12 and 0.6 < x[B] - x[E] < 0.8
    assert abs(ic + ib - ie) < 1e-6 * ie
    assert abs(x[D] - 5.0) < 1e-6 and abs(x[A]) < 1e-9
    # supply current = everything drawn from VCC
    assert abs(-x[BR_CC] - (ic + (x[VCC] - x[B]) / 47e3 + (x[VCC] - x[D]) / 10e3)) < 1e-9
    # starts from rest: the first accepted point moves the bias nodes only by what the input couples in
    # (SIN source, 12.6 mV after 1 ns, through the 100 pF capacitor into the base)
    assert np.max(np.abs(r["wave"][1, [B, C, E, D]] - x[[B, C, E, D]])) < 1.3e-2


@pytest.mark.parametrize("method", [7, 8])
def test_replaying_a_run_on_its_own_accepted_steps_reproduces_it(method):
    """TranParams::replay_h / replay_order (verification mode used by the at-size tests): a run replayed on its own
    accepted steps -- rejected attempts skipped -- takes the same Newton iterations and lands on the same waveform."""
    w = wl.ring_oscillator_array(1, 11)
    ref = ref_circuit_from_workload(oracle_ref.RefCircuit, w)
    ref.set_flags(transient=1)
    a = ref.tran_run(w["x"], 1.5e-9, 1e-12, [0, 5], w["linear"], w["sources"], method=method)
    assert a["rc"] == 0 and a["stats"]["rejected"] > 0          # the interesting case: the original run rejected steps
    acc = a["steps"][a["steps"][:, 4] > 0]
    ref2 = ref_circuit_from_workload(oracle_ref.RefCircuit, w)
    ref2.set_flags(transient=1)
    b = ref2.tran_run(w["x"], 1.5e-9, 1e-12, [0, 5], w["linear"], w["sources"], method=method,
                      replay=(acc[:, 1], acc[:, 3].astype(np.int32)))
    assert b["rc"] == 0 and b["stats"]["rejected"] == 0
    assert np.allclose(b["t"], a["t"], rtol=1e-13, atol=0)
    assert np.array_equal(b["steps"][:, 2], acc[:, 2])
    assert np.allclose(b["wave"], a["wave"], rtol=1e-9, atol=1e-12)
