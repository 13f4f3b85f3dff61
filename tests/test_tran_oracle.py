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
