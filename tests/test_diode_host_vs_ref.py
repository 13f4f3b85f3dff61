"""CPU-only: diode evaluator source (host mirror) against the reference's own Diode objects (oracle/_ref)."""
import numpy as np
import pytest

import oracle_ref
from b4_common import rel_err
from dev_common import DIODE_CARDS, DIODE_SLOT_COL, DIODE_SLOT_ROW, HostDevices, assemble, diode_circuit

pytestmark = pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")

CASES = {"tran1": dict(transient=1, newtonIter=1), "tran0": dict(transient=1, newtonIter=0),
         "dcop_init": dict(dcop=1, tranop=1, initJct=1, newtonIter=0), "nolimit": dict(transient=1, newtonIter=2, voltageLimiter=0)}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("card", sorted(DIODE_CARDS))
def test_diode(card, case):
    hd = HostDevices()
    ref = diode_circuit(oracle_ref.RefCircuit, card, seed=3)
    rng = np.random.default_rng(4)
    flags = CASES[case]
    ref.set_flags(**flags)
    x = rng.uniform(-9.0, 1.2, ref.n)
    x[1::2] = rng.uniform(-0.3, 0.3, len(x[1::2]))
    nsto, csto = rng.normal(0.3, 0.4, ref.n_sto), rng.normal(0.3, 0.4, ref.n_sto)
    ref.set_state(curr_sto=csto, next_sto=nsto)
    want = ref.load(x)
    per, lids = [], []
    for i in range(ref.n_inst):
        e = ref.diode_export(i)
        V = [x[g] if g >= 0 else 0.0 for g in e["lids"]]
        o = hd.diode(e, flags, V, csto[e["sto0"]], nsto[e["sto0"]])
        per.append(o); lids.append(e["lids"])
        assert rel_err(o["store"], ref.get_state()["next_sto"][e["sto0"]:e["sto0"] + 3], 1e-30) < 1e-12
    asm = assemble(per, lids, DIODE_SLOT_ROW, DIODE_SLOT_COL, ref.n, ref.rowptr, ref.colind)
    for k in want:
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(asm[k], want[k], scale) < 1e-12, (card, case, k)
