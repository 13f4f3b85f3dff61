"""CPU-only: MOSFET level 1 and Gummel-Poon BJT evaluator sources (host build of the kernel headers) against
the reference's own device objects (oracle/_ref): assembled F, Q, limiter vectors, dFdx, dQdx, store and
state vectors at 1e-12."""
import numpy as np
import pytest

import oracle_ref
from b4_common import rel_err
from dev_common import BJT_CARDS, MOS1_CARDS, MVS_CARDS, SIMPLE, HostDevices, assemble, simple_circuit

pytestmark = pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")

CASES = {"tran1": dict(transient=1, newtonIter=1), "tran0": dict(transient=1, newtonIter=0),
         "dcop_init": dict(dcop=1, tranop=1, initJct=1, newtonIter=0), "dcop2": dict(dcop=1, tranop=1, newtonIter=2),
         "nolimit": dict(transient=1, newtonIter=2, voltageLimiter=0),
         # DCOP continuation (MOSFET homotopy): vds / gate drive scaled inside the drain-current code (N_DEV_MOSFET1.C:2978-3001)
         "homotopy": dict(dcop=1, tranop=1, newtonIter=1, artParameter=1, gainScale=0.35, nltermScale=0.6)}


def run(kind, card, case, seed=3):
    type_id, key, nodes, nstore, nstate, srow, scol = SIMPLE[kind]
    hd = HostDevices()
    ref = simple_circuit(oracle_ref.RefCircuit, kind, card, seed=seed)
    rng = np.random.default_rng(seed + 1)
    flags = CASES[case]
    ref.set_flags(**flags)
    x = rng.uniform(-1.5, 1.5, ref.n) if kind != "mvs" else rng.uniform(-0.6, 1.0, ref.n)
    nsto, csto = rng.normal(0.2, 0.5, ref.n_sto), rng.normal(0.2, 0.5, ref.n_sto)
    csta = rng.normal(0.0, 1e-14, ref.n_sta)
    ref.set_state(curr_sto=csto, next_sto=nsto, curr_sta=csta)
    want = ref.load(x)
    st = ref.get_state()
    per, lids = [], []
    for i in range(ref.n_inst):
        e = ref.dev_export(i, key)
        V = [x[g] if g >= 0 else 0.0 for g in e["lids"]]
        o = hd.simple(type_id, e, flags, V, csto[e["sto0"]:e["sto0"] + nstore], nsto[e["sto0"]:e["sto0"] + nstore],
                      csta[e["sta0"]:e["sta0"] + nstate], nodes, len(srow), nstore, nstate)
        per.append(o); lids.append(e["lids"])
        assert rel_err(o["store"], st["next_sto"][e["sto0"]:e["sto0"] + nstore], 1e-25) < 1e-12, (i, "store")
        assert rel_err(o["state"], st["next_sta"][e["sta0"]:e["sta0"] + nstate], 1e-25) < 1e-12, (i, "state")
        assert o["orig"] == ref.lib.xref_inst_converged(ref.h, i), (i, "isConverged")
    asm = assemble(per, lids, srow, scol, ref.n, ref.rowptr, ref.colind)
    for k in want:
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(asm[k], want[k], scale) < 1e-12, (kind, card, case, k)


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("card", sorted(MOS1_CARDS))
def test_mos1(card, case):
    run("mos1", card, case)


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("card", sorted(BJT_CARDS))
def test_bjt(card, case):
    run("bjt", card, case)


@pytest.mark.parametrize("case", ["tran1", "dcop2"])
@pytest.mark.parametrize("card", sorted(MVS_CARDS))
def test_adms_mvs(card, case):
    """ADMS-generated model (MVS 2.0.0 ETSOI, N_DEV_ADMSmvs_2_0_0_etsoi.C): 3 internal nodes + a branch equation; the
    reference object runs through the generic per-instance DeviceMaster loops (Core/N_DEV_DeviceMaster.h:666-823)"""
    run("mvs", card, case)


@pytest.mark.parametrize("begin", [1, 0])
@pytest.mark.parametrize("case", ["tran1", "nolimit", "dcop2"])
def test_bjt_excess_phase(case, begin):
    """Model PTF != 0: Weil's approximation of the excess phase (Instance::oldDAEExcessPhaseCalculation1 / 2,
    N_DEV_BJT.C:2706-2799).  In transient the collector current follows iBE / qB through the step history kept in the
    store entry CEXBC (current and last store); the first step out of a break point seeds both history values.  DC
    operating point: no effect."""
    type_id, key, nodes, nstore, nstate, srow, scol = SIMPLE["bjt"]
    hd = HostDevices()
    ref = simple_circuit(oracle_ref.RefCircuit, "bjt", "ptf", seed=5)
    rng = np.random.default_rng(11)
    flags = CASES[case]
    ref.set_flags(**flags)
    dt0, dt1 = 3e-11, 2e-11
    ref.set_step(dt0, dt1, begin)
    x = rng.uniform(-0.2, 0.9, ref.n)
    nsto, csto, lsto = (rng.normal(0.2, 0.5, ref.n_sto) for _ in range(3))
    csto[3::4] = rng.uniform(1e-4, 2e-3, len(csto[3::4])); lsto[3::4] = csto[3::4] * rng.uniform(0.7, 1.2, len(csto[3::4]))
    ref.set_state(curr_sto=csto, next_sto=nsto, curr_sta=rng.normal(0.0, 1e-14, ref.n_sta))
    ref.last_store(lsto)
    want = ref.load(x)
    st = ref.get_state(); last_after = ref.last_store()
    per, lids = [], []
    for i in range(ref.n_inst):
        e = ref.dev_export(i, key)
        assert e["rec"][25] != 0.0          # excessPhaseFac = PTF (rad) * TF
        s0 = e["sto0"]
        V = [x[g] if g >= 0 else 0.0 for g in e["lids"]]
        o, cex = hd.bjt_xp(e, flags, [dt0, dt1, begin], V, csto[s0:s0 + 3], nsto[s0:s0 + 3],
                           [0.0, 0.0] if begin else [csto[s0 + 3], lsto[s0 + 3]])
        per.append(o); lids.append(e["lids"])
        assert rel_err(o["store"], st["next_sto"][s0:s0 + 3], 1e-25) < 1e-12
        tran = not flags.get("dcop")
        assert int(cex[0]) == ((1 if tran else 0) | (2 if tran and begin else 0))
        if tran:
            assert abs(cex[1] - st["next_sto"][s0 + 3]) <= 1e-12 * abs(cex[1]), (i, "next CEXBC")
            if begin:
                assert abs(cex[2] - st["curr_sto"][s0 + 3]) <= 1e-12 * abs(cex[2]) and st["curr_sto"][s0 + 3] == last_after[s0 + 3]
            else:
                assert st["curr_sto"][s0 + 3] == csto[s0 + 3] and last_after[s0 + 3] == lsto[s0 + 3]
        else:
            assert st["next_sto"][s0 + 3] == nsto[s0 + 3]
    asm = assemble(per, lids, srow, scol, ref.n, ref.rowptr, ref.colind)
    for k in want:
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(asm[k], want[k], scale) < 1e-12, (case, begin, k)
    if not flags.get("dcop"):      # the history really enters: the same point without it gives another collector current
        ref.set_step(dt0, dt1, 1)
        assert np.max(np.abs(ref.load(x)["f"] - want["f"])) > 1e-6 or begin
