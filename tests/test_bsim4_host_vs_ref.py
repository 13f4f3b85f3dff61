"""CPU-only: kernel source (host mirror) against the live reference objects (oracle/_ref), including the
name-by-name comparison of ~230 cached intermediates; skipped where oracle/_ref is not built."""
import numpy as np
import pytest

import oracle_ref
from b4_common import VARIANTS, assemble_general, isolated_devices, rel_err

pytestmark = pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")


@pytest.mark.parametrize("variant", ["default", "igc2", "capmod0", "capmod1", "rgate3", "rbody", "rdsmod", "pocket"])
def test_intermediates_and_assembly(host_mirror, variant):
    ref = isolated_devices(oracle_ref.RefCircuit, 4, variant, seed=21)
    rng = np.random.default_rng(5)
    flags = dict(transient=1, newtonIter=1)
    ref.set_flags(**flags)
    x = rng.uniform(-0.2, 1.2, ref.n)
    nsto = rng.normal(0, 0.3, ref.n_sto)
    von = rng.uniform(0.2, 0.6, ref.n_inst)
    ref.set_state(next_sto=nsto, curr_sto=nsto); ref.set_von(von)
    want = ref.load(x)
    per, lids = [], []
    for i in range(ref.n_inst):
        e = ref.export(i)
        V = np.array([x[g] if g >= 0 else 0.0 for g in e["lids"]])
        o = host_mirror.eval(e, flags, V, nsto[e["sto0"]:e["sto0"] + 13], True, von[i])
        per.append(o); lids.append(e["lids"])
        md, mi = ref.mid(i)
        for k, v in md.items():
            assert abs(o["mid_d"][k] - v) <= 1e-12 * max(abs(v), 1e-300), (variant, i, k)
        for k, v in mi.items():
            assert o["mid_i"][k] == v, (variant, i, k)
    asm = assemble_general(host_mirror, per, lids, ref.n, ref.rowptr, ref.colind)
    for k in want:
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(asm[k], want[k], scale) < 1e-12


def test_variants_really_change_the_topology():
    sizes = {}
    for v in ("default", "rgate", "rgate3", "rbody", "rdsmod"):
        ref = isolated_devices(oracle_ref.RefCircuit, 1, v, seed=1)
        sizes[v] = ref.n
    assert sizes["default"] == 8 and sizes["rgate"] > 8 and sizes["rgate3"] > sizes["rgate"]
    assert sizes["rbody"] == 8 + 2 * 3 and sizes["rdsmod"] == 8 + 2 * 2


@pytest.mark.parametrize("variant", ["default", "igc", "capmod0", "gidl"])
def test_lead_currents_of_a_four_terminal_device_are_its_own_row_contributions(host_mirror, variant):
    """Master::loadDAEVectors' lead-current block (N_DEV_MOSFET_B4.C:10933-10987) for a device without internal nodes:
    leadF / leadQ of (id, ig, is, ib) equal the instance's F / Q contributions to its drain, gate, source and bulk
    rows -- the identity xgpu_b4_lead_load relies on -- and junctionV = (Vd - Vs, Vg - Vs, 0, 0)."""
    ref = isolated_devices(oracle_ref.RefCircuit, 4, variant, seed=21, lead=True)
    rng = np.random.default_rng(5)
    flags = dict(transient=1, newtonIter=1)
    ref.set_flags(**flags)
    x = rng.uniform(-0.2, 1.2, ref.n)
    nsto = rng.normal(0, 0.3, ref.n_sto)
    von = rng.uniform(0.2, 0.6, ref.n_inst)
    ref.set_state(next_sto=nsto, curr_sto=nsto); ref.set_von(von)
    ref.load(x)
    lead = ref.lead()
    assert np.all(lead["branch0"] == 4 * np.arange(ref.n_inst))
    kD, kGE, kS, kB, kDP, kSP, kGP, kGM, kBP, kSB, kDB = range(11)
    for i in range(ref.n_inst):
        e = ref.export(i)
        V = np.array([x[g] if g >= 0 else 0.0 for g in e["lids"]])
        o = host_mirror.eval(e, flags, V, nsto[e["sto0"]:e["sto0"] + 13], True, von[i])
        b = lead["branch0"][i]
        rows = ((kD, kDP), (kGE, kGP, kGM), (kS, kSP), (kB, kBP, kSB, kDB))          # rows that collapse onto d, g, s, b
        for k, rr in enumerate(rows):
            for key, vec in (("F", "leadF"), ("Q", "leadQ")):
                got = sum(o[key][r] for r in rr)
                want = lead[vec][b + k]
                assert abs(got - want) <= 1e-12 * max(abs(want), 1e-3 * np.max(np.abs(lead[vec]))), (variant, i, k, key)
        assert lead["junctionV"][b] == V[kD] - V[kS] and lead["junctionV"][b + 1] == V[kGE] - V[kS]
        assert lead["junctionV"][b + 2] == 0.0 and lead["junctionV"][b + 3] == 0.0


@pytest.mark.parametrize("extra", [dict(TEMP=125.0), dict(TEMP=-40.0, NF=2.0), dict(M=3.0),
                                   dict(SA=2e-7, SB=3e-7, SD=1e-7, NF=2.0)], ids=["hot", "cold_nf2", "m3", "stress"])
@pytest.mark.parametrize("variant", ["default", "igc2", "gidl", "default_v470", "igc2_v461"])
def test_instance_parameters_reach_the_evaluator_through_the_records(host_mirror, variant, extra):
    """instance temperature, finger count, multiplicity and stress parameters act through
    processParams / updateTemperature of the reference (N_DEV_MOSFET_B4p82.C:120-2734 and the 4.7.0 / 4.6.1 twins),
    i.e. through the exported records: the evaluator must reproduce the reference on them unchanged"""
    from b4_common import host_mirror_case, records_from_ref
    ref = isolated_devices(oracle_ref.RefCircuit, 3, variant, seed=10, inst_extra=extra)
    rng = np.random.default_rng(1000)
    x = rng.uniform(-0.3, 1.3, ref.n)
    flags = dict(transient=1, newtonIter=1)
    ref.set_flags(**flags)
    nsto = rng.normal(0.3, 0.3, ref.n_sto)
    von = rng.uniform(0.2, 0.6, ref.n_inst)
    ref.set_state(curr_sto=nsto, next_sto=nsto, curr_sta=np.zeros(ref.n_sta)); ref.set_von(von)
    want = ref.load(x)
    g = dict(x=x, nsto=nsto, csto=nsto, von=von, rowptr=ref.rowptr, colind=ref.colind,
             flags=np.array([flags.get(f, 1 if f == "voltageLimiter" else 0) for f in oracle_ref.FLAG_NAMES]),
             curr_sta=np.zeros(ref.n_sta), n_sta=np.array(ref.n_sta), n_sto=np.array(ref.n_sto))
    for k, v in records_from_ref(ref).items():
        g["rec_" + k] = v
    _, asm = host_mirror_case(host_mirror, g)
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(asm[k], want[k], scale) < 1e-12, (variant, extra, k)
