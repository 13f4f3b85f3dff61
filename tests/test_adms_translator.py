"""The ADMS translator (xyce_b200/adms/translate.py: admsXml `_nosac` C++ -> single-source evaluator): every
translated model against the reference's own generated class compiled into oracle/_ref.
CPU part: the generated evaluator compiled for the host (same statements, no FMA contraction) must reproduce the
reference object to the last bits; the translation of MVS 2.0.0 ETSOI must also agree with the hand restatement
(adms_mvs_eval.h).  GPU part (marked gpu): the same through xgpu_simple_group_add / the generic kernel."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_ref
from adms_common import ADMS_CARDS, BIAS, LIMITED, adms_circuit, bias_vector, outvars_close, outvars_mismatch
from b4_common import rel_err, solver_state
from dev_common import HostDevices, assemble

pytestmark = pytest.mark.skipif(not oracle_ref.available(), reason="oracle/_ref not built")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import xyce_b200
from xyce_b200.capi import Engine

MODELS = {m["name"]: m for m in Engine.adms_gen_models()}
PAIRS = [(m, c) for m in sorted(ADMS_CARDS) for c in sorted(ADMS_CARDS[m]) if m not in LIMITED]
LIMITED_PAIRS = [(m, c) for m in LIMITED for c in sorted(ADMS_CARDS[m])]
LIMIT_CASES = {"tran_iter1": dict(transient=1, newtonIter=1), "tran_iter0": dict(transient=1, newtonIter=0),
               "dcop_initjct": dict(dcop=1, tranop=1, initJct=1, newtonIter=0), "dcop_iter2": dict(dcop=1, tranop=1, newtonIter=2),
               "nolimit": dict(transient=1, newtonIter=2, voltageLimiter=0)}


def test_registry_lists_the_translated_models():
    assert set(ADMS_CARDS) <= set(MODELS), "build with the reference tree present translates %s" % sorted(ADMS_CARDS)
    e = MODELS["mvs_2_0_0_etsoi"]
    assert (e["nodes"], e["ext"], e["slots"]) == (7, 3, 17) and "M:type" in e["fields"]
    k = MODELS["ekv_va"]
    assert (k["nodes"], k["ext"], k["slots"]) == (4, 4, 16) and "I:L" in k["fields"] and "M:L" not in k["fields"]
    assert "I:admsTemperature" in k["fields"]
    p = MODELS["PSP103VA"]
    assert (p["nodes"], p["ext"]) == (13, 4) and len(p["fields"]) > 600
    assert MODELS["hicumL2va"]["nodes"] == 15 and MODELS["JUNCAP200"]["nodes"] == 2


def host_eval(hd, name, info, rec, V, gmin=1e-12):
    n, s = info["nodes"], info["slots"]
    out = np.zeros(2 * n + 2 * s + info["nstore"])
    dp = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
    k = hd.lib.xbh_adms_gen_eval(name.encode(), dp(rec), dp(np.asarray(V, dtype=np.float64)), C.c_double(gmin), out.ctypes.data_as(C.POINTER(C.c_double)))
    assert k == len(out), k
    z = np.zeros(n)
    return dict(F=out[:n], Q=out[n:2 * n], FL=z, QL=z, JF=out[2 * n:2 * n + s], JQ=out[2 * n + s:2 * n + 2 * s], store=out[2 * n + 2 * s:])


@pytest.mark.parametrize("model,card", PAIRS)
@pytest.mark.parametrize("case", ["tran1", "dcop2"])
def test_generated_evaluator_on_host_equals_reference_object(model, card, case):
    info = MODELS[model]
    hd = HostDevices()
    ref = adms_circuit(oracle_ref.RefCircuit, model, card, info["ext"], n_dev=8, seed=2)
    flags = dict(tran1=dict(transient=1, newtonIter=1), dcop2=dict(dcop=1, tranop=1, newtonIter=2))[case]
    ref.set_flags(**flags)
    rng = np.random.default_rng(11)
    exports = [ref.adms_export(i, model) for i in range(ref.n_inst)]
    x = bias_vector(model, ref.n, [e["lids"] for e in exports], rng)
    want = ref.load(x)
    assert not any(np.any(np.isnan(v)) for v in want.values())
    per, lids = [], []
    for i in range(ref.n_inst):
        e = exports[i]
        assert len(e["rec"]) == len(info["fields"]) and len(e["lids"]) == info["nodes"]
        V = [x[g] if g >= 0 else 0.0 for g in e["lids"]]
        per.append(host_eval(hd, model, info, e["rec"], V)); lids.append(e["lids"])
    # output variables: what Instance::updatePrimaryState copied to the store vector
    st = ref.get_state()["next_sto"]
    for i, o in enumerate(per):
        if info["nstore"]:
            w_ = st[exports[i]["sto0"]:exports[i]["sto0"] + info["nstore"]]
            assert outvars_close(o["store"], w_, 1e-12), (model, card, "store", i)
    asm = assemble(per, lids, info["slot_row"], info["slot_col"], ref.n, ref.rowptr, ref.colind)
    for k in ("f", "q", "dFdx", "dQdx"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(asm[k], want[k], scale) < 1e-13, (model, card, case, k)      # same statements: last-bit agreement
    assert np.any(want["dFdx"])
    if model in ("ekv_va", "hic0_full", "hicumL2va", "PSP103VA", "JUNCAP200", "bsim6", "bsimcmg_110"):
        assert np.any(want["q"]) and np.any(want["dQdx"])      # dynamic contributions are exercised


@pytest.mark.parametrize("card", sorted(ADMS_CARDS["mvs_2_0_0_etsoi"]))
def test_translated_mvs_agrees_with_the_hand_restatement(card):
    """adms_mvs_eval.h (dual numbers, written by hand in round 1) and the translator's output for the same model."""
    from dev_common import SIMPLE
    type_id, key, nodes, nstore, nstate, srow, scol = SIMPLE["mvs"]
    info = MODELS["mvs_2_0_0_etsoi"]
    hd = HostDevices()
    ref = adms_circuit(oracle_ref.RefCircuit, "mvs_2_0_0_etsoi", card, 3, n_dev=6, seed=4)
    rng = np.random.default_rng(5)
    x = rng.uniform(-0.6, 1.0, ref.n)
    for i in range(ref.n_inst):
        eg, eh = ref.adms_export(i, "mvs_2_0_0_etsoi"), ref.dev_export(i, "mvs")
        V = [x[g] if g >= 0 else 0.0 for g in eg["lids"]]
        a = host_eval(hd, "mvs_2_0_0_etsoi", info, eg["rec"], V)
        b = hd.simple(type_id, eh, dict(transient=1, newtonIter=1), V, [], [], [], nodes, len(srow), 0, 0)
        assert list(info["slot_row"]) == list(srow) and list(info["slot_col"]) == list(scol)
        for k in ("F", "JF"):
            sc = 1e-3 * np.max(np.abs(b[k]))
            assert rel_err(a[k], b[k], sc) < 1e-12, (card, i, k)


@pytest.mark.gpu
@pytest.mark.parametrize("model,card", PAIRS)
def test_gpu_generic_adms_kernel_matches_reference_object(model, card):
    info = MODELS[model]
    ref = adms_circuit(oracle_ref.RefCircuit, model, card, info["ext"], n_dev=200, seed=5)
    ex = [ref.adms_export(i, model) for i in range(ref.n_inst)]
    eng = xyce_b200.Engine(0)
    eng.set_pattern(ref.rowptr, ref.colind)
    eng.set_sizes(ref.n_sta, ref.n_sto)
    eng.add_simple_group(info["type"], np.array([e["rec"] for e in ex]), [0] * len(ex), np.array([e["lids"] for e in ex]),
                         [e["sto0"] for e in ex], 1, [e["sta0"] for e in ex], 1)
    eng.finalize()
    rng = np.random.default_rng(6)
    for case, flags in (("tran1", dict(transient=1, newtonIter=1)), ("dcop2", dict(dcop=1, tranop=1, newtonIter=2))):
        x = bias_vector(model, ref.n, [e["lids"] for e in ex], rng)
        ref.set_flags(**flags)
        want = ref.load(x)
        got = eng.load_host(x, solver_state(**flags))
        for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
            scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
            assert rel_err(got[k], want[k], scale) < 1e-12, (model, card, case, k)
        assert np.any(want["dFdx"] != 0.0) and eng.all_converged()
        if info["nstore"]:
            ws, gs = ref.get_state()["next_sto"], eng.get_state(0)
            assert outvars_mismatch(gs, ws, 1e-10, info["nstore"], illcond_share=0.05, illcond_tol=1e-3) is None, (model, card, case, "store")
    eng.close()


@pytest.mark.gpu
def test_gpu_ekv_inverter_dcop_and_tran_match_reference_flow():
    """CMOS inverter from two instances of the TRANSLATED EKV 2.6 model (static + dynamic contributions, temperature
    update), PULSE input, load capacitor: DC operating point and .TRAN on the GPU against the same driver around the
    reference's generated class and Kundert Sparse -- identical step sequence and Newton counts, waveforms within
    RELTOL / ABSTOL."""
    info = MODELS["ekv_va"]
    IN, VDD, OUT, BR_IN, BR_DD = range(5)
    ref = oracle_ref.RefCircuit(5)
    tn, pn, inn = ADMS_CARDS["ekv_va"]["nmos"]
    tp, pp, inp = ADMS_CARDS["ekv_va"]["pmos"]
    ref.add_dev_model("adms:ekv_va", "nmod", tn, 1, pn)
    ref.add_dev_model("adms:ekv_va", "pmod", tp, 1, pp)
    ref.add_dev_instance("adms:ekv_va", "M:n", "nmod", [OUT, IN, -1, -1], inn)
    ref.add_dev_instance("adms:ekv_va", "M:p", "pmod", [OUT, IN, VDD, VDD], inp)
    g, c = [], []
    for node, br in ((IN, BR_IN), (VDD, BR_DD)):
        g.append((node, br, 1.0)); g.append((br, node, 1.0))
    c.append((OUT, OUT, 20e-15))
    lin = dict(g_row=np.array([t[0] for t in g], dtype=np.int32), g_col=np.array([t[1] for t in g], dtype=np.int32),
               g_val=np.array([t[2] for t in g]), c_row=np.array([t[0] for t in c], dtype=np.int32),
               c_col=np.array([t[1] for t in c], dtype=np.int32), c_val=np.array([t[2] for t in c]))
    # PULSE(0 1.8 0.1n 0.1n 0.1n 0.5n 1.2n) on the input, DC 1.8 V supply
    src = dict(row=np.array([BR_IN, BR_DD], dtype=np.int32), scale=np.ones(2), type=np.array([1, 0], dtype=np.int32),
               params=np.array([[0.0, 1.8, 1e-10, 1e-10, 1e-10, 5e-10, 1.2e-9], [1.8, 0, 0, 0, 0, 0, 0]]))
    ref.add_pattern_entries(np.concatenate([lin["g_row"], lin["c_row"]]), np.concatenate([lin["g_col"], lin["c_col"]]))
    ref.finalize()
    x0 = np.zeros(ref.n); x0[VDD] = 1.8; x0[OUT] = 1.8
    probes = list(range(ref.n))
    ref.set_flags(transient=1)
    want = ref.tran_run(x0, 1.0e-9, 1e-11, probes, lin, src, dcop=1)
    ex = [ref.adms_export(i, "ekv_va") for i in range(2)]
    eng = xyce_b200.Engine(0)
    eng.set_pattern(ref.rowptr, ref.colind)
    eng.set_sizes(ref.n_sta, ref.n_sto)
    eng.add_simple_group(info["type"], np.array([e["rec"] for e in ex]), [0, 0], np.array([e["lids"] for e in ex]),
                         [e["sto0"] for e in ex], 1, [e["sta0"] for e in ex], 1)
    eng.set_linear(lin["g_row"], lin["g_col"], lin["g_val"], lin["c_row"], lin["c_col"], lin["c_val"])
    eng.set_sources(src["row"], src["scale"], src["type"], src["params"])
    eng.finalize()
    got = eng.tran_run(x0, 1.0e-9, 1e-11, probes, dcop=1)
    eng.close()
    assert want["rc"] == 0 and got["rc"] == 0, got.get("error")
    assert got["stats"]["dcop_newton_iters"] == want["stats"]["dcop_newton_iters"] >= 1
    assert got["stats"]["accepted"] == want["stats"]["accepted"] and got["stats"]["rejected"] == want["stats"]["rejected"]
    assert np.array_equal(got["steps"][:, 2], want["steps"][:, 2])
    tol = 1e-3 * np.maximum(np.abs(want["wave"]), np.abs(got["wave"])) + 1e-6
    assert np.all(np.abs(got["wave"] - want["wave"]) <= tol)
    assert want["wave"][0, OUT] > 1.7 and np.min(want["wave"][:, OUT]) < 0.1      # the inverter switches


def host_eval_limited(hd, name, info, rec, V, flags, cs, ns):
    from dev_common import flag_arrays
    n, s_, k_ = info["nodes"], info["slots"], info["nstore"]
    out = np.zeros(4 * n + 2 * s_ + k_ + 1)
    fl, fd = flag_arrays(flags)
    dp = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (rec, V, cs, ns)]
    k = hd.lib.xbh_adms_gen_eval2(name.encode(), dp(keep[0]), dp(keep[1]), fl.ctypes.data_as(C.POINTER(C.c_int)), dp(fd), dp(keep[2]), dp(keep[3]),
                                  out.ctypes.data_as(C.POINTER(C.c_double)))
    assert k == len(out), k
    o = 2 * n + 2 * s_ + k_
    return dict(F=out[:n], Q=out[n:2 * n], JF=out[2 * n:2 * n + s_], JQ=out[2 * n + s_:2 * n + 2 * s_], store=out[2 * n + 2 * s_:o],
                FL=out[o:o + n], QL=out[o + n:o + 2 * n], orig=int(out[o + 2 * n]))


@pytest.mark.parametrize("model,card", LIMITED_PAIRS)
@pytest.mark.parametrize("case", sorted(LIMIT_CASES))
def test_limited_models_on_host_equal_reference_object(model, card, case):
    """$limit: VBIC 1.3 and Mextram 504 -- limited junction voltages (pnjlim / pnjlim_new / the model's own limRTH), the
    dFdxdVp / dQdxdVp correction vectors, the limited probes saved to the store vector and Instance::isConverged (origFlag)
    against the reference's generated classes, with the previous-iterate voltages far enough from the new ones that the
    limiters act."""
    info = MODELS[model]
    hd = HostDevices()
    ref = adms_circuit(oracle_ref.RefCircuit, model, card, info["ext"], n_dev=8, seed=3)
    flags = LIMIT_CASES[case]
    ref.set_flags(**flags)
    rng = np.random.default_rng(12)
    exports = [ref.adms_export(i, model) for i in range(ref.n_inst)]
    x = bias_vector(model, ref.n, [e["lids"] for e in exports], rng)
    csto, nsto = rng.uniform(0.0, 0.3, ref.n_sto), rng.uniform(0.0, 0.3, ref.n_sto)
    ref.set_state(curr_sto=csto, next_sto=nsto)
    want = ref.load(x)
    assert not any(np.any(np.isnan(v)) for v in want.values())
    st = ref.get_state()["next_sto"]
    per, lids, limited = [], [], 0
    for i, e in enumerate(exports):
        V = [x[g] if g >= 0 else 0.0 for g in e["lids"]]
        k0 = e["sto0"]
        o = host_eval_limited(hd, model, info, e["rec"], V, flags, csto[k0:k0 + info["nstore"]], nsto[k0:k0 + info["nstore"]])
        per.append(o); lids.append(e["lids"])
        assert outvars_close(o["store"], st[k0:k0 + info["nstore"]], 1e-12), (model, card, case, "store", i)
        assert o["orig"] == ref.lib.xref_inst_converged(ref.h, i), (i, "isConverged")
        limited += 1 - o["orig"]
    asm = assemble(per, lids, info["slot_row"], info["slot_col"], ref.n, ref.rowptr, ref.colind)
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(asm[k], want[k], scale) < 1e-13, (model, card, case, k)
    if case in ("tran_iter1", "tran_iter0", "dcop_iter2"):
        assert limited > 0 and np.any(want["dFdxdVp"])      # the limiters did act
    if case == "nolimit":
        assert limited == 0 and not np.any(want["dFdxdVp"])


@pytest.mark.gpu
@pytest.mark.parametrize("model,card", LIMITED_PAIRS)
def test_gpu_limited_models_match_reference_object(model, card):
    info = MODELS[model]
    ref = adms_circuit(oracle_ref.RefCircuit, model, card, info["ext"], n_dev=150, seed=5)
    ex = [ref.adms_export(i, model) for i in range(ref.n_inst)]
    eng = xyce_b200.Engine(0)
    eng.set_pattern(ref.rowptr, ref.colind)
    eng.set_sizes(ref.n_sta, ref.n_sto)
    eng.add_simple_group(info["type"], np.array([e["rec"] for e in ex]), [0] * len(ex), np.array([e["lids"] for e in ex]),
                         [e["sto0"] for e in ex], 1, [e["sta0"] for e in ex], 1)
    eng.finalize()
    rng = np.random.default_rng(6)
    for case in sorted(LIMIT_CASES):
        flags = LIMIT_CASES[case]
        x = bias_vector(model, ref.n, [e["lids"] for e in ex], rng)
        csto, nsto = rng.uniform(0.0, 0.3, ref.n_sto), rng.uniform(0.0, 0.3, ref.n_sto)
        ref.set_flags(**flags); ref.set_state(curr_sto=csto, next_sto=nsto)
        eng.set_state(0, nsto); eng.set_state(1, csto)
        want = ref.load(x)
        got = eng.load_host(x, solver_state(**flags))
        for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
            scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
            assert rel_err(got[k], want[k], scale) < 1e-12, (model, card, case, k)
        assert outvars_mismatch(eng.get_state(0), ref.get_state()["next_sto"], 1e-10, info["nstore"], illcond_share=0.05, illcond_tol=1e-3) is None, (model, card, case)
        assert eng.all_converged() == all(ref.lib.xref_inst_converged(ref.h, i) for i in range(ref.n_inst))
    eng.close()


@pytest.mark.gpu
def test_gpu_vbic_amplifier_dcop_and_tran_with_limiting_match_reference_flow():
    """Common-emitter stage around the TRANSLATED VBIC 1.3 ($limit: junction limiting through the store vectors, initJct
    start, origFlag in the convergence test): DC operating point from zero, then .TRAN with a SIN input, on the GPU against
    the same driver around the reference's generated class and Kundert Sparse -- identical DCOP Newton count, step sequence
    and per-step Newton counts, waveforms within RELTOL / ABSTOL."""
    info = MODELS["vbic13"]
    IN, VCC, B, Cn, BR_IN, BR_CC = range(6)
    ref = oracle_ref.RefCircuit(6)
    mt, mp, ip = ADMS_CARDS["vbic13"]["npn"]
    ref.add_dev_model("adms:vbic13", "qmod", mt, 1, dict(mp, RCX=10.0, RBX=20.0, RE=1.0, CJE=2e-14, CJC=1e-14, TF=1e-11))
    ref.add_dev_instance("adms:vbic13", "Q:1", "qmod", [Cn, B, -1], ip)
    g, c = [], []
    def res(a, b, r):
        for (i, j, v) in ((a, a, 1 / r), (a, b, -1 / r), (b, a, -1 / r), (b, b, 1 / r)):
            if i >= 0 and j >= 0: g.append((i, j, v))
    for node, br in ((IN, BR_IN), (VCC, BR_CC)):
        g.append((node, br, 1.0)); g.append((br, node, 1.0))
    res(VCC, Cn, 5e2); res(IN, B, 2e4); res(VCC, B, 2e5)
    c.append((Cn, Cn, 1e-13))
    lin = dict(g_row=np.array([t[0] for t in g], dtype=np.int32), g_col=np.array([t[1] for t in g], dtype=np.int32),
               g_val=np.array([t[2] for t in g]), c_row=np.array([t[0] for t in c], dtype=np.int32),
               c_col=np.array([t[1] for t in c], dtype=np.int32), c_val=np.array([t[2] for t in c]))
    # SIN(0.9 0.2 1GHz) at the input, 3 V supply
    src = dict(row=np.array([BR_IN, BR_CC], dtype=np.int32), scale=np.ones(2), type=np.array([2, 0], dtype=np.int32),
               params=np.array([[0.9, 0.2, 1e9, 0, 0, 0, 0], [3.0, 0, 0, 0, 0, 0, 0]]))
    ref.add_pattern_entries(np.concatenate([lin["g_row"], lin["c_row"]]), np.concatenate([lin["g_col"], lin["c_col"]]))
    ref.finalize()
    x0 = np.zeros(ref.n)
    probes = list(range(ref.n))
    ref.set_flags(transient=1)
    want = ref.tran_run(x0, 2.0e-9, 2e-11, probes, lin, src, dcop=1)
    e = ref.adms_export(0, "vbic13")
    eng = xyce_b200.Engine(0)
    eng.set_pattern(ref.rowptr, ref.colind)
    eng.set_sizes(ref.n_sta, ref.n_sto)
    eng.add_simple_group(info["type"], np.array([e["rec"]]), [0], np.array([e["lids"]]), [e["sto0"]], 1, [e["sta0"]], 1)
    eng.set_linear(lin["g_row"], lin["g_col"], lin["g_val"], lin["c_row"], lin["c_col"], lin["c_val"])
    eng.set_sources(src["row"], src["scale"], src["type"], src["params"])
    eng.finalize()
    got = eng.tran_run(x0, 2.0e-9, 2e-11, probes, dcop=1)
    eng.close()
    assert want["rc"] == 0 and got["rc"] == 0, got.get("error")
    assert got["stats"]["dcop_newton_iters"] == want["stats"]["dcop_newton_iters"] >= 4      # from zero: the limiters walk the junctions up
    assert got["stats"]["accepted"] == want["stats"]["accepted"] and got["stats"]["rejected"] == want["stats"]["rejected"]
    assert np.array_equal(got["steps"][:, 2], want["steps"][:, 2])
    tol = 1e-3 * np.maximum(np.abs(want["wave"]), np.abs(got["wave"])) + 1e-6
    assert np.all(np.abs(got["wave"] - want["wave"]) <= tol)
    assert 0.2 < want["wave"][0, Cn] < 2.9 and np.ptp(want["wave"][:, Cn]) > 0.1      # biased in the active region, the collector follows the input
