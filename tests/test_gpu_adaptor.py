"""The COMPILED Xyce-side adaptor (adaptor/N_DEV_GpuMaster_B4.h : MOSFET_B4::GpuMaster, a subclass of the reference's
BSIM4 Master created through the factory hook) against the stock Master, both driven through the same Device virtuals
-- updateState -> loadDAEVectors -> loadDAEMatrices -> isConverged (Core/N_DEV_Device.h:312-531) -- on the reference's
own objects, Linear::Matrix addressing and ExternData vectors, in one process.  The adaptor extracts the records from the
reference's Instance / Model / SizeDependParam objects itself (no Python in between) and reaches the GPU through the C ABI."""
import numpy as np
import pytest

import oracle_ref
from b4_common import FLAG_NAMES, VARIANTS, isolated_devices, rel_err

pytestmark = pytest.mark.gpu


class GpuRef(oracle_ref.RefCircuit):
    def __init__(self, n):
        super().__init__(n)
        self.use_gpu_master(True)


CASES = {"tran1": dict(transient=1, newtonIter=1), "tran_init": dict(transient=1, initTran=1, newtonIter=0),
         "dcop_jct": dict(dcop=1, tranop=1, initJct=1, newtonIter=0), "nolimit": dict(transient=1, newtonIter=2, voltageLimiter=0)}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("variant", ["default", "rgate3", "rbody", "rdsmod", "igc2_v470", "mob1_v461"])
def test_gpu_master_equals_stock_master_through_the_device_virtuals(variant, case):
    assert variant in VARIANTS
    stock = isolated_devices(oracle_ref.RefCircuit, 24, variant, seed=3)
    gpu = isolated_devices(GpuRef, 24, variant, seed=3)
    gpu.gpu_attach(0)
    assert gpu.n == stock.n and np.array_equal(gpu.rowptr, stock.rowptr) and np.array_equal(gpu.colind, stock.colind)
    rng = np.random.default_rng(9)
    x = rng.uniform(-0.3, 1.3, stock.n)
    csto, nsto = rng.normal(0.3, 0.3, stock.n_sto), rng.normal(0.3, 0.3, stock.n_sto)
    von = rng.uniform(0.2, 0.6, stock.n_inst)
    flags = CASES[case]
    out = []
    for c in (stock, gpu):
        c.set_flags(**flags)
        c.set_state(curr_sto=csto, next_sto=nsto)
        c.set_von(von)
    gpu_eng_von(gpu, von)
    want, got = stock.load(x), gpu.load(x)
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(got[k], want[k], scale) < 1e-12, k
    ws, gs = stock.get_state(), gpu.get_state()
    # store slots the reference never writes (vged, vgmd) keep their input on both sides
    assert rel_err(gs["next_sto"], ws["next_sto"], 1e-30) < 1e-12
    assert rel_err(gs["next_sta"], ws["next_sta"], 1e-30) < 1e-12
    if case == "tran_init":
        assert rel_err(gs["curr_sta"], ws["curr_sta"], 1e-30) < 1e-12
    assert gpu.all_converged() == stock.all_converged()


def gpu_eng_von(gpu, von):
    """the carried limiter threshold lives in the GPU context (Instance::von on the stock side)"""
    gpu.lib.xref_gpu_set_von(gpu.h, oracle_ref.dptr(np.ascontiguousarray(von, dtype=np.float64)))


class GpuRefAll(oracle_ref.RefCircuit):
    """every device type that has an adaptor goes to its GPU master (adaptor/N_DEV_GpuMaster_Simple.h)"""
    def __init__(self, n):
        super().__init__(n)
        self.use_gpu_master(2)


def _compare_through_virtuals(stock, gpu, x, flags, seed=9, check_store=True, adms=False, nstore=None):
    gpu.gpu_attach(0)
    assert gpu.n == stock.n and np.array_equal(gpu.rowptr, stock.rowptr) and np.array_equal(gpu.colind, stock.colind)
    rng = np.random.default_rng(seed)
    csto, nsto = rng.normal(0.2, 0.4, stock.n_sto), rng.normal(0.2, 0.4, stock.n_sto)
    csta = rng.normal(0.0, 1e-14, stock.n_sta)
    for c in (stock, gpu):
        c.set_flags(**flags)
        c.set_state(curr_sto=csto, next_sto=nsto, curr_sta=csta)
    want, got = stock.load(x), gpu.load(x)
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(got[k], want[k], scale) < 1e-12, k
    ws, gs = stock.get_state(), gpu.get_state()
    if stock.n_sto and check_store:
        from adms_common import outvars_close
        assert outvars_close(gs["next_sto"], ws["next_sto"], 1e-10 if adms else 1e-12, nstore, illcond_share=0.05 if adms else 0.0, illcond_tol=1e-3)
    if stock.n_sta:
        assert rel_err(gs["next_sta"], ws["next_sta"], 1e-25) < 1e-12
    assert gpu.all_converged() == stock.all_converged()
    assert np.any(want["dFdx"])


@pytest.mark.parametrize("case", ["tran1", "dcop_jct", "nolimit"])
@pytest.mark.parametrize("kind,card", [("mos1", "basic"), ("mos1", "pmos_rs"), ("bjt", "basic"), ("bjt", "res_pnp")])
def test_simple_gpu_masters_equal_the_stock_masters_through_the_device_virtuals(kind, card, case):
    """GpuSimpleMaster<MOSFET1::Master, ...> / <BJT::Master, ...> created in place of the stock Masters: records extracted
    from the reference's Instance / Model objects by the adaptor itself, GPU reached through the C ABI."""
    from dev_common import simple_circuit
    stock = simple_circuit(oracle_ref.RefCircuit, kind, card, n_dev=30, seed=4)
    gpu = simple_circuit(GpuRefAll, kind, card, n_dev=30, seed=4)
    x = np.random.default_rng(5).uniform(-1.5, 1.5, stock.n)
    _compare_through_virtuals(stock, gpu, x, CASES[case])


@pytest.mark.parametrize("begin", [1, 0])
def test_bjt_gpu_master_with_excess_phase(begin):
    """PTF != 0 behind GpuSimpleMaster<BJT::Master, ...>: the adaptor hands the last store vector over and, on the first step
    out of a break point, brings the seeded history (current and last store entry CEXBC) back -- N_DEV_BJT.C:2706-2799."""
    from dev_common import simple_circuit
    stock = simple_circuit(oracle_ref.RefCircuit, "bjt", "ptf", n_dev=30, seed=4)
    gpu = simple_circuit(GpuRefAll, "bjt", "ptf", n_dev=30, seed=4)
    gpu.gpu_attach(0)
    rng = np.random.default_rng(12)
    x = rng.uniform(-0.2, 0.9, stock.n)
    nsto, csto, lsto = (rng.normal(0.2, 0.4, stock.n_sto) for _ in range(3))
    csto[3::4] = rng.uniform(1e-4, 2e-3, len(csto[3::4])); lsto[3::4] = csto[3::4] * rng.uniform(0.7, 1.2, len(csto[3::4]))
    csta = rng.normal(0.0, 1e-14, stock.n_sta)
    for c in (stock, gpu):
        c.set_flags(**CASES["tran1"]); c.set_step(3e-11, 2e-11, begin)
        c.set_state(curr_sto=csto, next_sto=nsto, curr_sta=csta); c.last_store(lsto)
    want, got = stock.load(x), gpu.load(x)
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(got[k], want[k], scale) < 1e-12, k
    ws, gs = stock.get_state(), gpu.get_state()
    for k in ("next_sto", "curr_sto"):
        assert rel_err(gs[k], ws[k], 1e-30) < 1e-12, k
    assert rel_err(gpu.last_store(), stock.last_store(), 1e-30) < 1e-12
    assert np.array_equal(ws["curr_sto"][3::4], csto[3::4]) == (not begin)


@pytest.mark.parametrize("case", ["tran1", "dcop_jct"])
def test_diode_gpu_master(case):
    from dev_common import DIODE_CARDS, diode_circuit
    for card in sorted(DIODE_CARDS):
        stock = diode_circuit(oracle_ref.RefCircuit, card, n_dev=20, seed=2)
        gpu = diode_circuit(GpuRefAll, card, n_dev=20, seed=2)
        x = np.random.default_rng(3).uniform(-1.0, 1.0, stock.n)
        _compare_through_virtuals(stock, gpu, x, CASES[case])


@pytest.mark.parametrize("model,card", [("mvs_2_0_0_etsoi", "nmos"), ("mvs_2_0_0_hemt", "wide"), ("ekv_va", "pmos"), ("ekv_va", "short_hot"),
                                         ("hicumL2va", "res"), ("hic0_full", "default"), ("PSP103VA", "pmos_rg"), ("JUNCAP200", "sized"),
                                         ("bsim6", "pmos_rg"), ("bsimcmg_110", "nfin"), ("DIODE_CMC", "rs")])
def test_translated_adms_models_behind_the_generic_device_master(model, card):
    """the admsXml-generated models use DeviceMaster<Traits> itself (no Master subclass): GpuSimpleMaster<DeviceMaster<Traits>,
    generated filler> takes its place; the record comes from the translator's adms_fill_<model>()."""
    from adms_common import adms_circuit, bias_vector
    import xyce_b200
    info = {m["name"]: m for m in xyce_b200.capi.Engine.adms_gen_models()}[model]
    stock = adms_circuit(oracle_ref.RefCircuit, model, card, info["ext"], n_dev=25, seed=6)
    gpu = adms_circuit(GpuRefAll, model, card, info["ext"], n_dev=25, seed=6)
    lids = [stock.adms_export(i, model)["lids"] for i in range(stock.n_inst)]
    x = bias_vector(model, stock.n, lids, np.random.default_rng(7))
    # the store vector of these models holds their output variables (operating-point quantities for .PRINT such as
    # HICUM's rcx_t, GMi, CPIi), published by the generic kernel like Instance::updatePrimaryState does
    _compare_through_virtuals(stock, gpu, x, CASES["tran1"], adms=True, nstore=info["nstore"])


@pytest.mark.parametrize("kind,card", [("b4", "default"), ("b4", "rdsmod"), ("diode", None), ("mos1", "pmos_rs"), ("bjt", "res_pnp")])
def test_gpu_masters_forward_lead_currents(kind, card):
    """loadLeadCurrent (.PRINT I(...) / P(...)): leadF, leadQ, junctionV handed to Device::loadDAEVectors come back from the
    GPU masters as the stock Masters write them (branch-data LIDs; entries nobody writes stay as they were)."""
    from dev_common import DIODE_CARDS, diode_circuit, simple_circuit
    if kind == "b4":
        stock = isolated_devices(oracle_ref.RefCircuit, 20, card, seed=3, lead=True)
        gpu = isolated_devices(GpuRef, 20, card, seed=3, lead=True)
    elif kind == "diode":
        c0 = sorted(DIODE_CARDS)[0]
        stock = diode_circuit(oracle_ref.RefCircuit, c0, n_dev=30, seed=4, lead=True)
        gpu = diode_circuit(GpuRefAll, c0, n_dev=30, seed=4, lead=True)
    else:
        stock = simple_circuit(oracle_ref.RefCircuit, kind, card, n_dev=30, seed=4, lead=True)
        gpu = simple_circuit(GpuRefAll, kind, card, n_dev=30, seed=4, lead=True)
    gpu.gpu_attach(0)
    rng = np.random.default_rng(10)
    x = rng.uniform(-0.3, 1.2, stock.n)
    csto, nsto = rng.normal(0.3, 0.3, stock.n_sto), rng.normal(0.3, 0.3, stock.n_sto)
    csta = rng.normal(0.0, 1e-14, stock.n_sta)
    von = rng.uniform(0.2, 0.6, stock.n_inst)
    for c in (stock, gpu):
        c.set_flags(**CASES["tran1"])
        c.set_state(curr_sto=csto, next_sto=nsto, curr_sta=csta)
        if kind == "b4":
            c.set_von(von)
    if kind == "b4":
        gpu_eng_von(gpu, von)
    stock.load(x); gpu.load(x)
    want, got = stock.lead(), gpu.lead()
    assert np.array_equal(want["branch0"], got["branch0"]) and np.any(want["leadF"])
    for k in ("leadF", "leadQ", "junctionV"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(got[k], want[k], scale) < 1e-12, (kind, k)
