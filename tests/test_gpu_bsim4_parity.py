"""GPU parity: CUDA BSIM4 evaluation + assembly (through the C ABI, host buffers) against the
reference's own BSIM4 objects (oracle/_ref).  Tolerance: 1e-12 relative, scaled by the magnitude
of the entry or, for sums that cancel, by the largest entry of the same vector/matrix
(BASELINE.json north_star; SURVEY.md 8a accumulation-order note)."""
import numpy as np
import pytest

import oracle_ref
from b4_common import VARIANTS, engine_from_ref, isolated_devices, rel_err, solver_state

pytestmark = pytest.mark.gpu
TOL = 1e-12

FLAG_CASES = {
    "tran_iter1": dict(transient=1, newtonIter=1),
    "tran_iter0": dict(transient=1, newtonIter=0),
    "tran_init": dict(transient=1, newtonIter=0, initTran=1),
    "dcop_initjct": dict(dcop=1, tranop=1, transient=1, initJct=1, newtonIter=0),
    "dcop_iter2": dict(dcop=1, tranop=1, transient=1, newtonIter=2),
    "dc_nocharge": dict(dcop=1, newtonIter=1),
    "nolimit": dict(transient=1, newtonIter=1, voltageLimiter=0),
}


def run_case(variant, flags, n_pairs=16, seed=3, store_noise=0.3, arith=None, options=None, sorted_bins=False):
    assert oracle_ref.available(), "oracle/_ref/libxyce_ref.so must travel with the snapshot"
    ref = isolated_devices(oracle_ref.RefCircuit, n_pairs, variant, seed=seed, sorted_bins=sorted_bins)
    eng, rec = engine_from_ref(ref)
    if arith is not None:
        eng.set_option("b4_arith", arith)
    for k, v in (options or {}).items():
        eng.set_option(k, v)
    rng = np.random.default_rng(seed + 100)
    x = rng.uniform(-0.3, 1.3, ref.n)
    nsto = rng.normal(0.3, store_noise, ref.n_sto)
    csto = rng.normal(0.3, store_noise, ref.n_sto)
    von = rng.uniform(0.2, 0.6, ref.n_inst)
    ref.set_flags(**flags)
    ref.set_state(curr_sto=csto, next_sto=nsto, curr_sta=np.zeros(ref.n_sta))
    ref.set_von(von)
    eng.set_state(0, nsto); eng.set_state(1, csto)
    eng.b4_set_von(0, von)
    want = ref.load(x)
    got = eng.load_host(x, solver_state(**flags))
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(got[k], want[k], scale) < TOL, (variant, flags, k)
        # and entry-wise wherever no cancellation is involved
    # store / state vectors: entry-wise, with the same cancellation floor per slot kind (1e-3 of the largest
    # entry of that slot over all instances; e.g. Cgd is a difference of capacitances)
    st = ref.get_state()
    for which, key in ((0, "next_sto"), (2, "next_sta"), (3, "curr_sta")):
        got_s, want_s = eng.get_state(which), st[key]
        per = len(want_s) // ref.n_inst
        if per * ref.n_inst == len(want_s) and per > 0:
            got_s, want_s = got_s.reshape(ref.n_inst, per), want_s.reshape(ref.n_inst, per)
            floor = np.maximum(1e-3 * np.max(np.abs(want_s), axis=0, keepdims=True), 1e-300)
            assert np.max(np.abs(got_s - want_s) / np.maximum(np.abs(want_s), floor)) < TOL, key
        else:
            assert rel_err(got_s, want_s, 1e-30) < TOL, key
    assert rel_err(eng.b4_get_von(0, ref.n_inst), ref.get_von(), 1e-30) < TOL
    # DeviceMgr::allDevicesConverged = AND of Instance::isConverged() (= !limitedFlag for BSIM4)
    assert eng.all_converged() == all(ref.lib.xref_inst_converged(ref.h, i) for i in range(ref.n_inst))
    eng.close()


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_variants_transient(variant):
    run_case(variant, FLAG_CASES["tran_iter1"])


@pytest.mark.parametrize("case", sorted(FLAG_CASES))
def test_flag_cases_default_card(case):
    run_case("default", FLAG_CASES[case])


@pytest.mark.parametrize("case", ["tran_iter0", "dcop_initjct", "nolimit"])
@pytest.mark.parametrize("variant", ["rgate3", "rbody", "rdsmod", "igc"])
def test_flag_cases_general_topology(variant, case):
    run_case(variant, FLAG_CASES[case])


@pytest.mark.parametrize("arith", [0, 1, 2])
@pytest.mark.parametrize("variant", ["default", "igc", "capmod1", "rgate3", "rbody"])
def test_every_arithmetic_variant(variant, arith):
    # 0 strict (no FMA, IEEE division), 1 FMA contraction, 2 FMA + reciprocal division (library default)
    run_case(variant, FLAG_CASES["tran_iter1"], arith=arith)


@pytest.mark.parametrize("shape", [(64, 4), (64, 6), (96, 4), (128, 2), (128, 3), (128, 4), (256, 1), (384, 1), (512, 1)])
@pytest.mark.parametrize("uniform,lockstep", [(0, 0), (1, 0), (1, 1)])
def test_every_kernel_shape(shape, uniform, lockstep):
    # record access (per-thread loads / uniform records in the parameter block), lock-step barriers and
    # every compiled block shape give the same stamps; 300 pairs -> several blocks, ragged last block
    run_case("default", FLAG_CASES["tran_iter1"], n_pairs=300 if shape[0] < 512 else 700, sorted_bins=True,
             options=dict(b4_threads=shape[0], b4_minblocks=shape[1], b4_uniform=uniform, b4_lockstep=lockstep))


@pytest.mark.parametrize("variant", ["igc", "capmod1", "rdsmod", "rgate3"])
def test_lockstep_and_uniform_other_cards(variant):
    run_case(variant, FLAG_CASES["tran_iter1"], n_pairs=100, sorted_bins=True, options=dict(b4_uniform=1, b4_lockstep=1))
    run_case(variant, FLAG_CASES["dcop_initjct"], options=dict(b4_uniform=0))


def test_many_bins_fall_back_to_per_thread_records():
    # more than 64 (model, bin) runs in one group: the engine must pick the per-thread-record kernel and
    # reject lock-step (which needs block-uniform records)
    ref = isolated_devices(oracle_ref.RefCircuit, 40, "default", seed=9)     # N,P,N,P,... = 80 runs
    eng, _ = engine_from_ref(ref)
    x = np.random.default_rng(1).uniform(-0.3, 1.3, ref.n)
    ref.set_flags(**FLAG_CASES["tran_iter1"])
    want, got = ref.load(x), eng.load_host(x, solver_state(**FLAG_CASES["tran_iter1"]))
    assert rel_err(got["dFdx"], want["dFdx"], 1e-3 * np.max(np.abs(want["dFdx"]))) < TOL
    eng.set_option("b4_lockstep", 1)
    with pytest.raises(RuntimeError):
        eng.load_host(x, solver_state(**FLAG_CASES["tran_iter1"]))
    eng.close()


def test_pass_through_limiters():
    # previous iterate == present voltages: limiters take the pass-through branch (C2 operating point)
    run_case("default", FLAG_CASES["tran_iter1"], store_noise=0.0)


@pytest.mark.parametrize("variant", ["default", "igc", "capmod1", "gidl", "rgate", "rgate3", "rbody", "rdsmod", "rsh"])
def test_lead_currents(variant):
    """loadLeadCurrent (.PRINT I(M1)): leadF, leadQ, junctionV at the branch-data LIDs against Master::loadDAEVectors
    (N_DEV_MOSFET_B4.C:10933-10987) on the reference objects, 4-terminal devices and devices with internal nodes."""
    import torch
    ref = isolated_devices(oracle_ref.RefCircuit, 100, variant, seed=4, lead=True)
    eng, _ = engine_from_ref(ref)
    rng = np.random.default_rng(12)
    x = rng.uniform(-0.2, 1.2, ref.n)
    sto = rng.normal(0.3, 0.3, ref.n_sto)
    von = rng.uniform(0.2, 0.6, ref.n_inst)
    flags = dict(transient=1, newtonIter=1)
    ref.set_flags(**flags); ref.set_state(curr_sto=sto, next_sto=sto); ref.set_von(von)
    eng.set_state(0, sto); eng.set_state(1, sto); eng.b4_set_von(0, von)
    ref.load(x)
    want = ref.lead()
    eng.b4_lead_set(0, want["branch0"])
    eng.load_host(x, solver_state(**flags))
    nb = len(want["leadF"])
    out = [torch.full((nb,), 7.0, dtype=torch.float64, device="cuda") for _ in range(3)]
    eng.b4_lead_load(eng.device_buffer(0), *[t.data_ptr() for t in out])
    eng.sync()
    for key, t in zip(("leadF", "leadQ", "junctionV"), out):
        got = t.cpu().numpy()
        scale = 1e-3 * np.max(np.abs(want[key])) if np.any(want[key]) else 1e-300
        assert rel_err(got, want[key], scale) < TOL, (variant, key)
    eng.close()


# ---- against the committed golden fixtures (no oracle library needed at run time) ----
from b4_common import load_golden  # noqa: E402

GOLD = load_golden()


@pytest.mark.parametrize("case", sorted(GOLD))
def test_gpu_matches_reference_golden(case):
    import xyce_b200
    g = GOLD[case]
    eng = xyce_b200.Engine(0)
    eng.set_pattern(g["rowptr"], g["colind"])
    eng.set_sizes(int(g["n_sta"]), int(g["n_sto"]))
    eng.b4_set_models(g["rec_model_d"], g["rec_model_i"], g["rec_size_d"])
    eng.b4_add_group(g["rec_inst_d"], g["rec_inst_i"], g["rec_model_idx"], g["rec_size_idx"], g["rec_lids"],
                     g["rec_sto0"], 1, g["rec_sta0"], 1)
    eng.finalize()
    eng.set_state(0, g["nsto"]); eng.set_state(1, g["csto"]); eng.b4_set_von(0, g["von"])
    from b4_common import FLAG_NAMES
    flags = dict(zip(FLAG_NAMES, [int(v) for v in g["flags"]]))
    got = eng.load_host(g["x"], solver_state(**flags))
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        want = g["ref_" + k]
        scale = 1e-3 * np.max(np.abs(want)) if np.any(want) else 1e-300
        assert rel_err(got[k], want, scale) < TOL, (case, k)
    # store vector: entry-wise with the cancellation floor per slot kind, as in run_case (the published Cgs / Cgd are
    # differences of capacitances: -1.8e-20 F beside 1e-16 F terms in igc2_v461)
    sto, want_sto = eng.get_state(0).reshape(-1, 22), g["next_sto"].reshape(-1, 22)
    floor = np.maximum(1e-3 * np.max(np.abs(want_sto), axis=0, keepdims=True), 1e-300)
    assert np.max(np.abs(sto - want_sto) / np.maximum(np.abs(want_sto), floor)) < TOL
    assert rel_err(eng.get_state(2), g["next_sta"], 1e-30) < TOL
    assert rel_err(eng.get_state(3), g["curr_sta"], 1e-30) < TOL
    assert rel_err(eng.b4_get_von(0, len(g["von"])), g["von_out"], 1e-30) < TOL
    eng.close()


def test_fused_load_equals_separate_calls_bitwise():
    """xgpu_load_dae (one eval launch + three fused assembly launches) returns exactly the sums of
    xgpu_update_state + xgpu_load_vectors + xgpu_load_matrices: same contributions, same order."""
    from xyce_b200 import workloads as wl
    import torch
    w = wl.ring_oscillator_array(7, 31)
    eng = wl.build_engine(w)
    ss = solver_state(transient=1, newtonIter=1)
    n, nnz = w["n_unknowns"], eng.nnz
    dev = dict(dtype=torch.float64, device="cuda")
    x = torch.tensor(np.random.default_rng(0).uniform(0, 1, n), **dev)
    sta = [torch.zeros(w["n_state"], **dev) for _ in range(2)]
    sto = [torch.tensor(w["store"], **dev) for _ in range(2)]
    a = [torch.zeros(n, **dev) for _ in range(4)] + [torch.zeros(nnz, **dev) for _ in range(2)]
    b = [torch.full_like(t, 7.0) for t in a]
    st = (x.data_ptr(), sta[0].data_ptr(), sta[1].data_ptr(), sto[0].data_ptr(), sto[1].data_ptr(), ss)
    eng.update_state(*st)
    eng.load_vectors(*[t.data_ptr() for t in a[:4]], accumulate=False)
    eng.load_matrices(a[4].data_ptr(), a[5].data_ptr(), accumulate=False)
    eng.sync()
    # an evaluation updates the carried limiter state (store vector, von): rewind it before the second pass
    for t in sto:
        t.copy_(torch.tensor(w["store"], **dev))
    eng.b4_set_von(0, w["von"])
    eng.load_dae(*st, *[t.data_ptr() for t in b], accumulate=False)
    eng.sync()
    for p, q in zip(a, b):
        assert torch.equal(p, q)
    eng.close()


@pytest.mark.parametrize("zero_copy", [1, 0])
def test_load_host_with_pinned_buffers(zero_copy):
    """xgpu_load_host on pinned, mapped host buffers (the assembly kernel writes them directly over PCIe) gives bitwise
    the results of the pageable-buffer copy path"""
    import ctypes as C
    import torch
    ref = isolated_devices(oracle_ref.RefCircuit, 300, "default", seed=9, sorted_bins=True)
    eng, _ = engine_from_ref(ref)
    eng.set_option("zero_copy_out", zero_copy)
    rng = np.random.default_rng(10)
    x = rng.uniform(-0.2, 1.2, ref.n)
    sto = rng.normal(0.3, 0.3, ref.n_sto)
    von = rng.uniform(0.2, 0.6, ref.n_inst)
    ss = solver_state(transient=1, newtonIter=1)
    outs = []
    for pinned in (False, True):
        eng.set_state(0, sto); eng.set_state(1, sto); eng.b4_set_von(0, von)
        if not pinned:
            outs.append(eng.load_host(x, ss))
            continue
        hx = torch.tensor(x, dtype=torch.float64).pin_memory()
        bufs = [torch.full((ref.n,), 3.0, dtype=torch.float64).pin_memory() for _ in range(4)] + \
               [torch.full((eng.nnz,), 3.0, dtype=torch.float64).pin_memory() for _ in range(2)]
        ptr = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
        assert eng.lib.xgpu_load_host(eng.h, ptr(hx), C.byref(ss), *[ptr(t) for t in bufs]) == 0
        outs.append(dict(zip(("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"), [t.numpy().copy() for t in bufs])))
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    eng.close()


@pytest.mark.parametrize("variant", ["default", "igc", "igc2", "gidl", "capmod1", "rdsmod"])
def test_per_instance_stamps_with_row_wise_scale(variant):
    """Every terminal of every device sits on its own node, so each CSR row holds ONE instance's stamp row before any
    assembly with other devices.  Entries are compared relative to max(|entry|, 1e-3 * largest entry of the SAME row of the
    SAME instance) instead of a system-wide floor: the gate row (tunnelling currents, 1e-12 S and below) is held to its own
    scale beside the drain row (1e-3 S); the only cancellation left under the floor is the one inside a KCL row.  F and Q
    rows: relative to the largest |row term| of the same instance."""
    if variant not in VARIANTS:
        pytest.skip("no such card variant")
    ref = isolated_devices(oracle_ref.RefCircuit, 24, variant, seed=21)
    eng, _ = engine_from_ref(ref)
    rng = np.random.default_rng(22)
    x = rng.uniform(-0.2, 1.2, ref.n)
    sto = rng.normal(0.3, 0.2, ref.n_sto)
    von = rng.uniform(0.2, 0.6, ref.n_inst)
    flags = FLAG_CASES["tran_iter1"]
    ref.set_flags(**flags); ref.set_state(curr_sto=sto, next_sto=sto); ref.set_von(von)
    eng.set_state(0, sto); eng.set_state(1, sto); eng.b4_set_von(0, von)
    want, got = ref.load(x), eng.load_host(x, solver_state(**flags))
    rowptr = ref.rowptr
    worst = 0.0
    for k in ("dFdx", "dQdx"):
        w, g = want[k], got[k]
        for r in range(ref.n):
            a, b = rowptr[r], rowptr[r + 1]
            if b == a:
                continue
            floor = 1e-3 * np.max(np.abs(w[a:b]))
            if floor == 0.0:
                assert not np.any(g[a:b]), (k, r)
                continue
            worst = max(worst, float(np.max(np.abs(g[a:b] - w[a:b]) / np.maximum(np.abs(w[a:b]), floor))))
    assert worst < TOL, ("matrix", worst)
    # vectors: unknowns of one instance = its 4 external nodes (4 * i ...) and its internal nodes
    for k in ("f", "q"):
        w, g = want[k], got[k]
        for i in range(ref.n_inst):
            rows = [l for l in ref.inst_info(i)["lids"] if 0 <= l < ref.n]
            floor = 1e-3 * np.max(np.abs(w[rows]))
            if floor == 0.0:
                continue
            worst = max(worst, float(np.max(np.abs(g[rows] - w[rows]) / np.maximum(np.abs(w[rows]), floor))))
    assert worst < TOL, ("vectors", worst)
    eng.close()


@pytest.mark.parametrize("variant,want_spec", [("default", 0), ("igc", 1), ("capmod1", -1)])
def test_mode_specialised_objects_are_picked_by_the_cards_mode_tuple(variant, want_spec):
    """bsim4_spec_tuples.def: groups whose cards carry a listed mode tuple run that tuple's kernel object (same stamps
    at 1e-12, checked inside run_case), every other card the generic build."""
    ref = isolated_devices(oracle_ref.RefCircuit, 150, variant, seed=31, sorted_bins=True)
    eng, _ = engine_from_ref(ref)
    rng = np.random.default_rng(32)
    x = rng.uniform(-0.2, 1.2, ref.n)
    flags = FLAG_CASES["tran_iter1"]
    ref.set_flags(**flags)
    sto = rng.normal(0.3, 0.2, ref.n_sto); von = rng.uniform(0.2, 0.6, ref.n_inst)
    ref.set_state(curr_sto=sto, next_sto=sto); ref.set_von(von)
    eng.set_state(0, sto); eng.set_state(1, sto); eng.b4_set_von(0, von)
    want, got = ref.load(x), eng.load_host(x, solver_state(**flags))
    assert eng.b4_group_spec(0) == want_spec
    for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
        scale = 1e-3 * np.max(np.abs(want[k])) if np.any(want[k]) else 1e-300
        assert rel_err(got[k], want[k], scale) < TOL, (variant, k)
    eng.set_option("b4_spec", 0)
    eng.set_state(0, sto); eng.set_state(1, sto); eng.b4_set_von(0, von)
    gen = eng.load_host(x, solver_state(**flags))
    assert eng.b4_group_spec(0) == -1
    for k in ("f", "dFdx", "dQdx"):
        assert rel_err(gen[k], got[k], 1e-3 * np.max(np.abs(got[k]))) < TOL
    eng.close()
