"""Shared BSIM4 test helpers: model cards, synthetic netlists, host-mirror wrapper, assembly in numpy."""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
import sys
sys.path.insert(0, os.path.join(HERE, ".."))
HOST_SO = os.path.join(HERE, "host_mirror", "libxb_host.so")

from xyce_b200.workloads import NMOS_CARD, PMOS_CARD  # noqa: E402  (benchmark cards are the base cards)

# model-card variants that switch on otherwise-dormant code paths
VARIANTS = {
    "default": ({}, {}),
    "igc": (dict(IGCMOD=1, IGBMOD=1), dict(IGCMOD=1, IGBMOD=1)),
    "igc2": (dict(IGCMOD=2, IGBMOD=1, TEMPMOD=2), dict(IGCMOD=2, IGBMOD=1, TEMPMOD=2)),
    "capmod0": (dict(CAPMOD=0), dict(CAPMOD=0, XPART=1.0)),
    "capmod1": (dict(CAPMOD=1, XPART=0.0), dict(CAPMOD=1, XPART=0.5)),
    "capmod2_xpart": (dict(CAPMOD=2, XPART=1.0), dict(CAPMOD=2, XPART=0.5)),
    "mob1": (dict(MOBMOD=1), dict(MOBMOD=2)),
    "mob3": (dict(MOBMOD=3), dict(MOBMOD=4)),
    "mob5": (dict(MOBMOD=5), dict(MOBMOD=6)),
    "diomod0": (dict(DIOMOD=0), dict(DIOMOD=2, BVS=5.0, XJBVS=1.0)),
    "gidl": (dict(AGIDL=1e-9, BGIDL=1e9, AGISL=1e-9, BGISL=1e9), dict(GIDLMOD=1, AGIDL=1e-9, BGIDL=1e9, AGISL=1e-9, BGISL=1e9)),
    "rdsmod": (dict(RDSMOD=1, RSH=5.0), dict(RDSMOD=1, RSH=5.0)),
    "rsh": (dict(RSH=8.0), dict(RSH=8.0)),
    "rgate": (dict(RGATEMOD=1, RSHG=2.0), dict(RGATEMOD=2, RSHG=2.0)),
    "rgate3": (dict(RGATEMOD=3, RSHG=2.0), dict(RGATEMOD=3, RSHG=2.0)),
    "rbody": (dict(RBODYMOD=1), dict(RBODYMOD=1)),
    "pocket": (dict(DVTP0=1e-7, DVTP1=0.1, LAMBDA=1e-9, VTL=2e5, ALPHA0=1e-7, BETA0=20.0),
               dict(DVTP0=1e-7, DVTP1=0.1, ALPHA0=1e-7, BETA0=20.0)),
    "cvcharge": (dict(CVCHARGEMOD=1), dict(CVCHARGEMOD=1)),
    "mtrl": (dict(MTRLMOD=1, EOT=1.7e-9, PHIG=4.2, EPSRGATE=11.7), dict(TNOIMOD=1)),
}

# the 4.7.0 and 4.6.1 evaluators (N_DEV_MOSFET_B4p70.C, N_DEV_MOSFET_B4p61.C) on the variants whose code differs
for _base, _vers in (("default", (4.7, 4.61)), ("igc2", (4.7, 4.61)), ("gidl", (4.7, 4.61)), ("mtrl", (4.7, 4.61)), ("pocket", (4.7, 4.61)), ("capmod0", (4.61,)),
                     ("mob1", (4.61,)), ("rdsmod", (4.61,)), ("diomod0", (4.61,)), ("igc", (4.61,))):
    for _v in _vers:
        VARIANTS["%s_v%d" % (_base, round(_v * 100))] = tuple({**c, "VERSION": _v} for c in VARIANTS[_base])
# mobMod 4-6 do not exist before 4.8 (the 4.7.0 code would read an unset VgsteffVth): high-k mobility on both cards
VARIANTS["mob3_v470"] = (dict(MOBMOD=3, VERSION=4.7), dict(MOBMOD=3, VERSION=4.7))

FLAG_NAMES = ["dcop", "tranop", "acop", "transient", "dcsweep", "initJct", "initFix", "initTran",
              "newtonIter", "locaEnabled", "artParameter", "voltageLimiter"]


def build_host_mirror():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "host_mirror")])
    return HOST_SO


class HostMirror:
    """ctypes wrapper of tests/host_mirror/libxb_host.so (the kernel source compiled for CPU)."""

    def __init__(self):
        build_host_mirror()
        self.lib = C.CDLL(HOST_SO)
        self.lib.xbh_b4_mid_names.restype = C.c_char_p
        self.mid_d_names = self.lib.xbh_b4_mid_names(0).decode().split()
        self.mid_i_names = self.lib.xbh_b4_mid_names(1).decode().split()
        self.slot_row = np.zeros(62, dtype=np.int32)
        self.slot_col = np.zeros(62, dtype=np.int32)
        self.lib.xbh_b4_slot_tables(self.slot_row.ctypes.data_as(C.POINTER(C.c_int)),
                                    self.slot_col.ctypes.data_as(C.POINTER(C.c_int)))

    def eval(self, rec, flags, V12, sto_old13, have_old, von_prev):
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        fl = np.array([flags.get(k, 1 if k == "voltageLimiter" else 0) for k in FLAG_NAMES], dtype=np.int32)
        fd = np.array([flags.get("gmin", 1e-12), flags.get("gainScale", 1.0), flags.get("nltermScale", 1.0)])
        out = dict(F=np.zeros(11), Q=np.zeros(11), FL=np.zeros(11), QL=np.zeros(11), JF=np.zeros(62),
                   JQ=np.zeros(62), store=np.zeros(22), state=np.zeros(3))
        md = np.zeros(len(self.mid_d_names))
        mi = np.zeros(len(self.mid_i_names), dtype=np.int32)
        V = np.ascontiguousarray(V12, dtype=np.float64)
        so = np.ascontiguousarray(sto_old13, dtype=np.float64)
        self.lib.xbh_b4_eval(dp(rec["model_d"]), ip(rec["model_i"]), dp(rec["size_d"]), dp(rec["inst_d"]),
                             ip(rec["inst_i"]), ip(fl), dp(fd), dp(V), dp(so), int(have_old), C.c_double(von_prev),
                             dp(out["F"]), dp(out["Q"]), dp(out["FL"]), dp(out["QL"]), dp(out["JF"]), dp(out["JQ"]),
                             dp(out["store"]), dp(out["state"]), dp(md), ip(mi))
        out["mid_d"] = dict(zip(self.mid_d_names, md.tolist()))
        out["mid_i"] = dict(zip(self.mid_i_names, mi.tolist()))
        return out


def isolated_devices(ref_cls, n_pairs, variant="default", seed=0, inst_extra=None, sorted_bins=False, lead=False):
    """n_pairs nmos + n_pairs pmos, every terminal on its own node (4 nodes per device).  sorted_bins orders
    the instances by (type, L, W) like an adaptor does, so that equal (model, bin) records form long runs."""
    rng = np.random.default_rng(seed)
    ndev = 2 * n_pairs
    c = ref_cls(4 * ndev)
    nv, pv = VARIANTS[variant]
    c.add_model("nch", "NMOS", {**NMOS_CARD, **nv})
    c.add_model("pch", "PMOS", {**PMOS_CARD, **pv})
    specs = [((i % 2 == 0), float(rng.choice([6e-8, 1e-7, 2.5e-7])), float(rng.choice([2e-7, 1e-6, 4e-6]))) for i in range(ndev)]
    if sorted_bins:
        specs.sort(key=lambda t: (not t[0], t[1], t[2]))
    for i, (is_n, L, W) in enumerate(specs):
        ip = dict(L=L, W=W, AD=2e-13, AS=2e-13, PD=2.4e-6, PS=2.4e-6)
        if inst_extra:
            ip.update(inst_extra)
        c.add_instance("M:%d" % i, "nch" if is_n else "pch", [4 * i, 4 * i + 1, 4 * i + 2, 4 * i + 3], ip)
    if lead:
        c.enable_lead_currents()
    c.finalize()
    return c


def assemble_general(hm, per_inst, lids, n, rowptr, colind):
    """Scatter per-instance general-stamp outputs into global vectors / CSR values (numpy, small cases)."""
    f, q, fl, ql = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    jf, jq = np.zeros(len(colind)), np.zeros(len(colind))
    for o, l in zip(per_inst, lids):
        for r in range(11):
            g = l[r]
            if g < 0:
                continue
            f[g] += o["F"][r]; q[g] += o["Q"][r]; fl[g] += o["FL"][r]; ql[g] += o["QL"][r]
        for s in range(62):
            gr, gc = l[hm.slot_row[s]], l[hm.slot_col[s]]
            if gr < 0 or gc < 0 or (o["JF"][s] == 0.0 and o["JQ"][s] == 0.0):
                continue
            seg = colind[rowptr[gr]:rowptr[gr + 1]]
            k = rowptr[gr] + int(np.searchsorted(seg, gc))
            assert colind[k] == gc, "stamp entry outside the reference's CSR pattern"
            jf[k] += o["JF"][s]; jq[k] += o["JQ"][s]
    return dict(f=f, q=q, dFdxdVp=fl, dQdxdVp=ql, dFdx=jf, dQdx=jq)


def records_from_ref(ref):
    """Collect de-duplicated model / size-bin records and per-instance records from a RefCircuit."""
    models, sizes = {}, {}
    md, mi, sd = [], [], []
    inst_d, inst_i, midx, sidx, lids, sto0, sta0 = [], [], [], [], [], [], []
    for i in range(ref.n_inst):
        e = ref.export(i)
        if e["model_id"] not in models:
            models[e["model_id"]] = len(md); md.append(e["model_d"]); mi.append(e["model_i"])
        if e["size_id"] not in sizes:
            sizes[e["size_id"]] = len(sd); sd.append(e["size_d"])
        inst_d.append(e["inst_d"]); inst_i.append(e["inst_i"])
        midx.append(models[e["model_id"]]); sidx.append(sizes[e["size_id"]])
        lids.append(e["lids"]); sto0.append(e["sto0"]); sta0.append(e["sta0"])
    return dict(model_d=np.array(md), model_i=np.array(mi, dtype=np.int32), size_d=np.array(sd),
                inst_d=np.array(inst_d), inst_i=np.array(inst_i, dtype=np.int32),
                model_idx=np.array(midx, dtype=np.int32), size_idx=np.array(sidx, dtype=np.int32),
                lids=np.array(lids, dtype=np.int32), sto0=np.array(sto0, dtype=np.int32),
                sta0=np.array(sta0, dtype=np.int32))


def engine_from_ref(ref, device=0):
    """Upload a RefCircuit's topology and Xyce-computed constants into a GPU Engine (what a
    GpuMaster adaptor does at setup time)."""
    import xyce_b200
    rec = records_from_ref(ref)
    eng = xyce_b200.Engine(device)
    eng.set_pattern(ref.rowptr, ref.colind)
    eng.set_sizes(ref.n_sta, ref.n_sto)
    eng.b4_set_models(rec["model_d"], rec["model_i"], rec["size_d"])
    eng.b4_add_group(rec["inst_d"], rec["inst_i"], rec["model_idx"], rec["size_idx"], rec["lids"],
                     rec["sto0"], 1, rec["sta0"], 1)
    eng.finalize()
    return eng, rec


def solver_state(**kw):
    import xyce_b200
    m = dict(dcop="dcopFlag", tranop="tranopFlag", acop="acopFlag", transient="transientFlag",
             dcsweep="dcsweepFlag", initJct="initJctFlag", initFix="initFixFlag", initTran="initTranFlag",
             newtonIter="newtonIter", locaEnabled="locaEnabledFlag", artParameter="artParameterFlag",
             voltageLimiter="voltageLimiterFlag")
    return xyce_b200.SolverState(**{m.get(k, k): v for k, v in kw.items()})


def rel_err(a, b, scale=None):
    """max |a-b| / max(|b|, scale): element-wise relative error with an absolute floor."""
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return 0.0
    s = np.maximum(np.abs(b), 1e-300 if scale is None else scale)
    return float(np.max(np.abs(a - b) / s))


def load_golden():
    """tests/golden/b4_cases.npz -> {case: dict}; outputs of the reference's own BSIM4 code (scripts/make_golden.py)."""
    z = np.load(os.path.join(HERE, "golden", "b4_cases.npz"))
    cases = {}
    for name in z["cases"]:
        pre = str(name) + "/"
        cases[str(name)] = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    return cases


def host_mirror_case(hm, g):
    """Evaluate one golden case with the host mirror of the kernel source and assemble in numpy."""
    flags = dict(zip(FLAG_NAMES, [int(v) for v in g["flags"]]))
    newton0 = flags["newtonIter"] == 0
    use_curr = newton0 and (not flags["dcop"] or flags["locaEnabled"])
    have_old = not (newton0 and not use_curr)
    sto_src = g["csto"] if use_curr else g["nsto"]
    per, lids = [], []
    for i in range(len(g["von"])):
        rec = {k: np.ascontiguousarray(g["rec_" + k][idx]) for k, idx in
               (("model_d", g["rec_model_idx"][i]), ("model_i", g["rec_model_idx"][i]),
                ("size_d", g["rec_size_idx"][i]), ("inst_d", i), ("inst_i", i))}
        l = g["rec_lids"][i]
        V = np.array([g["x"][t] if t >= 0 else 0.0 for t in l])
        so = sto_src[g["rec_sto0"][i]:g["rec_sto0"][i] + 13]
        per.append(hm.eval(rec, flags, V, so, have_old, float(g["von"][i])))
        lids.append(l)
    return per, assemble_general(hm, per, lids, len(g["x"]), g["rowptr"], g["colind"])


def ref_circuit_from_workload(ref_cls, w):
    """Build the reference-object circuit (oracle/_ref) for a workloads.* dict with BSIM4 + linear devices."""
    from xyce_b200 import workloads as wl
    c = ref_cls(w["n_unknowns"])
    c.add_model("nch", "NMOS", wl.NMOS_CARD)
    c.add_model("pch", "PMOS", wl.PMOS_CARD)
    for i in range(w["n_inst"]):
        is_n = w["kind"][i] == 0
        nodes = [int(v) for v in w["lids"][i][:4]]
        c.add_instance("M:%d" % i, "nch" if is_n else "pch", nodes, wl.NMOS_INST if is_n else wl.PMOS_INST)
    L = w["linear"]
    c.add_pattern_entries(np.concatenate([L["g_row"], L["c_row"]]), np.concatenate([L["g_col"], L["c_col"]]))
    c.finalize()
    return c
