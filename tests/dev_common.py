"""Helpers for the small compact models (diode, MOSFET level 1, BJT, rlc): netlists of isolated devices
evaluated by the reference objects (oracle/_ref), by the host mirror and by the GPU."""
import ctypes as C
import numpy as np

from b4_common import FLAG_NAMES, HOST_SO, build_host_mirror

DIODE_CARDS = {
    "basic": dict(IS=1e-14, N=1.05, RS=0.0, CJO=2e-12, VJ=0.8, M=0.45, TT=5e-9),
    "rs_bv": dict(IS=2e-14, N=1.1, RS=2.5, CJO=1e-12, VJ=0.7, M=0.33, TT=1e-9, BV=6.0, IBV=1e-6),
    "sidewall": dict(IS=1e-14, JSW=1e-13, NS=1.2, CJSW=1e-12, VJSW=0.7, MJSW=0.3, CJO=1e-12, RS=1.0, ISR=1e-13, NR=2.0, IKF=1e-2),
    "level2": dict(LEVEL=2, IS=1e-14, N=1.0, RS=0.5, CJO=1e-12, ISR=1e-12, NR=2.0, IKF=5e-3, BV=8.0, IBV=1e-5, NBV=2.0),
}


def flag_arrays(flags):
    fl = np.array([flags.get(k, 1 if k == "voltageLimiter" else 0) for k in FLAG_NAMES], dtype=np.int32)
    fd = np.array([flags.get("gmin", 1e-12), flags.get("gainScale", 1.0), flags.get("nltermScale", 1.0)])
    return fl, fd


def diode_circuit(ref_cls, card, n_dev=6, seed=0, lead=False):
    rng = np.random.default_rng(seed)
    c = ref_cls(2 * n_dev)
    p = dict(DIODE_CARDS[card])
    level = int(p.pop("LEVEL", 1))
    c.add_dev_model("d", "dmod", "D", level, p)
    for i in range(n_dev):
        c.add_dev_instance("d", "D:%d" % i, "dmod", [2 * i, 2 * i + 1], dict(AREA=float(rng.choice([1.0, 2.5]))))
    if lead:
        c.enable_lead_currents()
    c.finalize()
    return c


class HostDevices:
    def __init__(self):
        build_host_mirror()
        self.lib = C.CDLL(HOST_SO)

    def diode(self, e, flags, V3, vd_curr, vd_next):
        fl, fd = flag_arrays(flags)
        out = np.zeros(30)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        V = np.ascontiguousarray(V3, dtype=np.float64)
        self.lib.xbh_diode_eval(dp(np.ascontiguousarray(e["rec"])), int(e["flags"]), ip(fl), dp(fd), dp(V),
                                C.c_double(vd_curr), C.c_double(vd_next), dp(out))
        return dict(F=out[0:3], Q=out[3:6], FL=out[6:9], QL=out[9:12], JF=out[12:19], JQ=out[19:26], store=out[26:29],
                    orig=int(out[29]))


    def simple(self, type_id, e, flags, V, curr_sto, next_sto, curr_sta, nodes, slots, nstore, nstate):
        """MOSFET level 1 (type 2) / BJT (type 3) through the host build of the kernel source"""
        fl, fd = flag_arrays(flags)
        out = np.zeros(4 * nodes + 2 * slots + nstore + nstate + 1)
        dp = lambda a: np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (e["rec"], V, curr_sto, next_sto, curr_sta)]
        k = self.lib.xbh_simple_eval(int(type_id), dp(keep[0]), int(e["flags"]), ip(fl), dp(fd), dp(keep[1]), dp(keep[2]),
                                     dp(keep[3]), dp(keep[4]), out.ctypes.data_as(C.POINTER(C.c_double)))
        assert k == len(out)
        n, s = nodes, slots
        return dict(F=out[0:n], Q=out[n:2 * n], FL=out[2 * n:3 * n], QL=out[3 * n:4 * n], JF=out[4 * n:4 * n + s],
                    JQ=out[4 * n + s:4 * n + 2 * s], store=out[4 * n + 2 * s:4 * n + 2 * s + nstore],
                    state=out[4 * n + 2 * s + nstore:4 * n + 2 * s + nstore + nstate], orig=int(out[-1]))

    def bjt_xp(self, e, flags, step3, V, curr_sto, next_sto, cex2):
        """BJT with the step history of the excess phase; returns (outputs like simple(), [mode, next, init])"""
        nodes, slots, nstore, nstate = 7, 24, 3, 6
        fl, fd = flag_arrays(flags)
        out = np.zeros(4 * nodes + 2 * slots + nstore + nstate + 1)
        cex_out = np.zeros(3)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (e["rec"], step3, V, curr_sto, next_sto, cex2)]
        self.lib.xbh_bjt_eval_xp.restype = C.c_int
        k = self.lib.xbh_bjt_eval_xp(dp(keep[0]), int(e["flags"]), fl.ctypes.data_as(C.POINTER(C.c_int)), dp(fd), dp(keep[1]),
                                     dp(keep[2]), dp(keep[3]), dp(keep[4]), dp(keep[5]), dp(out), dp(cex_out))
        assert k == len(out)
        n, s = nodes, slots
        o = dict(F=out[0:n], Q=out[n:2 * n], FL=out[2 * n:3 * n], QL=out[3 * n:4 * n], JF=out[4 * n:4 * n + s],
                 JQ=out[4 * n + s:4 * n + 2 * s], store=out[4 * n + 2 * s:4 * n + 2 * s + nstore],
                 state=out[4 * n + 2 * s + nstore:4 * n + 2 * s + nstore + nstate], orig=int(out[-1]))
        return o, cex_out


MOS1_SLOT_ROW = [0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5]
MOS1_SLOT_COL = [0, 4, 1, 3, 4, 5, 2, 5, 1, 3, 4, 5, 0, 1, 3, 4, 5, 1, 2, 3, 4, 5]
BJT_SLOT_ROW = [0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6]
BJT_SLOT_COL = [0, 4, 1, 4, 5, 6, 2, 6, 3, 4, 0, 1, 3, 4, 5, 6, 1, 4, 5, 6, 2, 4, 5, 6]
# (type id, devtype key of the oracle harness, nodes, store entries used, state entries)
MVS_SLOT_ROW = [0, 0, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 6, 6, 6, 6]
MVS_SLOT_COL = [3, 0, 2, 4, 4, 3, 5, 0, 4, 3, 5, 2, 6, 5, 4, 1, 3]
SIMPLE = {"mos1": (2, "m1", 6, 6, 8, MOS1_SLOT_ROW, MOS1_SLOT_COL), "bjt": (3, "q", 7, 3, 6, BJT_SLOT_ROW, BJT_SLOT_COL),
          "mvs": (5, "mvs", 7, 0, 0, MVS_SLOT_ROW, MVS_SLOT_COL)}
# ADMS-generated MVS 2.0.0 ETSOI model cards (src/DeviceModelPKG/ADMS/N_DEV_ADMSmvs_2_0_0_etsoi.C:119-250 defaults = "nmos")
MVS_CARDS = {
    "nmos": ("NMOS", dict(TYPE=1)),
    "pmos": ("PMOS", dict(TYPE=-1, W=2e-6, LGDR=60e-9, RS0=200e-6, N0=1.5, DELTA=0.15, ND=0.05, MU_EFF=0.6, KSEE=0.15)),
    "short": ("NMOS", dict(TYPE=1, LGDR=32e-9, DLG=6e-9, BETA=1.8, THETA=2.2, TJUN=350.0, NU=0.6, ENERGY_DIFF_VOLT=0.1)),
}

MOS1_CARDS = {
    "basic": ("NMOS", dict(VTO=0.7, KP=1.1e-4, GAMMA=0.4, PHI=0.65, LAMBDA=0.02, TOX=2e-8, CBD=2e-14, CBS=2e-14, IS=1e-14,
                            CGSO=2e-10, CGDO=2e-10, CGBO=1e-10)),
    "pmos_rs": ("PMOS", dict(VTO=-0.8, KP=4e-5, GAMMA=0.5, PHI=0.7, LAMBDA=0.03, TOX=2e-8, RD=15.0, RS=12.0, CJ=3e-4, CJSW=2e-10,
                              MJ=0.4, MJSW=0.3, PB=0.85, JS=1e-6, CGSO=1.5e-10, CGDO=1.5e-10)),
    "sheet": ("NMOS", dict(VTO=0.5, KP=8e-5, GAMMA=0.3, PHI=0.6, TOX=1.5e-8, RSH=20.0, CJ=2e-4, MJ=0.5, CJSW=1e-10, MJSW=0.5,
                            NSUB=1e16, UO=500.0)),
}
BJT_CARDS = {
    "basic": ("NPN", dict(IS=1e-15, BF=120.0, BR=2.0, VAF=80.0, CJE=1e-12, CJC=5e-13, TF=3e-10, TR=1e-8)),
    "res_pnp": ("PNP", dict(IS=2e-15, BF=80.0, BR=1.5, VAF=60.0, VAR=20.0, IKF=5e-2, IKR=1e-2, ISE=1e-13, NE=1.6, ISC=1e-13, NC=1.8,
                             RB=50.0, RBM=10.0, IRB=1e-3, RC=8.0, RE=1.5, CJE=1.2e-12, VJE=0.8, MJE=0.35, CJC=6e-13, VJC=0.7,
                             MJC=0.4, XCJC=0.7, CJS=2e-13, VJS=0.6, MJS=0.3, TF=2e-10, XTF=2.0, VTF=3.0, ITF=0.05, TR=5e-9)),
    "hicur": ("NPN", dict(IS=5e-16, BF=200.0, IKF=1e-2, NK=0.6, RB=100.0, RE=1.0, CJE=8e-13, CJC=4e-13, TF=1e-10, XTF=1.0, ITF=0.02)),
}

# excess phase: PTF degrees at f = 1 / (2 pi TF) -> Weil's approximation on the step history (N_DEV_BJT.C:2706-2799);
# kept out of BJT_CARDS: it needs the step sizes set (0 / 0 otherwise, in the reference too)
BJT_PTF_CARDS = {
    "ptf": ("NPN", dict(IS=1e-15, BF=150.0, BR=2.0, VAF=70.0, IKF=2e-2, RB=40.0, RC=5.0, CJE=1e-12, CJC=5e-13, TF=2e-10, TR=5e-9,
                        PTF=35.0)),
}


def simple_circuit(ref_cls, kind, card, n_dev=6, seed=0, lead=False):
    """n_dev isolated devices, every terminal on its own node"""
    rng = np.random.default_rng(seed)
    nt = 4
    c = ref_cls(nt * n_dev)
    if kind == "mos1":
        mtype, p = MOS1_CARDS[card]
        c.add_dev_model("m1", "mmod", mtype, 1, p)
        for i in range(n_dev):
            ip = dict(L=float(rng.choice([1e-6, 2e-6])), W=float(rng.choice([5e-6, 2e-5])), AD=2e-11, AS=2e-11, PD=2e-5, PS=2e-5,
                      NRD=1.0, NRS=1.0)
            c.add_dev_instance("m1", "M:%d" % i, "mmod", [nt * i, nt * i + 1, nt * i + 2, nt * i + 3], ip)
    elif kind == "mvs":
        mtype, p = MVS_CARDS[card]
        c.add_dev_model("mvs", "vsmod", mtype, 1, p)
        for i in range(n_dev):
            c.add_dev_instance("mvs", "M:%d" % i, "vsmod", [nt * i, nt * i + 1, nt * i + 2], {})
    else:
        mtype, p = {**BJT_CARDS, **BJT_PTF_CARDS}[card]
        c.add_dev_model("q", "qmod", mtype, 1, p)
        for i in range(n_dev):
            c.add_dev_instance("q", "Q:%d" % i, "qmod", [nt * i, nt * i + 1, nt * i + 2, nt * i + 3], dict(AREA=float(rng.choice([1.0, 3.0]))))
    if lead:
        c.enable_lead_currents()
    c.finalize()
    return c


DIODE_SLOT_ROW = [0, 0, 1, 1, 2, 2, 2]
DIODE_SLOT_COL = [0, 2, 1, 2, 0, 1, 2]


def assemble(per, lids, slot_row, slot_col, n, rowptr, colind):
    f, q, fl, ql = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    jf, jq = np.zeros(len(colind)), np.zeros(len(colind))
    for o, l in zip(per, lids):
        for r in range(len(l)):
            if l[r] < 0:
                continue
            f[l[r]] += o["F"][r]; q[l[r]] += o["Q"][r]; fl[l[r]] += o["FL"][r]; ql[l[r]] += o["QL"][r]
        for s in range(len(slot_row)):
            gr, gc = l[slot_row[s]], l[slot_col[s]]
            if gr < 0 or gc < 0:
                continue
            k = rowptr[gr] + int(np.searchsorted(colind[rowptr[gr]:rowptr[gr + 1]], gc))
            assert colind[k] == gc
            jf[k] += o["JF"][s]; jq[k] += o["JQ"][s]
    return dict(f=f, q=q, dFdxdVp=fl, dQdxdVp=ql, dFdx=jf, dQdx=jq)


def mixed_netlist():
    """BASELINE config 5 shape: diode clipper + Gummel-Poon common-emitter stage + MOSFET level 1 inverter + R, C, V in
    one netlist, built on the reference's own Diode / BJT / MOSFET1 objects.  Returns (ref, lin, src, x0, probes)."""
    import oracle_ref
    IN, A, VCC, B, C, E, D, BR_IN, BR_CC = range(9)
    ref = oracle_ref.RefCircuit(9)
    dp = dict(DIODE_CARDS["rs_bv"]); dp.pop("LEVEL", None)
    qt, qp = BJT_CARDS["basic"]; mt, mp = MOS1_CARDS["basic"]
    ref.add_dev_model("d", "dmod", "D", 1, dp)
    ref.add_dev_model("q", "qmod", qt, 1, dict(qp, RB=20.0, RC=5.0, RE=0.5))
    ref.add_dev_model("m1", "mmod", mt, 1, dict(mp, RD=10.0, RS=10.0))
    ref.add_dev_instance("d", "D:1", "dmod", [A, -1], dict(AREA=1.0))
    ref.add_dev_instance("d", "D:2", "dmod", [-1, A], dict(AREA=1.0))
    ref.add_dev_instance("q", "Q:1", "qmod", [C, B, E, -1], dict(AREA=1.0))
    ref.add_dev_instance("m1", "M:1", "mmod", [D, A, -1, -1], dict(L=2e-6, W=2e-5, AD=2e-11, AS=2e-11, PD=2e-5, PS=2e-5))
    g, c = [], []
    def res(a, b, r):
        gg = 1.0 / r
        for (i, j, v) in ((a, a, gg), (a, b, -gg), (b, a, -gg), (b, b, gg)):
            if i >= 0 and j >= 0: g.append((i, j, v))
    def cap(a, b, v):
        for (i, j, s) in ((a, a, v), (a, b, -v), (b, a, -v), (b, b, v)):
            if i >= 0 and j >= 0: c.append((i, j, s))
    def vsrc(node, br):
        g.append((node, br, 1.0)); g.append((br, node, 1.0))
    vsrc(IN, BR_IN); vsrc(VCC, BR_CC)
    res(IN, A, 1e3); cap(IN, B, 1e-10); res(VCC, B, 47e3); res(B, -1, 10e3); res(VCC, C, 2.2e3); res(E, -1, 470.0)
    res(VCC, D, 10e3); cap(D, -1, 1e-12); cap(C, -1, 2e-12)
    lin = dict(g_row=np.array([t[0] for t in g], dtype=np.int32), g_col=np.array([t[1] for t in g], dtype=np.int32),
               g_val=np.array([t[2] for t in g]), c_row=np.array([t[0] for t in c], dtype=np.int32),
               c_col=np.array([t[1] for t in c], dtype=np.int32), c_val=np.array([t[2] for t in c]))
    src = dict(row=np.array([BR_IN, BR_CC], dtype=np.int32), scale=np.ones(2), type=np.array([2, 0], dtype=np.int32),
               params=np.array([[0.0, 2.0, 1e6, 0, 0, 0, 0], [5.0, 0, 0, 0, 0, 0, 0]]))
    ref.add_pattern_entries(np.concatenate([lin["g_row"], lin["c_row"]]), np.concatenate([lin["g_col"], lin["c_col"]]))
    ref.finalize()
    x0 = np.zeros(ref.n); x0[VCC] = 5.0
    probes = [IN, A, B, C, D, BR_CC]
    return ref, lin, src, x0, probes


def mixed_engine(ref, lin, src):
    """The GPU engine for mixed_netlist(): records exported from the reference objects (what an adaptor uploads)."""
    import xyce_b200
    eng = xyce_b200.Engine(0)
    eng.set_pattern(ref.rowptr, ref.colind)
    eng.set_sizes(ref.n_sta, ref.n_sto)
    ex = [ref.diode_export(i) for i in (0, 1)]
    eng.add_simple_group(1, np.array([e["rec"] for e in ex]), [e["flags"] for e in ex], np.array([e["lids"] for e in ex]),
                         [e["sto0"] for e in ex], 1, [e["sta0"] for e in ex], 1)
    for tid, key, idx in ((3, "q", 2), (2, "m1", 3)):
        e = ref.dev_export(idx, key)
        eng.add_simple_group(tid, np.array([e["rec"]]), [e["flags"]], np.array([e["lids"]]), [e["sto0"]], 1, [e["sta0"]], 1)
    eng.set_linear(lin["g_row"], lin["g_col"], lin["g_val"], lin["c_row"], lin["c_col"], lin["c_val"])
    eng.set_sources(src["row"], src["scale"], src["type"], src["params"])
    eng.finalize()
    return eng
