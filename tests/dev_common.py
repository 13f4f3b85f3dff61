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


def diode_circuit(ref_cls, card, n_dev=6, seed=0):
    rng = np.random.default_rng(seed)
    c = ref_cls(2 * n_dev)
    p = dict(DIODE_CARDS[card])
    level = int(p.pop("LEVEL", 1))
    c.add_dev_model("d", "dmod", "D", level, p)
    for i in range(n_dev):
        c.add_dev_instance("d", "D:%d" % i, "dmod", [2 * i, 2 * i + 1], dict(AREA=float(rng.choice([1.0, 2.5]))))
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
