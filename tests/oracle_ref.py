"""ctypes wrapper around oracle/_ref/libxyce_ref.so (the reference's own BSIM4 objects).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs.  Nothing under xyce_b200/ may import this module.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "..", "oracle", "_ref", "libxyce_ref.so")

FLAG_NAMES = ["dcop", "tranop", "acop", "transient", "dcsweep", "initJct", "initFix", "initTran",
              "newtonIter", "locaEnabled", "artParameter", "voltageLimiter"]


def available():
    return os.path.exists(REF_SO)


def _lib():
    lib = C.CDLL(REF_SO)
    lib.xref_new.restype = C.c_void_p
    lib.xref_b4_names.restype = C.c_char_p
    return lib


def _keys(d):
    ks = list(d.keys())
    arr = (C.c_char_p * len(ks))(*[k.encode() for k in ks])
    vals = np.array([float(d[k]) for k in ks], dtype=np.float64)
    return len(ks), arr, vals


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int)) if a is not None else None


class RefCircuit:
    """A netlist of BSIM4 instances evaluated by the reference's own C++ objects."""

    def __init__(self, n_ext_nodes):
        self.lib = _lib()
        self.h = C.c_void_p(self.lib.xref_new())
        self.lib.xref_set_num_external_nodes(self.h, int(n_ext_nodes))
        self.n_inst = 0
        self.n = None

    def add_model(self, name, mtype, params):
        n, ks, vs = _keys(params)
        rc = self.lib.xref_b4_add_model(self.h, name.encode(), mtype.encode(), n, ks, dptr(vs))
        assert rc == 0

    def add_instance(self, name, model, nodes, params):
        n, ks, vs = _keys(params)
        nd = np.array(nodes, dtype=np.int32)
        rc = self.lib.xref_b4_add_instance(self.h, name.encode(), model.encode(), iptr(nd), n, ks, dptr(vs))
        assert rc == 0
        self.n_inst += 1

    def add_dev_model(self, devtype, name, mtype, level, params):
        n, ks, vs = _keys(params)
        rc = self.lib.xref_add_model(self.h, devtype.encode(), name.encode(), mtype.encode(), int(level), n, ks, dptr(vs))
        assert rc == 0

    def add_dev_instance(self, devtype, name, model, nodes, params):
        n, ks, vs = _keys(params)
        nd = np.array(nodes, dtype=np.int32)
        rc = self.lib.xref_add_instance(self.h, devtype.encode(), name.encode(), model.encode(), len(nd), iptr(nd), n, ks, dptr(vs))
        assert rc == 0
        self.n_inst += 1

    def inst_info(self, idx):
        lids = np.zeros(16, dtype=np.int32)
        v = [C.c_int() for _ in range(4)]
        k = self.lib.xref_inst_info(self.h, idx, iptr(lids), 16, *[C.byref(x) for x in v])
        return dict(lids=lids[:k].copy(), sta0=v[0].value, sto0=v[1].value, nsta=v[2].value, nsto=v[3].value)

    def diode_export(self, idx):
        rec = np.zeros(64)
        flags = C.c_int()
        lids = np.zeros(3, dtype=np.int32)
        k = self.lib.xref_diode_export(self.h, idx, dptr(rec), C.byref(flags), iptr(lids))
        info = self.inst_info(idx)
        return dict(rec=rec[:k].copy(), flags=flags.value, lids=lids, sto0=info["sto0"], sta0=info["sta0"])

    def dev_export(self, idx, devtype):
        """record / flag word / node LIDs of a MOSFET level 1 ("m1") or BJT ("q") instance"""
        rec = np.zeros(96)
        flags = C.c_int()
        nn = {"m1": 6, "q": 7, "mvs": 7}[devtype]
        lids = np.zeros(nn, dtype=np.int32)
        if devtype == "mvs":          # ADMS-generated MVS 2.0.0 ETSOI: no flag word
            k = self.lib.xref_mvs_export(self.h, idx, dptr(rec), iptr(lids))
        else:
            fn = {"m1": self.lib.xref_mos1_export, "q": self.lib.xref_bjt_export}[devtype]
            k = fn(self.h, idx, dptr(rec), C.byref(flags), iptr(lids))
        info = self.inst_info(idx)
        return dict(rec=rec[:k].copy(), flags=flags.value, lids=lids, sto0=info["sto0"], sta0=info["sta0"])

    def adms_export(self, idx, name):
        """record (field order of the ADMS translator's evaluator) and unknown LIDs of an instance of a translated model"""
        rec = np.zeros(4096)            # PSP103 has 636 fields
        lids = np.zeros(64, dtype=np.int32)
        nl = C.c_int()
        k = self.lib.xref_adms_export(self.h, idx, name.encode(), dptr(rec), iptr(lids), C.byref(nl))
        assert k >= 0, "model %s is not in the oracle's ADMS registry" % name
        info = self.inst_info(idx)
        return dict(rec=rec[:k].copy(), flags=0, lids=lids[:nl.value].copy(), sto0=info["sto0"], sta0=info["sta0"])

    def enable_lead_currents(self):
        """DeviceInstance::enableLeadCurrentCalc on every instance (before finalize)"""
        self.lib.xref_enable_lead_currents(self.h)

    def lead(self):
        n = self.lib.xref_num_branch_data(self.h)
        out = [np.zeros(n) for _ in range(3)]
        self.lib.xref_get_lead(self.h, dptr(out[0]), dptr(out[1]), dptr(out[2]))
        return dict(leadF=out[0], leadQ=out[1], junctionV=out[2],
                    branch0=np.array([self.lib.xref_inst_branch0(self.h, i) for i in range(self.n_inst)], dtype=np.int32))

    def finalize(self):
        self.n = self.lib.xref_finalize(self.h)
        self.nnz = self.lib.xref_nnz(self.h)
        self.rowptr = np.zeros(self.n + 1, dtype=np.int32)
        self.colind = np.zeros(self.nnz, dtype=np.int32)
        self.lib.xref_pattern(self.h, iptr(self.rowptr), iptr(self.colind))
        self.n_sta = self.lib.xref_num_state(self.h)
        self.n_sto = self.lib.xref_num_store(self.h)
        return self.n

    def set_flags(self, gmin=1e-12, gainScale=1.0, nltermScale=1.0, **fl):
        f = np.zeros(len(FLAG_NAMES), dtype=np.int32)
        f[FLAG_NAMES.index("voltageLimiter")] = 1
        for k, v in fl.items():
            f[FLAG_NAMES.index(k)] = int(v)
        d = np.array([gmin, gainScale, nltermScale], dtype=np.float64)
        self.lib.xref_set_flags(self.h, iptr(f), dptr(d))
        self.flags = dict(zip(FLAG_NAMES, f.tolist()), gmin=gmin, gainScale=gainScale, nltermScale=nltermScale)

    def set_state(self, curr_sto=None, next_sto=None, curr_sta=None):
        cv = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        a, b, c = cv(curr_sto), cv(next_sto), cv(curr_sta)
        self.lib.xref_set_state(self.h, dptr(a), dptr(b), dptr(c))

    def set_step(self, curr_dt, last_dt, begin_integration):
        self.lib.xref_set_step.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
        self.lib.xref_set_step(self.h, curr_dt, last_dt, int(begin_integration))

    def last_store(self, vals=None):
        out = np.zeros(self.n_sto)
        v = None if vals is None else np.ascontiguousarray(vals, dtype=np.float64)
        self.lib.xref_last_store.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.xref_last_store(self.h, None if v is None else v.ctypes.data, out.ctypes.data)
        return out

    def get_state(self):
        cs, ns = np.zeros(self.n_sto), np.zeros(self.n_sto)
        ca, na = np.zeros(self.n_sta), np.zeros(self.n_sta)
        self.lib.xref_get_state(self.h, dptr(cs), dptr(ns), dptr(ca), dptr(na))
        return dict(curr_sto=cs, next_sto=ns, curr_sta=ca, next_sta=na)

    def set_von(self, von):
        v = np.ascontiguousarray(von, dtype=np.float64)
        self.lib.xref_b4_set_von(self.h, dptr(v))

    def get_von(self):
        v = np.zeros(self.n_inst)
        self.lib.xref_b4_get_von(self.h, dptr(v))
        return v

    def load(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = {k: np.zeros(self.n) for k in ("f", "q", "dFdxdVp", "dQdxdVp")}
        out["dFdx"] = np.zeros(self.nnz)
        out["dQdx"] = np.zeros(self.nnz)
        rc = self.lib.xref_load(self.h, dptr(x), dptr(out["f"]), dptr(out["q"]), dptr(out["dFdxdVp"]),
                                dptr(out["dQdxdVp"]), dptr(out["dFdx"]), dptr(out["dQdx"]))
        assert rc == 0
        return out

    # ---- the Xyce-side GPU adaptor (adaptor/N_DEV_GpuMaster_B4.h) in place of the stock BSIM4 Master ----
    def use_gpu_master(self, on=True):
        """call before the first BSIM4 model / instance is added"""
        self.lib.xref_use_gpu_master(self.h, int(on))

    def gpu_attach(self, device=0):
        """after finalize(): extract the records from the reference objects and upload them through the C ABI"""
        rc = self.lib.xref_gpu_attach(self.h, int(device))
        assert rc == 0, "xref_gpu_attach failed (%d)" % rc

    def all_converged(self):
        return bool(self.lib.xref_all_converged(self.h))

    def load_repeat(self, x, reps):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.lib.xref_set_solution(self.h, dptr(x))
        self.lib.xref_load_repeat(self.h, int(reps))

    def add_pattern_entries(self, rows, cols):
        r, c = np.ascontiguousarray(rows, dtype=np.int32), np.ascontiguousarray(cols, dtype=np.int32)
        self.lib.xref_add_pattern_entries(self.h, len(r), iptr(r), iptr(c))

    def tran_run(self, x0, tstop, tstep, probes, linear, sources, delmax=0.0, max_out=200000, method=0, dcop=0,
                 replay=None, pwl=None):
        """Transient run: tran_driver.h control flow around the reference device code + ksparse.
        replay = (h[], order[]): integrate on exactly these accepted steps (TranParams::replay_h) instead of
        selecting steps -- used to compare a sub-circuit with a larger run on that run's own time points."""
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        if replay is not None:
            rh, ro = f64(replay[0]), i32(replay[1])
            self.lib.xref_tran_replay(len(rh), dptr(rh), iptr(ro))
        if pwl is not None:
            tv = f64(np.asarray(pwl, dtype=np.float64).reshape(-1, 2))
            self.lib.xref_tran_pwl(len(tv), dptr(tv))
        par = f64([tstop, tstep, delmax, method, dcop])
        L = {k: (i32(v) if k.endswith(("row", "col")) else f64(v)) for k, v in linear.items()}
        S = dict(row=i32(sources["row"]), scale=f64(sources["scale"]), type=i32(sources["type"]), params=f64(sources["params"]))
        probes = i32(probes)
        times, wave = np.zeros(max_out), np.zeros((max_out, len(probes)))
        steps, stats = np.zeros((max_out, 5)), np.zeros(16)
        n_out, n_steps = C.c_int(), C.c_int()
        rc = self.lib.xref_tran_run(self.h, dptr(par), dptr(f64(x0)), len(L["g_row"]), iptr(L["g_row"]), iptr(L["g_col"]),
                                    dptr(L["g_val"]), len(L["c_row"]), iptr(L["c_row"]), iptr(L["c_col"]), dptr(L["c_val"]),
                                    len(S["row"]), iptr(S["row"]), dptr(S["scale"]), iptr(S["type"]), dptr(S["params"]),
                                    len(probes), iptr(probes), max_out, C.byref(n_out), dptr(times), dptr(wave),
                                    max_out, C.byref(n_steps), dptr(steps), dptr(stats))
        keys = ("accepted", "rejected", "newton_iters", "jacobian_loads", "residual_loads", "linear_solves",
                "lu_analyses", "lu_refactors", "time_points", "attempts", "driver_rc", "dcop_newton_iters", "dcop_status")
        return dict(rc=rc, t=times[:n_out.value], wave=wave[:n_out.value], steps=steps[:n_steps.value],
                    stats=dict(zip(keys, stats.tolist())))

    def names(self, which):
        return self.lib.xref_b4_names(which).decode().split()

    def export(self, idx):
        cnt = np.zeros(7, dtype=np.int32)
        self.lib.xref_b4_counts(iptr(cnt))
        md, mi = np.zeros(cnt[0]), np.zeros(cnt[1], dtype=np.int32)
        sd = np.zeros(cnt[2])
        idd, ii = np.zeros(cnt[3]), np.zeros(cnt[4], dtype=np.int32)
        mid, sid = C.c_longlong(), C.c_longlong()
        lids = np.zeros(12, dtype=np.int32)
        sta0, sto0 = C.c_int(), C.c_int()
        self.lib.xref_b4_export(self.h, idx, dptr(md), iptr(mi), dptr(sd), dptr(idd), iptr(ii),
                                C.byref(mid), C.byref(sid), iptr(lids), C.byref(sta0), C.byref(sto0))
        return dict(model_d=md, model_i=mi, size_d=sd, inst_d=idd, inst_i=ii, model_id=mid.value,
                    size_id=sid.value, lids=lids, sta0=sta0.value, sto0=sto0.value)

    def mid(self, idx):
        cnt = np.zeros(7, dtype=np.int32)
        self.lib.xref_b4_counts(iptr(cnt))
        d, i = np.zeros(cnt[5]), np.zeros(cnt[6], dtype=np.int32)
        self.lib.xref_b4_mid(self.h, idx, dptr(d), iptr(i))
        return dict(zip(self.names(5), d.tolist())), dict(zip(self.names(6), i.tolist()))
