"""ctypes binding of include/xyce_b200.h.  No CPU fallback: a missing library or GPU raises."""
import ctypes as C
import os
import numpy as np

# XYCE_B200_LIB selects another build of the same library (kernel experiments: scripts/build_variant.sh)
LIB_PATH = os.environ.get("XYCE_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libxyce_b200.so")
_lib = None

FLAG_FIELDS = ["dcopFlag", "tranopFlag", "acopFlag", "transientFlag", "dcsweepFlag", "initJctFlag", "initFixFlag",
               "initTranFlag", "newtonIter", "locaEnabledFlag", "artParameterFlag", "voltageLimiterFlag"]
REAL_FIELDS = ["gmin", "gainScale", "nltermScale", "vgstConst", "vdsScaleMin", "sizeScale", "currTimeStep", "lastTimeStep"]


class TranParams(C.Structure):
    """Mirror of xgpu_tran_params; zeros select the reference defaults."""
    _fields_ = [("tstop", C.c_double), ("tstep", C.c_double), ("delmax", C.c_double), ("maxNewtonStep", C.c_int),
                ("deltaXTol", C.c_double), ("absTol", C.c_double), ("relTol", C.c_double), ("RHSTol", C.c_double),
                ("relErrorTol", C.c_double), ("absErrorTol", C.c_double), ("maxOrder", C.c_int), ("maxSteps", C.c_int),
                ("method", C.c_int), ("dcop", C.c_int)]


class SolverState(C.Structure):
    """Mirror of xgpu_solver_state (Device::SolverState + DeviceOptions subset)."""
    _fields_ = [(n, C.c_int) for n in FLAG_FIELDS] + [(n, C.c_double) for n in REAL_FIELDS] + [("beginIntegrationFlag", C.c_int)]

    def __init__(self, **kw):
        super().__init__()
        # DeviceOptions defaults: Core/N_DEV_DeviceOptions.C:79-150
        self.voltageLimiterFlag = 1
        self.gmin = 1e-12
        self.gainScale = 1.0
        self.nltermScale = 1.0
        self.vgstConst = 4.5
        self.vdsScaleMin = 0.3
        self.sizeScale = 1.0
        for k, v in kw.items():
            setattr(self, k, v)


def load_library():
    """Load the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("xyce_b200: %s is missing -- run __graft_entry__.build() (nvcc, sm_100a); "
                           "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.xgpu_last_error.restype = C.c_char_p
    lib.xgpu_b4_field_names.restype = C.c_char_p
    lib.xgpu_device_buffer.restype = C.c_void_p
    lib.xgpu_launch_count.restype = C.c_longlong
    lib.xgpu_jacobian_combine.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    """One GPU context (one process per GPU)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.xgpu_create(int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError("xgpu_create failed with code %d (a CUDA device is required; no CPU fallback)" % rc)
        self.h = h
        self.n = 0
        self.nnz = 0
        # optional overrides of the kernel variant (see xgpu_set_option in include/xyce_b200.h)
        for env, opt in (("XYCE_B200_B4_ARITH", "b4_arith"), ("XYCE_B200_B4_MINBLOCKS", "b4_minblocks"),
                         ("XYCE_B200_B4_THREADS", "b4_threads"), ("XYCE_B200_B4_UNIFORM", "b4_uniform"),
                         ("XYCE_B200_B4_LOCKSTEP", "b4_lockstep"), ("XYCE_B200_B4_SPEC", "b4_spec")):
            if os.environ.get(env):
                self.set_option(opt, int(os.environ[env]))

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("xyce_b200 error %d: %s" % (rc, self.lib.xgpu_last_error(self.h).decode()))

    def close(self):
        if self.h:
            self.lib.xgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.lib.xgpu_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_option(self, name, value):
        self._chk(self.lib.xgpu_set_option(self.h, name.encode(), int(value)))

    def sync(self):
        self._chk(self.lib.xgpu_sync(self.h))

    def set_pattern(self, rowptr, colind):
        rowptr, colind = _i32(rowptr), _i32(colind)
        self.n, self.nnz = len(rowptr) - 1, len(colind)
        self._chk(self.lib.xgpu_pattern_set(self.h, self.n, _ip(rowptr), _ip(colind)))

    def build_pattern(self, n_unknowns):
        """Derive the CSR pattern from the device stamps (call after the groups were added)."""
        self._chk(self.lib.xgpu_pattern_build(self.h, int(n_unknowns)))
        self.n, self.nnz = int(n_unknowns), int(self.lib.xgpu_pattern_nnz(self.h))
        rowptr, colind = np.zeros(self.n + 1, dtype=np.int32), np.zeros(self.nnz, dtype=np.int32)
        self._chk(self.lib.xgpu_pattern_get(self.h, _ip(rowptr), _ip(colind)))
        return rowptr, colind

    def measure_fp64_peak(self):
        v = C.c_double()
        self._chk(self.lib.xgpu_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def pipe_info(self):
        """(pipelines?, rows in the first window, nonzeros in the first window) of the pipelined host-buffer path"""
        v = (C.c_longlong * 3)()
        self._chk(self.lib.xgpu_pipe_info(self.h, v))
        return int(v[0]), int(v[1]), int(v[2])

    def b4_group_spec(self, group=0):
        """id of the mode-specialised kernel object the group's last evaluation ran, -1 = generic build"""
        return int(self.lib.xgpu_b4_group_spec(self.h, int(group)))

    @staticmethod
    def adms_gen_models():
        """Models of the ADMS translator compiled into the library (xgpu_adms_gen_info): list of dicts with
        name, type (for add_simple_group), nodes, ext, slots, fields, slot_row, slot_col."""
        lib = load_library()
        out = []
        for idx in range(lib.xgpu_adms_gen_count()):
            name, fields = C.c_char_p(), C.c_char_p()
            info = (C.c_int32 * 5)()
            rows, cols = (C.c_int32 * 512)(), (C.c_int32 * 512)()
            if lib.xgpu_adms_gen_info(idx, C.byref(name), C.byref(fields), info, rows, cols) != 0:
                continue
            out.append(dict(name=name.value.decode(), type=info[0], nodes=info[1], ext=info[2], slots=info[3], nstore=int(lib.xgpu_simple_store_count(info[0])),
                            fields=fields.value.decode().split(), slot_row=list(rows[:info[3]]), slot_col=list(cols[:info[3]])))
        return out

    def newton_step_host(self, x, ss, qscalar, fscalar, hist=None, dx=None, rhs=None):
        """xgpu_newton_step_host: one Newton iteration (load, stamp, J and r, LU refactor, solves) behind host buffers."""
        x = _f64(x)
        dx = np.empty(self.n) if dx is None else dx
        hist = None if hist is None else _f64(hist)
        self._chk(self.lib.xgpu_newton_step_host(self.h, _dp(x), C.byref(ss), C.c_double(qscalar), C.c_double(fscalar),
                                                  _dp(hist), _dp(dx), _dp(rhs)))
        return dx

    def selftest_fastmath(self, which, a, b=None):
        """fast-variant exp (0) / log (1) / a/b (2) of the BSIM4 kernel, evaluated on the device."""
        a = _f64(a)
        b = None if b is None else _f64(b)
        out = np.empty_like(a)
        self._chk(self.lib.xgpu_selftest_fastmath(self.h, int(which), len(a), _dp(a), _dp(b), _dp(out)))
        return out

    def set_sizes(self, n_state, n_store):
        self.n_state, self.n_store = int(n_state), int(n_store)
        self._chk(self.lib.xgpu_sizes_set(self.h, self.n_state, self.n_store))

    def field_names(self, which):
        return self.lib.xgpu_b4_field_names(which).decode().split()

    def b4_set_models(self, model_d, model_i, size_d):
        model_d, model_i, size_d = _f64(model_d), _i32(model_i), _f64(size_d)
        self._chk(self.lib.xgpu_b4_models_set(self.h, model_d.shape[0], _dp(model_d), _ip(model_i),
                                              size_d.shape[0], _dp(size_d)))

    def b4_add_group(self, inst_d, inst_i, model_idx, size_idx, lids12, sto_lid0, sto_stride, sta_lid0, sta_stride):
        inst_d, inst_i = _f64(inst_d), _i32(inst_i)
        a = [_i32(x) for x in (model_idx, size_idx, lids12, sto_lid0, sta_lid0)]
        gid = self.lib.xgpu_b4_group_add(self.h, inst_d.shape[0], _dp(inst_d), _ip(inst_i), _ip(a[0]), _ip(a[1]),
                                         _ip(a[2]), _ip(a[3]), int(sto_stride), _ip(a[4]), int(sta_stride))
        if gid < 0:
            raise RuntimeError("xgpu_b4_group_add failed (%d): %s" % (gid, self.lib.xgpu_last_error(self.h).decode()))
        return gid

    def add_simple_group(self, dev_type, rec, flags, lids, sto_lid0, sto_stride=1, sta_lid0=None, sta_stride=1):
        """dev_type: 1 diode, 2 MOSFET level 1, 3 BJT, 4 ADMS-shaped RLC; rec: [n, nfields] records."""
        rec, flags, lids, sto_lid0 = _f64(rec), _i32(flags), _i32(lids), _i32(sto_lid0)
        sta = _i32(sta_lid0) if sta_lid0 is not None else None
        gid = self.lib.xgpu_simple_group_add(self.h, int(dev_type), rec.shape[0], _dp(rec), _ip(flags), _ip(lids),
                                             _ip(sto_lid0), int(sto_stride), _ip(sta), int(sta_stride))
        if gid < 0:
            raise RuntimeError("xgpu_simple_group_add failed (%d): %s" % (gid, self.lib.xgpu_last_error(self.h).decode()))
        return gid

    def finalize(self):
        self._chk(self.lib.xgpu_finalize(self.h))

    def b4_set_von(self, group, von):
        von = _f64(von)
        self._chk(self.lib.xgpu_b4_von_set(self.h, group, _dp(von)))

    def b4_get_von(self, group, n):
        von = np.zeros(n)
        self._chk(self.lib.xgpu_b4_von_get(self.h, group, _dp(von)))
        return von

    def set_state(self, which, vals):
        vals = _f64(vals)
        self._chk(self.lib.xgpu_state_set(self.h, which, _dp(vals)))

    def get_state(self, which):
        out = np.zeros(self.n_store if which < 2 or which == 4 else self.n_state)      # 4 = last store
        self._chk(self.lib.xgpu_state_get(self.h, which, _dp(out)))
        return out

    def load_host(self, x, ss, want_matrices=True):
        """Host-buffer path: H2D solution, evaluate + assemble on the GPU, D2H results."""
        x = _f64(x)
        out = {k: np.zeros(self.n) for k in ("f", "q", "dFdxdVp", "dQdxdVp")}
        out["dFdx"] = np.zeros(self.nnz) if want_matrices else None
        out["dQdx"] = np.zeros(self.nnz) if want_matrices else None
        self._chk(self.lib.xgpu_load_host(self.h, _dp(x), C.byref(ss), _dp(out["f"]), _dp(out["q"]),
                                          _dp(out["dFdxdVp"]), _dp(out["dQdxdVp"]), _dp(out["dFdx"]), _dp(out["dQdx"])))
        return out

    def load_host_jr(self, x, ss, qscalar, fscalar):
        """Host-buffer path for a host-side Newton solver: J = qscalar dQdx + fscalar dFdx and the device part of the residual."""
        x = _f64(x)
        rhs, jac = np.zeros(self.n), np.zeros(self.nnz)
        self._chk(self.lib.xgpu_load_host_jr(self.h, _dp(x), C.byref(ss), C.c_double(qscalar), C.c_double(fscalar), _dp(rhs), _dp(jac)))
        return rhs, jac

    # ---- bordered solve / multi-GPU ----
    @staticmethod
    def comm_unique_id():
        lib = load_library()
        buf = (C.c_ubyte * 128)()
        rc = lib.xgpu_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError("xgpu_comm_unique_id failed with code %d (libnccl.so.2 missing?)" % rc)
        return bytes(buf)

    def comm_init(self, id128, rank, world):
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(id128))
        self._chk(self.lib.xgpu_comm_init(self.h, buf, int(rank), int(world)))

    def p2p_handle(self):
        buf = (C.c_ubyte * 64)()
        self._chk(self.lib.xgpu_p2p_handle(self.h, buf))
        return bytes(buf)

    def p2p_attach(self, handles_by_rank):
        blob = b"".join(bytes(h) for h in handles_by_rank)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._chk(self.lib.xgpu_p2p_attach(self.h, buf))

    def p2p_error(self):
        return int(self.lib.xgpu_p2p_error(self.h))

    def border_set(self, n_border):
        self._chk(self.lib.xgpu_border_set(self.h, int(n_border)))

    def border_info(self):
        ni, ns, ng = C.c_int(), C.c_int(), C.c_longlong()
        self._chk(self.lib.xgpu_border_info(self.h, C.byref(ni), C.byref(ns), C.byref(ng)))
        return ni.value, ns.value, ng.value

    def shared_reduce(self, d_f, d_q, d_fl, d_ql):
        p = C.c_void_p
        self._chk(self.lib.xgpu_shared_reduce(self.h, p(d_f), p(d_q), p(d_fl), p(d_ql)))

    def border_analyze(self, d_vals):
        rc = self.lib.xgpu_border_analyze(self.h, C.c_void_p(d_vals))
        if rc not in (0, 2):
            self._chk(rc)
        return rc

    def border_solve(self, d_vals, d_rhs, d_x, rhs_border_reduced=False):
        rc = self.lib.xgpu_border_solve(self.h, C.c_void_p(d_vals), C.c_void_p(d_rhs), C.c_void_p(d_x), int(bool(rhs_border_reduced)))
        if rc not in (0, 2, 3):
            self._chk(rc)
        return rc

    def device_buffer(self, which):
        return self.lib.xgpu_device_buffer(self.h, which)

    # device-pointer hot path (pointers as integers, e.g. torch.Tensor.data_ptr())
    def update_state(self, d_sol, d_next_sta, d_curr_sta, d_next_sto, d_curr_sto, ss):
        p = C.c_void_p
        self._chk(self.lib.xgpu_update_state(self.h, p(d_sol), p(d_next_sta), p(d_curr_sta), p(d_next_sto),
                                             p(d_curr_sto), C.byref(ss)))

    def load_vectors(self, d_f, d_q, d_fl, d_ql, accumulate=False):
        p = C.c_void_p
        self._chk(self.lib.xgpu_load_vectors(self.h, p(d_f), p(d_q), p(d_fl), p(d_ql), int(accumulate)))

    def load_matrices(self, d_dfdx, d_dqdx, accumulate=False):
        p = C.c_void_p
        self._chk(self.lib.xgpu_load_matrices(self.h, p(d_dfdx), p(d_dqdx), int(accumulate)))

    def load_dae(self, d_sol, d_next_sta, d_curr_sta, d_next_sto, d_curr_sto, ss, d_f, d_q, d_fl, d_ql, d_dfdx, d_dqdx,
                 accumulate=False):
        p = C.c_void_p
        self._chk(self.lib.xgpu_load_dae(self.h, p(d_sol), p(d_next_sta), p(d_curr_sta), p(d_next_sto), p(d_curr_sto),
                                         C.byref(ss), p(d_f), p(d_q), p(d_fl), p(d_ql), p(d_dfdx), p(d_dqdx), int(accumulate)))

    def jacobian_combine(self, qs, d_dqdx, fs, d_dfdx, d_jac):
        p = C.c_void_p
        self._chk(self.lib.xgpu_jacobian_combine(self.h, qs, p(d_dqdx), fs, p(d_dfdx), p(d_jac)))

    # ---- linear devices, sources, transient ----
    def set_linear(self, g_row, g_col, g_val, c_row, c_col, c_val):
        a = [_i32(g_row), _i32(g_col), _f64(g_val), _i32(c_row), _i32(c_col), _f64(c_val)]
        self._chk(self.lib.xgpu_linear_set(self.h, len(a[0]), _ip(a[0]), _ip(a[1]), _dp(a[2]),
                                           len(a[3]), _ip(a[3]), _ip(a[4]), _dp(a[5])))

    def set_sources(self, row, scale, stype, params7):
        a = [_i32(row), _f64(scale), _i32(stype), _f64(params7)]
        self._chk(self.lib.xgpu_sources_set(self.h, len(a[0]), _ip(a[0]), _dp(a[1]), _ip(a[2]), _dp(a[3])))

    def add_linear_devices(self, kind, nodes3, value=None, stype=None, params7=None):
        """kind: "R", "C", "L" (value) or "V", "I" (source type + params); nodes3 = [n][pos, neg, branch]"""
        k = "RCLVI".index(kind)
        nd = _i32(np.asarray(nodes3, dtype=np.int32).reshape(-1, 3))
        v = None if value is None else _f64(value)
        t = None if stype is None else _i32(stype)
        p = None if params7 is None else _f64(np.asarray(params7, dtype=np.float64).reshape(-1, 7))
        self._chk(self.lib.xgpu_linear_devices_add(self.h, k, len(nd), _ip(nd), None if v is None else _dp(v),
                                                   None if t is None else _ip(t), None if p is None else _dp(p)))

    def set_pwl_table(self, tv_pairs):
        """(time, value) pairs of all PWL sources (type 5: params = {td, offset, count, repeat, repeattime})"""
        tv = _f64(np.asarray(tv_pairs, dtype=np.float64).reshape(-1, 2))
        self._chk(self.lib.xgpu_sources_pwl_set(self.h, len(tv), _dp(tv)))

    def tran_run(self, x0, tstop, tstep, probes, delmax=0.0, max_out=200000, **kw):
        tp = TranParams(tstop=tstop, tstep=tstep, delmax=delmax, **kw)
        probes = _i32(probes)
        x0 = _f64(x0)
        times, wave = np.zeros(max_out), np.zeros((max_out, len(probes)))
        steps, stats = np.zeros((max_out, 5)), np.zeros(16)
        n_out, n_steps = C.c_int(), C.c_int()
        rc = self.lib.xgpu_tran_run(self.h, C.byref(tp), _dp(x0), len(probes), _ip(probes), max_out, C.byref(n_out),
                                    _dp(times), _dp(wave), max_out, C.byref(n_steps), _dp(steps), _dp(stats))
        keys = ("accepted", "rejected", "newton_iters", "jacobian_loads", "residual_loads", "linear_solves",
                "lu_analyses", "lu_refactors", "time_points", "attempts", "driver_rc", "dcop_newton_iters", "dcop_status",
                "setup_s", "run_s", "max_readback_wait_s")
        return dict(rc=rc, t=times[:n_out.value], wave=wave[:n_out.value], steps=steps[:n_steps.value],
                    stats=dict(zip(keys, stats.tolist())),
                    error=self.lib.xgpu_last_error(self.h).decode() if rc else "")

    # ---- sparse LU ----
    def lu_analyze(self, d_vals):
        rc = self.lib.xgpu_lu_analyze(self.h, C.c_void_p(d_vals))
        if rc not in (0, 2):
            self._chk(rc)
        return rc

    def b4_lead_set(self, group, branch_lid0):
        b = _i32(branch_lid0)
        self._chk(self.lib.xgpu_b4_lead_set(self.h, int(group), _ip(b)))

    def simple_lead_set(self, group, branch_lid0):
        b = _i32(branch_lid0)
        self._chk(self.lib.xgpu_simple_lead_set(self.h, int(group), _ip(b)))

    def lead_load(self, d_sol, d_leadF, d_leadQ, d_junctionV):
        self._chk(self.lib.xgpu_lead_load(self.h, C.c_void_p(d_sol), C.c_void_p(d_leadF), C.c_void_p(d_leadQ), C.c_void_p(d_junctionV)))

    def b4_lead_load(self, d_sol, d_leadF, d_leadQ, d_junctionV):
        self._chk(self.lib.xgpu_b4_lead_load(self.h, C.c_void_p(d_sol), C.c_void_p(d_leadF), C.c_void_p(d_leadQ), C.c_void_p(d_junctionV)))

    def lu_import(self, row_perm, col_perm, block_ptr, Lp, Li, Up, Ui, row_scale=None):
        """plan from an external factorization (xgpu_lu_import); raises on malformed input"""
        a = [_i32(v) for v in (row_perm, col_perm, block_ptr, Lp, Li, Up, Ui)]
        rs = _f64(row_scale) if row_scale is not None else None
        self._chk(self.lib.xgpu_lu_import(self.h, _ip(a[0]), _ip(a[1]), len(a[2]) - 1, _ip(a[2]), _ip(a[3]), _ip(a[4]),
                                          _ip(a[5]), _ip(a[6]), _dp(rs) if rs is not None else None))

    def lu_export(self):
        sz = np.zeros(4, dtype=np.int32)
        self._chk(self.lib.xgpu_lu_export_sizes(self.h, _ip(sz)))
        n, nb, nl, nu = [int(v) for v in sz]
        out = dict(row_perm=np.zeros(n, dtype=np.int32), col_perm=np.zeros(n, dtype=np.int32),
                   block_ptr=np.zeros(nb + 1, dtype=np.int32), Lp=np.zeros(n + 1, dtype=np.int32),
                   Li=np.zeros(max(nl, 1), dtype=np.int32), Lx=np.zeros(max(nl, 1)), Up=np.zeros(n + 1, dtype=np.int32),
                   Ui=np.zeros(max(nu, 1), dtype=np.int32), Ux=np.zeros(max(nu, 1)))
        self._chk(self.lib.xgpu_lu_export(self.h, _ip(out["row_perm"]), _ip(out["col_perm"]), _ip(out["block_ptr"]),
                                          _ip(out["Lp"]), _ip(out["Li"]), _dp(out["Lx"]), _ip(out["Up"]), _ip(out["Ui"]),
                                          _dp(out["Ux"])))
        for k in ("Li", "Lx"):
            out[k] = out[k][:nl]
        for k in ("Ui", "Ux"):
            out[k] = out[k][:nu]
        return out

    def lu_refactor(self, d_vals):
        """0 ok, 2 zero / non-finite pivot, 3 a pivot of the fixed sequence fails KLU's threshold test (re-analyse)"""
        rc = self.lib.xgpu_lu_refactor(self.h, C.c_void_p(d_vals))
        if rc not in (0, 2, 3):
            self._chk(rc)
        return rc

    def lu_solve(self, d_vals, d_rhs, d_x):
        self._chk(self.lib.xgpu_lu_solve(self.h, C.c_void_p(d_vals), C.c_void_p(d_rhs), C.c_void_p(d_x)))

    def lu_info(self):
        info = np.zeros(8)
        self._chk(self.lib.xgpu_lu_info(self.h, _dp(info)))
        keys = ("n", "blocks", "largest_block", "nnz_L", "nnz_U", "offdiag", "levels", "refactor_flops")
        return dict(zip(keys, info.tolist()))

    def all_converged(self):
        v = C.c_int()
        self._chk(self.lib.xgpu_all_converged(self.h, C.byref(v)))
        return bool(v.value)

    def launch_count(self):
        return int(self.lib.xgpu_launch_count(self.h))
