// xyce_b200 -- CUDA backend of the Newton / transient driver (tran_driver.h) and its C ABI:
// linear-device replay (FilteredMatrix semantics), independent sources, device-resident vectors,
// norms, the KLU-pattern LU, and the time loop itself.  All vectors stay in HBM for the whole run;
// per Newton iteration the host receives only a handful of scalars (norms, convergence flag).
#include "pdl.cuh"
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>

#include "ctx.h"
#include "tran_driver.h"
#include "vecops.cuh"

using namespace xb;

namespace {

#define XS_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return xg_fail(ctx, 100 + (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

template <class T> cudaError_t up(T **d, const std::vector<T> &v) {
  cudaFree(*d); *d = nullptr;
  cudaError_t e = cudaMalloc((void **)d, (v.empty() ? 1 : v.size()) * sizeof(T));
  if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

// merge duplicate (row, col) pairs, drop ground (-1) entries
void merge_coo(int n, const int32_t *r, const int32_t *c, const double *v, XgLinearPart &L) {
  std::map<std::pair<int, int>, double> m;
  for (int k = 0; k < n; ++k) if (r[k] >= 0 && c[k] >= 0) m[std::make_pair(r[k], c[k])] += v[k];
  L.h_row.clear(); L.h_col.clear(); L.h_val.clear();
  for (auto &e : m) { L.h_row.push_back(e.first.first); L.h_col.push_back(e.first.second); L.h_val.push_back(e.second); }
}

// the same, added to the entries already there
void append_coo(int n, const int32_t *r, const int32_t *c, const double *v, XgLinearPart &L) {
  std::vector<int32_t> rr(L.h_row), cc(L.h_col); std::vector<double> vv(L.h_val);
  rr.insert(rr.end(), r, r + n); cc.insert(cc.end(), c, c + n); vv.insert(vv.end(), v, v + n);
  merge_coo((int)vv.size(), rr.data(), cc.data(), vv.data(), L);
}

__global__ void and_flags_kernel(const int *flags, int n, int *out) {
  xb::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i] == 0) *out = 0;      // every writer stores the same value
}

__global__ void gather_kernel(const double *x, const int *idx, int n, double *out) {
  xb::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[idx[i]];
}

__global__ void set_sources_kernel(double *b, const int *rows, const double *vals, int ns) {
  xb::pdl_wait();
  // one thread: sources may share a row; order is the source order (deterministic)
  if (blockIdx.x == 0 && threadIdx.x == 0) for (int k = 0; k < ns; ++k) b[rows[k]] += vals[k];
}
// up to 24 source values travel in the kernel parameter block: no staging copy, no host synchronisation
struct SrcVals { int n; int rows[24]; double vals[24]; };
__global__ void set_sources_byval_kernel(double *b, SrcVals sv) {
  xb::pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) for (int k = 0; k < sv.n; ++k) b[sv.rows[k]] += sv.vals[k];
}

using xb::sim::pulse_value;

struct GpuBackend {
  xgpu_ctx *ctx;
  cudaStream_t s;
  int n_;
  std::vector<double *> v;          // kNumVec device vectors
  double *dFdx = nullptr, *dQdx = nullptr, *J = nullptr, *scratch = nullptr;
  double *sta[2] = {nullptr, nullptr}, *sto[3] = {nullptr, nullptr, nullptr};      // sto: next, current, last
  int *d_src_rows = nullptr; double *d_src_vals = nullptr;
  int *d_probe = nullptr; double *d_probe_out = nullptr;
  std::vector<int> probes;
  std::vector<double> times, wave;
  bool lu_analyzed = false;
  int rc_alloc = 0;
  xgpu_solver_state ss{};
  long long lu_refactors = 0, lu_analyses = 0;

  int n() const { return n_; }
  void copy(int d, int a) { cudaMemcpyAsync(v[d], v[a], n_ * sizeof(double), cudaMemcpyDeviceToDevice, s); }
  void fill(int d, double val) { vec::fill(v[d], val, n_, s); ++ctx->launches; }
  void scale(int d, double a) { vec::axpby(v[d], a, v[d], 0.0, v[d], n_, s); ++ctx->launches; }
  void axpby(int d, double a, int x, double b, int y) { vec::axpby(v[d], a, v[x], b, v[y], n_, s); ++ctx->launches; }
  void axpy(int d, double a, int x) { vec::axpby(v[d], 1.0, v[d], a, v[x], n_, s); ++ctx->launches; }
  // Reductions over ALL unknowns of the circuit.  One GPU: over the local vector.  Several ranks: every rank reduces
  // its interior part and (identically) the replicated border part; the interior parts travel in one small all-gather
  // and are combined in rank order, so every rank computes bitwise the same number and takes the same decisions.
  bool multi() const { return xg_dist_multi(ctx); }
  double dreduce(vec::Reduce mode, const double *x, const double *w) {
    ctx->launches += 2;
    if (!multi()) return vec::reduce(mode, x, w, n_, scratch, s);
    const XgDist *d = ctx->dist;
    vec::reduce_dev(mode, x, w, d->ni, scratch, s);
    vec::reduce_dev(mode, x + d->ni, w ? w + d->ni : nullptr, d->ns, scratch + 2048, s);
    ctx->launches += 2;
    xg_dist_allgather(ctx, scratch, 1);
    cudaMemcpyAsync(h_norms + 4, scratch + 2048, sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    const bool is_max = (mode == vec::kMaxAbs || mode == vec::kWMaxAbs);
    double acc = h_norms[4];
    for (int r = 0; r < d->world; ++r) acc = is_max ? ((acc != acc || d->h_pack[r] != d->h_pack[r]) ? acc + d->h_pack[r] : std::max(acc, d->h_pack[r])) : acc + d->h_pack[r];
    return acc;
  }
  double n_global() const { return (ctx->dist && multi()) ? (double)ctx->dist->n_global : (double)n_; }
  double norm2(int x) { return std::sqrt(dreduce(vec::kSumSq, v[x], nullptr)); }
  double norm_inf(int x) { return dreduce(vec::kMaxAbs, v[x], nullptr); }
  double wmax_norm(int x, int w) { return dreduce(vec::kWMaxAbs, v[x], v[w]); }
  double wrms_norm(int x, int w) { return std::sqrt(dreduce(vec::kWSumSq, v[x], v[w]) / n_global()); }
  void sol_weights(int d, double rel, double ab, int a, int b) { vec::sol_weights(v[d], rel, ab, v[a], v[b], n_, s); ++ctx->launches; }
  void abs_weights(int d, double rel, double ab, int a) { vec::abs_weights(v[d], rel, ab, v[a], n_, s); ++ctx->launches; }

  bool load_rhs(const sim::Flags &fl, double time) {
    ss.dcopFlag = fl.dcop; ss.tranopFlag = fl.tranop; ss.transientFlag = fl.transient; ss.initTranFlag = fl.initTran;
    ss.newtonIter = fl.newtonIter; ss.initJctFlag = fl.initJct; ss.initFixFlag = fl.initFix; ss.currTimeStep = fl.currTimeStep;
    ss.lastTimeStep = fl.lastTimeStep; ss.beginIntegrationFlag = fl.beginIntegration;
    int rc = xgpu_update_state(ctx, v[sim::vNextSol], sta[0], sta[1], sto[0], sto[1], &ss);
    rc |= xgpu_load_vectors(ctx, v[sim::vF], v[sim::vQ], v[sim::vFlim], v[sim::vQlim], 0);
    // linear devices: F += G x, Q += C x  (N_LOA_CktLoader.C:774-782)
    const XgLinearPart &G = ctx->linG, &C = ctx->linC;
    vec::spmv_add(G.nrows, G.rows, G.ptr, G.col, G.val, v[sim::vNextSol], v[sim::vF], s);
    vec::spmv_add(C.nrows, C.rows, C.ptr, C.col, C.val, v[sim::vNextSol], v[sim::vQ], s);
    ctx->launches += (G.nrows > 0) + (C.nrows > 0);
    // several ranks: the border rows hold partial sums -- add them up over the ranks (the reference's reverse export
    // with Add, N_LOA_CktLoader.C:816-829); the sources on border rows are replicated, B needs no reduction
    if (multi()) { double *vv[4] = {v[sim::vF], v[sim::vQ], v[sim::vFlim], v[sim::vQlim]}; rc |= xg_dist_reduce_border_rows(ctx, vv, 4); }
    // independent sources evaluated on the host (DeviceMgr::updateSources), B assembled on the device
    vec::fill(v[sim::vB], 0.0, n_, s); ++ctx->launches;
    const int ns = (int)ctx->sources.size();
    if (ns > 0 && ns <= 24) {
      SrcVals sv; sv.n = ns;
      for (int k = 0; k < ns; ++k) {
        const XgSource &q = ctx->sources[k];
        sv.rows[k] = q.row; sv.vals[k] = q.scale * xb::sim::source_value(q.type, q.p, time, ctx->pwl.data(), fl.bpTol);
      }
      xb::launch_pdl(set_sources_byval_kernel, dim3(1), dim3(32), 0, s, v[sim::vB], sv); ++ctx->launches;
    } else if (ns > 0) {
      std::vector<double> vals(ns);
      for (int k = 0; k < ns; ++k) {
        const XgSource &q = ctx->sources[k];
        vals[k] = q.scale * xb::sim::source_value(q.type, q.p, time, ctx->pwl.data(), fl.bpTol);
      }
      cudaMemcpyAsync(d_src_vals, vals.data(), ns * sizeof(double), cudaMemcpyHostToDevice, s);
      cudaStreamSynchronize(s);     // vals is a stack-lifetime buffer
      xb::launch_pdl(set_sources_kernel, dim3(1), dim3(32), 0, s, v[sim::vB], d_src_rows, d_src_vals, ns); ++ctx->launches;
    }
    return rc == 0;
  }

  void load_jacobian(double qscalar, double fscalar) {
    xgpu_load_matrices(ctx, dFdx, dQdx, 0);
    const XgLinearPart &G = ctx->linG, &C = ctx->linC;
    vec::scatter_add(G.nnz, G.pos, G.val, dFdx, s);
    vec::scatter_add(C.nnz, C.pos, C.val, dQdx, s);
    ctx->launches += (G.nnz > 0) + (C.nnz > 0);
    xgpu_jacobian_combine(ctx, qscalar, dQdx, fscalar, dFdx, J);
  }

  int solve() {
    int rc;
    if (ctx->dist && (ctx->dist->ns > 0 || multi())) {      // bordered form: interior LU per rank + replicated border system
      if (!ctx->dist->analyzed) {
        rc = xgpu_border_analyze(ctx, J); ++lu_analyses;
        if (rc != 0 && rc != 2) { vec::fill(v[sim::vDX], 0.0, n_, s); return rc; }
      }
      const int before = ctx->dist->reanalyses;
      rc = xg_border_solve(ctx, J, v[sim::vRHS], v[sim::vDX], multi() ? 1 : 0); ++lu_refactors;
      lu_analyses += ctx->dist->reanalyses - before;
      if (rc != 0) vec::fill(v[sim::vDX], 0.0, n_, s);
      return rc;
    }
    if (!lu_analyzed) { rc = xgpu_lu_analyze(ctx, J); lu_analyzed = (rc == 0 || rc == 2); ++lu_analyses; }
    else {
      rc = xgpu_lu_refactor(ctx, J); ++lu_refactors;
      if (rc == 2 || rc == 3) { rc = xgpu_lu_analyze(ctx, J); ++lu_analyses; }      // bad or sub-threshold pivot: re-pivot on the host
    }
    if (rc != 0) { vec::fill(v[sim::vDX], 0.0, n_, s); return rc; }
    return xgpu_lu_solve(ctx, J, v[sim::vRHS], v[sim::vDX]);
  }

  // Loader::getBreakPoints / getMaxTimeStepSize over the independent sources (host side, like DeviceMgr)
  void breakpoints(double t, std::vector<double> &out) {
    for (const XgSource &q : ctx->sources) xb::sim::source_breakpoints(q.type, q.p, ctx->pwl.data(), t, out);
  }
  double max_source_step(double t) {
    double m = 1.0e99;
    for (const XgSource &q : ctx->sources) { const double v = xb::sim::source_max_step(q.type, q.p, t); if (v > 0.0) m = std::min(m, v); }
    return m;
  }

  bool all_devices_converged() {
    int one = 1;
    cudaMemcpyAsync(ctx->d_conv, &one, sizeof(int), cudaMemcpyHostToDevice, s);
    for (auto &g : ctx->groups) { xb::launch_pdl(and_flags_kernel, dim3((g.n + 255) / 256), dim3(256), 0, s, g.d_orig, g.n, ctx->d_conv); ++ctx->launches; }
    for (auto &g : ctx->sgroups) { xb::launch_pdl(and_flags_kernel, dim3((g.n + 255) / 256), dim3(256), 0, s, g.d_orig, g.n, ctx->d_conv); ++ctx->launches; }
    int r = 1;
    cudaMemcpyAsync(&r, ctx->d_conv, sizeof(int), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    return r != 0;
  }
  void residual_and_norms(const sim::ResidualForm &f, sim::NewtonNorms &out) {
    vec::ResidualArgs a{};
    a.rhs = v[sim::vRHS]; a.q = v[sim::vQ]; a.qh0 = v[sim::vQh0]; a.qh1 = v[sim::vQh1]; a.f = v[sim::vF]; a.b = v[sim::vB];
    a.qh2 = v[sim::vQh2]; a.qlim = v[sim::vQlim]; a.flim = v[sim::vFlim]; a.dx = v[sim::vDX]; a.w = v[sim::vSolWt];
    a.form = f.form; a.inv_h = f.inv_h; a.fs = f.fs; a.qlim_coef = f.qlim_coef; a.a0 = f.a0; a.a1 = f.a1; a.a2 = f.a2;
    a.order2 = f.order2; a.limiter = f.limiter; a.n = n_;
    a.nflag_arrays = 0;
    bool overflow = false;
    for (auto &g : ctx->groups) { if (a.nflag_arrays < 8) { a.flags[a.nflag_arrays] = g.d_orig; a.flag_n[a.nflag_arrays++] = g.n; } else overflow = true; }
    for (auto &g : ctx->sgroups) { if (a.nflag_arrays < 8) { a.flags[a.nflag_arrays] = g.d_orig; a.flag_n[a.nflag_arrays++] = g.n; } else overflow = true; }
    if (multi()) {
      // interior rows (with the device flags) and the replicated border rows separately; the interior results of all
      // ranks arrive in one all-gather (norm parts + convergence flag: DeviceMgr::allDevicesConverged's reduction)
      const XgDist *d = ctx->dist;
      vec::ResidualArgs ai = a, ab = a;
      ai.n = d->ni;
      ab.n = d->ns; ab.nflag_arrays = 0;
      ab.rhs += d->ni; ab.q += d->ni; ab.qh0 += d->ni; ab.qh1 += d->ni; ab.f += d->ni; ab.b += d->ni; ab.qh2 += d->ni;
      ab.qlim += d->ni; ab.flim += d->ni; ab.dx += d->ni; ab.w += d->ni;
      vec::residual_norms(ai, scratch, scratch + 4 * 1024, s);
      vec::residual_norms(ab, scratch, scratch + 4 * 1024 + 8, s); ctx->launches += 4;
      xg_dist_allgather(ctx, scratch + 4 * 1024, 4);
      cudaMemcpyAsync(h_norms, scratch + 4 * 1024 + 8, 4 * sizeof(double), cudaMemcpyDeviceToHost, s);
      const auto w0 = std::chrono::steady_clock::now();
      cudaStreamSynchronize(s);
      max_wait_s = std::max(max_wait_s, std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count());
      double s2 = h_norms[0], mx = h_norms[1], wm = h_norms[2], ok = 1.0;
      for (int r = 0; r < d->world; ++r) {
        const double *p = d->h_pack + 4 * r;
        s2 += p[0];
        mx = (mx != mx || p[1] != p[1]) ? mx + p[1] : std::max(mx, p[1]);
        wm = (wm != wm || p[2] != p[2]) ? wm + p[2] : std::max(wm, p[2]);
        ok = std::min(ok, p[3]);
      }
      out.rhs_norm2 = std::sqrt(s2); out.rhs_norm_inf = mx; out.dx_wmax = wm; out.devices_converged = ok != 0.0;
      return;
    }
    vec::residual_norms(a, scratch, scratch + 4 * 1024, s); ctx->launches += 2;
    cudaMemcpyAsync(h_norms, scratch + 4 * 1024, 4 * sizeof(double), cudaMemcpyDeviceToHost, s);
    const auto w0 = std::chrono::steady_clock::now();
    cudaStreamSynchronize(s);
    max_wait_s = std::max(max_wait_s, std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count());
    out.rhs_norm2 = std::sqrt(h_norms[0]); out.rhs_norm_inf = h_norms[1]; out.dx_wmax = h_norms[2];
    out.devices_converged = h_norms[3] != 0.0;
    if (overflow) out.devices_converged = all_devices_converged();     // more than 8 device groups: separate pass
  }
  double *h_norms = nullptr;          // pinned
  double max_wait_s = 0.0;            // longest single wait for the per-iteration readback (diagnostics)
  bool limiter_active() const { return ss.voltageLimiterFlag != 0; }
  void accept_state() {
    cudaMemcpyAsync(sta[1], sta[0], (size_t)ctx->n_state * sizeof(double), cudaMemcpyDeviceToDevice, s);
    // DataStore::updateSolDataArrays rotates last <- current <- next; the last store is only read by the BJT excess phase
    if (ctx->needs_last_sto) cudaMemcpyAsync(sto[2], sto[1], (size_t)ctx->n_store * sizeof(double), cudaMemcpyDeviceToDevice, s);
    cudaMemcpyAsync(sto[1], sto[0], (size_t)ctx->n_store * sizeof(double), cudaMemcpyDeviceToDevice, s);
  }
  void record(double t) {
    times.push_back(t);
    const int np = (int)probes.size();
    if (np == 0) return;
    xb::launch_pdl(gather_kernel, dim3((np + 255) / 256), dim3(256), 0, s, v[sim::vNextSol], d_probe, np, d_probe_out); ++ctx->launches;
    const size_t off = wave.size();
    wave.resize(off + np);
    cudaMemcpyAsync(&wave[off], d_probe_out, np * sizeof(double), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
  }
};

int build_linear_dev(xgpu_ctx *ctx, XgLinearPart &L) {
  // CSR over non-empty rows + position of every entry in the system pattern
  std::vector<int> rows, ptr(1, 0), col, pos;
  std::vector<double> val;
  const size_t m = L.h_row.size();
  for (size_t k = 0; k < m; ++k) {
    if (rows.empty() || rows.back() != L.h_row[k]) { if (!rows.empty()) ptr.push_back((int)col.size()); rows.push_back(L.h_row[k]); }
    col.push_back(L.h_col[k]); val.push_back(L.h_val[k]);
    const int r = L.h_row[k];
    const int32_t *b = ctx->colind.data() + ctx->rowptr[r], *e = ctx->colind.data() + ctx->rowptr[r + 1];
    const int32_t *p = std::lower_bound(b, e, L.h_col[k]);
    if (p == e || *p != L.h_col[k]) return xg_fail(ctx, 14, "a linear-device stamp entry is missing from the CSR pattern");
    pos.push_back((int)(p - ctx->colind.data()));
  }
  if (!rows.empty()) ptr.push_back((int)col.size());
  L.nrows = (int)rows.size(); L.nnz = (int)col.size();
  XS_CUDA(up(&L.rows, rows)); XS_CUDA(up(&L.ptr, ptr)); XS_CUDA(up(&L.col, col)); XS_CUDA(up(&L.pos, pos)); XS_CUDA(up(&L.val, val));
  return 0;
}

}  // namespace

int xg_finalize_linear(xgpu_ctx *ctx) {
  int rc = build_linear_dev(ctx, ctx->linG);
  if (rc) return rc;
  return build_linear_dev(ctx, ctx->linC);
}

extern "C" {

int xgpu_linear_set(xgpu_ctx *ctx, int nG, const int32_t *g_row, const int32_t *g_col, const double *g_val, int nC,
                    const int32_t *c_row, const int32_t *c_col, const double *c_val) {
  if (!ctx || nG < 0 || nC < 0) return 1;
  if (ctx->finalized) return xg_fail(ctx, 5, "linear_set after finalize");
  merge_coo(nG, g_row, g_col, g_val, ctx->linG);
  merge_coo(nC, c_row, c_col, c_val, ctx->linC);
  return 0;
}

int xgpu_sources_set(xgpu_ctx *ctx, int ns, const int32_t *row, const double *scale, const int32_t *type, const double *params7) {
  if (!ctx || ns < 0) return 1;
  ctx->sources.clear();
  for (int k = 0; k < ns; ++k) {
    if (row[k] < 0) continue;
    XgSource q; q.row = row[k]; q.scale = scale[k]; q.type = type[k];
    std::memcpy(q.p, params7 + 7 * (size_t)k, 7 * sizeof(double));
    ctx->sources.push_back(q);
  }
  return 0;
}

// Linear devices by device: the constant stamps of the reference's Master loads, merged into the replayed G / C pair, and
// the independent sources appended to the source list.
int xgpu_linear_devices_add(xgpu_ctx *ctx, int kind, int n, const int32_t *nodes3, const double *value, const int32_t *src_type,
                            const double *src_params7) {
  if (!ctx || n < 0 || kind < 0 || kind > 4 || (n > 0 && !nodes3)) return 1;
  if (ctx->finalized) return xg_fail(ctx, 5, "linear_devices_add after finalize");
  if (kind <= 2 && n > 0 && !value) return 1;
  if (kind >= 3 && n > 0 && (!src_type || !src_params7)) return 1;
  std::vector<int32_t> gr, gc, cr, cc; std::vector<double> gv, cv;
  auto G = [&](int r, int c, double v) { if (r >= 0 && c >= 0) { gr.push_back(r); gc.push_back(c); gv.push_back(v); } };
  auto Cq = [&](int r, int c, double v) { if (r >= 0 && c >= 0) { cr.push_back(r); cc.push_back(c); cv.push_back(v); } };
  for (int i = 0; i < n; ++i) {
    const int p = nodes3[3 * (size_t)i], m = nodes3[3 * (size_t)i + 1], b = nodes3[3 * (size_t)i + 2];
    if (kind >= 2 && kind <= 3 && b < 0) return xg_fail(ctx, 18, "inductor / voltage source without a branch unknown");
    switch (kind) {
      case 0: {      // Resistor: i = G (vp - vn)  (N_DEV_Resistor.C Master::loadDAEMatrices: +G -G -G +G)
        if (value[i] == 0.0) return xg_fail(ctx, 18, "resistor with R = 0");
        const double g = 1.0 / value[i];
        G(p, p, g); G(p, m, -g); G(m, p, -g); G(m, m, g);
      } break;
      case 1: {      // Capacitor: q = C (vp - vn)  (N_DEV_Capacitor.C Master::loadDAEMatrices, constant C)
        const double c = value[i];
        Cq(p, p, c); Cq(p, m, -c); Cq(m, p, -c); Cq(m, m, c);
      } break;
      case 2:        // Inductor: F[p] += i, F[n] -= i, F[b] -= vp - vn, Q[b] += L i  (N_DEV_Inductor.C:880-910, :960-985)
        G(p, b, 1.0); G(m, b, -1.0); G(b, p, -1.0); G(b, m, 1.0); Cq(b, b, value[i]);
        break;
      case 3: {      // Vsrc: F[p] += i, F[n] -= i, F[b] += vp - vn, B[b] += v(t)  (N_DEV_Vsrc.C:1323-1420, :1454-1457)
        G(p, b, 1.0); G(m, b, -1.0); G(b, p, 1.0); G(b, m, -1.0);
        XgSource q; q.row = b; q.scale = 1.0; q.type = src_type[i];
        std::memcpy(q.p, src_params7 + 7 * (size_t)i, 7 * sizeof(double));
        ctx->sources.push_back(q);
      } break;
      case 4:        // ISRC: B[p] -= i(t), B[n] += i(t)  (N_DEV_ISRC.C:1100-1130)
        for (int e = 0; e < 2; ++e) {
          const int row = e ? m : p;
          if (row < 0) continue;
          XgSource q; q.row = row; q.scale = e ? 1.0 : -1.0; q.type = src_type[i];
          std::memcpy(q.p, src_params7 + 7 * (size_t)i, 7 * sizeof(double));
          ctx->sources.push_back(q);
        }
        break;
    }
  }
  append_coo((int)gv.size(), gr.data(), gc.data(), gv.data(), ctx->linG);
  append_coo((int)cv.size(), cr.data(), cc.data(), cv.data(), ctx->linC);
  return 0;
}

// host-only views of the source routines the transient driver uses (no GPU, no context): CPU-only CI pins them on
// values worked from the reference's formulas
double xgpu_source_value(int type, const double *params7, const double *pwl_tv_pairs, double t, double bp_tol) {
  return xb::sim::source_value(type, params7, t, pwl_tv_pairs, bp_tol);
}
int xgpu_source_breakpoints(int type, const double *params7, const double *pwl_tv_pairs, double t, int max_out, double *out) {
  std::vector<double> v;
  xb::sim::source_breakpoints(type, params7, pwl_tv_pairs, t, v);
  for (int i = 0; i < (int)v.size() && i < max_out; ++i) out[i] = v[i];
  return (int)v.size();
}
double xgpu_source_max_step(int type, const double *params7, double t) { return xb::sim::source_max_step(type, params7, t); }

int xgpu_sources_pwl_set(xgpu_ctx *ctx, int n_points, const double *tv_pairs) {
  if (!ctx || n_points < 0 || (n_points > 0 && !tv_pairs)) return 1;
  ctx->pwl.assign(tv_pairs, tv_pairs + 2 * (size_t)n_points);
  return 0;
}

int xgpu_tran_run(xgpu_ctx *ctx, const xgpu_tran_params *tp, const double *h_x0, int n_probes, const int32_t *probes,
                  int max_out, int *n_out, double *h_times, double *h_wave, int max_steps_out, int *n_steps_out,
                  double *h_step_info, double *stats16) {
  if (!ctx || !tp || !h_x0 || !n_out || !n_steps_out) return 1;
  if (!ctx->finalized) return xg_fail(ctx, 15, "xgpu_finalize has not been called");
  XS_CUDA(cudaSetDevice(ctx->device));
  const auto t_begin = std::chrono::steady_clock::now();
  GpuBackend B;
  B.ctx = ctx; B.s = ctx->stream; B.n_ = ctx->n;
  B.lu_analyzed = ctx->lu_ready;      // a plan from an earlier run is reused (refactor; re-analysed on a bad pivot)
  const size_t n = (size_t)ctx->n;
  B.v.assign(sim::kNumVec, nullptr);
  // one device pool (vectors, matrices, reduction scratch, source values, probe outputs) and one int arena
  // (source rows, probe indices), both kept in the context between runs
  std::vector<int> srows; for (auto &q : ctx->sources) srows.push_back(q.row);
  const size_t need_d = sim::kNumVec * n + 3 * (size_t)ctx->nnz + 8192 + srows.size() + 1 + (size_t)n_probes + 1;
  const size_t need_i = srows.size() + (size_t)n_probes + 2;
  if (ctx->tran_pool_len < need_d) {
    cudaFree(ctx->tran_pool); ctx->tran_pool = nullptr; ctx->tran_pool_len = 0;
    XS_CUDA(cudaMalloc((void **)&ctx->tran_pool, need_d * sizeof(double)));
    ctx->tran_pool_len = need_d;
  }
  if (ctx->tran_ints_len < need_i) {
    cudaFree(ctx->tran_ints); ctx->tran_ints = nullptr; ctx->tran_ints_len = 0;
    XS_CUDA(cudaMalloc((void **)&ctx->tran_ints, need_i * sizeof(int)));
    ctx->tran_ints_len = need_i;
  }
  if (!ctx->tran_pinned) XS_CUDA(cudaMallocHost((void **)&ctx->tran_pinned, 8 * sizeof(double)));
  double *pool = ctx->tran_pool;
  XS_CUDA(cudaMemsetAsync(pool, 0, need_d * sizeof(double), ctx->stream));
  for (int i = 0; i < sim::kNumVec; ++i) B.v[i] = pool + i * n;
  B.dFdx = pool + sim::kNumVec * n; B.dQdx = B.dFdx + ctx->nnz; B.J = B.dQdx + ctx->nnz; B.scratch = B.J + ctx->nnz;
  B.sto[0] = ctx->buf[7]; B.sto[1] = ctx->buf[8]; B.sto[2] = ctx->buf[11]; ctx->d_last_sto = ctx->buf[11]; B.sta[0] = ctx->buf[9]; B.sta[1] = ctx->buf[10];
  B.d_src_vals = B.scratch + 8192; B.d_probe_out = B.d_src_vals + srows.size() + 1;
  B.d_src_rows = ctx->tran_ints; B.d_probe = ctx->tran_ints + srows.size() + 1;
  if (!srows.empty()) XS_CUDA(cudaMemcpyAsync(B.d_src_rows, srows.data(), srows.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  B.probes.assign(probes, probes + n_probes);
  if (n_probes > 0) XS_CUDA(cudaMemcpyAsync(B.d_probe, B.probes.data(), (size_t)n_probes * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  XS_CUDA(cudaStreamSynchronize(ctx->stream));      // srows is a local
  B.h_norms = ctx->tran_pinned;
  std::memset(&B.ss, 0, sizeof(B.ss));
  B.ss.voltageLimiterFlag = 1; B.ss.gmin = 1e-12; B.ss.gainScale = 1.0; B.ss.nltermScale = 1.0;
  B.ss.vgstConst = 4.5; B.ss.vdsScaleMin = 0.3; B.ss.sizeScale = 1.0; B.ss.transientFlag = 1;
  XS_CUDA(cudaMemcpyAsync(B.v[sim::vNextSol], h_x0, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  XS_CUDA(cudaMemcpyAsync(B.v[sim::vCurrSol], h_x0, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));

  sim::TranParams P;
  P.tstop = tp->tstop; P.tstep = tp->tstep; P.delmax = tp->delmax;
  if (tp->maxNewtonStep > 0) P.maxNewtonStep = tp->maxNewtonStep;
  if (tp->deltaXTol > 0) P.deltaXTol = tp->deltaXTol;
  if (tp->absTol > 0) P.absTol = tp->absTol;
  if (tp->relTol > 0) P.relTol = tp->relTol;
  if (tp->RHSTol > 0) P.RHSTol = tp->RHSTol;
  if (tp->relErrorTol > 0) P.relErrorTol = tp->relErrorTol;
  if (tp->absErrorTol > 0) P.absErrorTol = tp->absErrorTol;
  if (tp->maxOrder > 0) P.maxOrder = tp->maxOrder;
  if (tp->maxSteps > 0) P.maxSteps = tp->maxSteps;
  if (tp->method != 0 && tp->method != 7 && tp->method != 8) return xg_fail(ctx, 16, "tran: method must be 0 / 7 (trapezoid) or 8 (Gear)");
  if (tp->method == 8) P.method = 8;
  P.dcop = tp->dcop ? 1 : 0;
  sim::TransientDriver<GpuBackend> drv(B, P);
  XS_CUDA(cudaStreamSynchronize(ctx->stream));
  const auto t_setup = std::chrono::steady_clock::now();
  const int rc = drv.run();
  XS_CUDA(cudaStreamSynchronize(ctx->stream));
  const auto t_run = std::chrono::steady_clock::now();
  XS_CUDA(cudaGetLastError());

  const int nt = (int)B.times.size();
  *n_out = std::min(nt, max_out);
  for (int i = 0; i < *n_out; ++i) {
    if (h_times) h_times[i] = B.times[i];
    if (h_wave) for (int p = 0; p < n_probes; ++p) h_wave[(size_t)i * n_probes + p] = B.wave[(size_t)i * n_probes + p];
  }
  const int nsr = (int)drv.steps.size();
  *n_steps_out = std::min(nsr, max_steps_out);
  if (h_step_info)
    for (int i = 0; i < *n_steps_out; ++i) {
      const sim::StepRecord &r = drv.steps[i];
      double *o = h_step_info + 5 * (size_t)i;
      o[0] = r.t; o[1] = r.h; o[2] = r.newton_iters; o[3] = r.order; o[4] = r.status;
    }
  if (stats16) {
    const sim::TranStats &t = drv.stats;
    const double st[16] = {(double)t.accepted, (double)t.rejected, (double)t.newton_total, (double)t.jacobian_loads,
                           (double)t.residual_loads, (double)t.linear_solves, (double)B.lu_analyses, (double)B.lu_refactors,
                           (double)nt, (double)nsr, (double)rc, (double)t.dcop_newton, (double)t.dcop_status,
                           std::chrono::duration<double>(t_setup - t_begin).count(), std::chrono::duration<double>(t_run - t_setup).count(),
                           B.max_wait_s};
    std::memcpy(stats16, st, sizeof(st));
  }
  if (rc != 0) return xg_fail(ctx, 200 + rc, rc == 2 ? "transient: time step too small / too many failures"
                                             : (rc == 4 ? "transient: the DC operating point did not converge" : "transient: step limit reached"));
  return 0;
}

}  // extern "C"
