// TEST INFRASTRUCTURE ONLY: compiles the single-source BSIM4 evaluator (xyce_b200/csrc/*.h)
// for the host so that CPU-only CI can check its arithmetic against the reference-built
// oracle (oracle/_ref) without a GPU.  Never loaded by the product; the product path is
// the CUDA library and fails loudly when that is missing.
#if defined(XB_TAINT_TRACE)
#include <execinfo.h>
#include <dlfcn.h>
#include <map>
#include <array>
#endif
#include <cstring>
#ifdef XB_COUNT_OPS
#include "../../xyce_b200/csrc/xb_real.h"
#define XB_REAL xb::CountReal
#endif
#ifdef XB_TAINT
#include "../../xyce_b200/csrc/xb_real.h"
#define XB_REAL xb::TaintReal
#endif
#include "../../xyce_b200/csrc/bsim4_instance.h"
#include "../../xyce_b200/csrc/diode_eval.h"
#include <string>
#include "../../xyce_b200/csrc/mos1_eval.h"
#include "../../xyce_b200/csrc/bjt_eval.h"
#include "../../xyce_b200/csrc/adms_mvs_eval.h"
#if __has_include("../../xyce_b200/csrc/gen_adms/registry.h") && !defined(XB_TAINT)
#include "../../xyce_b200/csrc/gen_adms/registry.h"      // evaluators written by the ADMS translator
#define XB_HAVE_ADMS_GEN 1
#endif

using namespace xb;
using namespace xb::b4;

namespace {
struct GeneralEmitter {
  double F[kNumRows], Q[kNumRows], FL[kNumRows], QL[kNumRows], JF[kNumSlots], JQ[kNumSlots];
  template <int R> void f(real v) { F[R] += to_double(v); }
  template <int R> void q(real v) { Q[R] += to_double(v); }
  template <int R> void fl(real v) { FL[R] += to_double(v); }
  template <int R> void ql(real v) { QL[R] += to_double(v); }
  template <int S> void jf(real v) { JF[S] += to_double(v); }
  template <int S> void jq(real v) { JQ[S] += to_double(v); }
};
}  // namespace

extern "C" {

const char *xbh_b4_mid_names(int which) {
#define NM(n) #n "\n"
  return which == 0 ? XB_B4_MID_D(NM) XB_B4_MID_EXTRA_D(NM) : XB_B4_MID_I(NM);
#undef NM
}
int xbh_b4_mid_count(int which) {
#define CNT(n) +1
  return which == 0 ? (0 XB_B4_MID_D(CNT) XB_B4_MID_EXTRA_D(CNT)) : (0 XB_B4_MID_I(CNT));
#undef CNT
}

int xbh_b4_eval(const double *model_d, const int *model_i, const double *size_d, const double *inst_d,
                const int *inst_i, const int *fl, const double *fd, const double *V12, const double *sto_old13,
                int have_old, double von_prev, double *F, double *Q, double *FL, double *QL, double *JF,
                double *JQ, double *store22, double *state3, double *mid_d, int *mid_i) {
  B4Model M; B4Size P; B4Inst I; SolverFlags S;
  int k;
#define GET(n) M.n = model_d[k++];
  k = 0; XB_B4_MODEL_D(GET)
#undef GET
#define GET(n) M.n = model_i[k++];
  k = 0; XB_B4_MODEL_I(GET)
#undef GET
#define GET(n) P.n = size_d[k++];
  k = 0; XB_B4_SIZE_D(GET)
#undef GET
#define GET(n) I.n = inst_d[k++];
  k = 0; XB_B4_INST_D(GET)
#undef GET
#define GET(n) I.n = inst_i[k++];
  k = 0; XB_B4_INST_I(GET)
#undef GET
  std::memset(&S, 0, sizeof(S));
  S.dcopFlag = fl[0]; S.tranopFlag = fl[1]; S.acopFlag = fl[2]; S.transientFlag = fl[3]; S.dcsweepFlag = fl[4];
  S.initJctFlag = fl[5]; S.initFixFlag = fl[6]; S.initTranFlag = fl[7]; S.newtonIter = fl[8];
  S.locaEnabledFlag = fl[9]; S.artParameterFlag = fl[10]; S.voltageLimiterFlag = fl[11];
  S.gmin = fd[0]; S.gainScale = fd[1]; S.nltermScale = fd[2]; S.vgstConst = 4.5; S.vdsScaleMin = 0.3;
  B4Mid W;
  std::memset(&W, 0, sizeof(W));
  GeneralEmitter e;
  std::memset(&e, 0, sizeof(e));
  real Vr[kNumNodes], so[13];
#ifdef XB_TAINT
  for (int t = 0; t < kNumNodes; ++t) Vr[t] = real(V12[t], true);      // the bias point: node voltages,
  for (int t = 0; t < 13; ++t) so[t] = real(sto_old13[t], true);       // previous limiting voltages,
  const real von_in(von_prev, true);                                   // carried threshold
  xb::taint_counts() = xb::TaintCounts{};
#else
  for (int t = 0; t < kNumNodes; ++t) Vr[t] = V12[t];
  for (int t = 0; t < 13; ++t) so[t] = sto_old13[t];
  const real von_in(von_prev);
#endif
#ifdef XB_COUNT_OPS
  xb::op_counts() = xb::OpCounts{};
#endif
  evaluate(S, M, P, I, Vr, so, have_old != 0, von_in, W, e);
  std::memcpy(F, e.F, sizeof(e.F)); std::memcpy(Q, e.Q, sizeof(e.Q));
  std::memcpy(FL, e.FL, sizeof(e.FL)); std::memcpy(QL, e.QL, sizeof(e.QL));
  std::memcpy(JF, e.JF, sizeof(e.JF)); std::memcpy(JQ, e.JQ, sizeof(e.JQ));
  for_each_store(W, [&](int s, real v) { store22[s] = to_double(v); });
  state3[sa_qb] = to_double(W.qb); state3[sa_qg] = to_double(W.qg); state3[sa_qd] = to_double(W.qd);
  k = 0;
#define PUT(n) mid_d[k++] = to_double(W.n);
  XB_B4_MID_D(PUT) XB_B4_MID_EXTRA_D(PUT)
#undef PUT
  k = 0;
#define PUT(n) mid_i[k++] = W.n;
  XB_B4_MID_I(PUT)
#undef PUT
  return 0;
}

// executed-operation counters of the last xbh_b4_eval call: add, mul, div, sqrt, exp, log, cmp
void xbh_op_counts(unsigned long long *out) {
#ifdef XB_COUNT_OPS
  const xb::OpCounts &c = xb::op_counts();
  out[0] = c.add; out[1] = c.mul; out[2] = c.div; out[3] = c.sqrt_; out[4] = c.exp_; out[5] = c.log_; out[6] = c.cmp;
#else
  for (int i = 0; i < 7; ++i) out[i] = 0;
#endif
}

#if defined(XB_TAINT_TRACE)
// call stacks of the bias-independent operations (tool build only: -O0 -g; scripts/hoistable_lines.py resolves them)
static std::map<std::array<void *, 9>, unsigned long long> g_taint_stacks;
void xb_taint_event(int kind) {
  void *fr[12];
  const int n = backtrace(fr, 12);
  std::array<void *, 9> key{};
  key[0] = (void *)(long)kind;
  for (int i = 1; i < n && i <= 8; ++i) key[i] = fr[i];
  ++g_taint_stacks[key];
}
int xbh_taint_dump(const char *path) {
  FILE *f = std::fopen(path, "w");
  if (!f) return 1;
  Dl_info info;
  dladdr((void *)&xbh_taint_dump, &info);
  std::fprintf(f, "base %p\n", info.dli_fbase);
  for (auto &e : g_taint_stacks) {
    std::fprintf(f, "%llu %ld", e.second, (long)e.first[0]);
    for (int i = 1; i < 9; ++i) std::fprintf(f, " %p", e.first[i]);
    std::fprintf(f, "\n");
  }
  std::fclose(f);
  g_taint_stacks.clear();
  return 0;
}
#endif

// TaintReal build: operations of the last xbh_b4_eval call, [bias-independent | bias-dependent] x [add, mul, div, sqrt,
// exp, log], then the divisions of a bias-dependent value by a bias-independent divisor
void xbh_taint_counts(unsigned long long *out13) {
#ifdef XB_TAINT
  const xb::TaintCounts &c = xb::taint_counts();
  for (int b = 0; b < 2; ++b) for (int k = 0; k < 6; ++k) out13[6 * b + k] = c.ops[b][k];
  out13[12] = c.div_const_divisor;
#else
  for (int i = 0; i < 13; ++i) out13[i] = 0;
#endif
}

static void fill_flags(SolverFlags &S, const int *fl, const double *fd) {
  std::memset(&S, 0, sizeof(S));
  S.dcopFlag = fl[0]; S.tranopFlag = fl[1]; S.acopFlag = fl[2]; S.transientFlag = fl[3]; S.dcsweepFlag = fl[4];
  S.initJctFlag = fl[5]; S.initFixFlag = fl[6]; S.initTranFlag = fl[7]; S.newtonIter = fl[8];
  S.locaEnabledFlag = fl[9]; S.artParameterFlag = fl[10]; S.voltageLimiterFlag = fl[11];
  S.gmin = fd[0]; S.gainScale = fd[1]; S.nltermScale = fd[2]; S.vgstConst = 4.5; S.vdsScaleMin = 0.3;
}

// diode: out = F[3] Q[3] FL[3] QL[3] JF[7] JQ[7] store[3] origFlag  (30 doubles)
int xbh_diode_eval(const double *rec, int flags, const int *fl, const double *fd, const double *V3, double vd_curr,
                   double vd_next, double *out) {
  SolverFlags S; fill_flags(S, fl, fd);
  xb::diode::Rec D;
  int k = 0;
#define GET(n) D.n = rec[k++];
  XB_DIODE_FIELDS(GET, GET)
#undef GET
  real V[3] = {V3[0], V3[1], V3[2]};
  xb::diode::Out o;
  xb::diode::evaluate(S, D, flags, V, real(vd_curr), real(vd_next), o);
  k = 0;
  for (int i = 0; i < 3; ++i) out[k++] = to_double(o.F[i]);
  for (int i = 0; i < 3; ++i) out[k++] = to_double(o.Q[i]);
  for (int i = 0; i < 3; ++i) out[k++] = to_double(o.FL[i]);
  for (int i = 0; i < 3; ++i) out[k++] = to_double(o.QL[i]);
  for (int i = 0; i < 7; ++i) out[k++] = to_double(o.JF[i]);
  for (int i = 0; i < 7; ++i) out[k++] = to_double(o.JQ[i]);
  out[k++] = to_double(o.Vd); out[k++] = to_double(o.Qd); out[k++] = to_double(o.Cd); out[k++] = o.origFlag;
  return k;
}

// BJT with excess phase (model PTF != 0): step3 = {currTimeStep, lastTimeStep, beginIntegrationFlag}, cex2 = {current, last}
// store entry CEXBC; out as xbh_simple_eval, cex_out = {mode bits, next, initial history value}
int xbh_bjt_eval_xp(const double *rec, int flags, const int *fl, const double *fd, const double *step3, const double *Vn,
                    const double *curr_sto, const double *next_sto, const double *cex2, double *out, double *cex_out) {
  SolverFlags S; fill_flags(S, fl, fd);
  S.currTimeStep = step3[0]; S.lastTimeStep = step3[1]; S.beginIntegrationFlag = (int)step3[2];
  namespace D = xb::bjt;
  D::Rec R; int j = 0, k = 0;
#define GET(n) R.n = rec[j++];
  XB_BJT_FIELDS(GET, GET)
#undef GET
  real V[D::kNodes], cs[3], ns[3];
  for (int i = 0; i < D::kNodes; ++i) V[i] = Vn[i];
  for (int i = 0; i < 3; ++i) { cs[i] = curr_sto[i]; ns[i] = next_sto[i]; }
  D::Out o;
  D::evaluate(S, R, flags, V, cs, ns, o, real(cex2[0]), real(cex2[1]));
  for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.F[i]);
  for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.Q[i]);
  for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.FL[i]);
  for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.QL[i]);
  for (int i = 0; i < D::kSlots; ++i) out[k++] = to_double(o.JF[i]);
  for (int i = 0; i < D::kSlots; ++i) out[k++] = to_double(o.JQ[i]);
  for (int i = 0; i < 3; ++i) out[k++] = to_double(o.store[i]);
  for (int i = 0; i < D::kNumState; ++i) out[k++] = to_double(o.state[i]);
  out[k++] = o.origFlag;
  cex_out[0] = o.cexbc_mode; cex_out[1] = to_double(o.cexbc_next); cex_out[2] = to_double(o.cexbc_init);
  return k;
}

// MOSFET level 1 (type 2) / BJT (type 3):
// out = F[n] Q[n] FL[n] QL[n] JF[s] JQ[s] store[..] state[..] origFlag; returns the number of doubles written
int xbh_simple_eval(int type, const double *rec, int flags, const int *fl, const double *fd, const double *Vn,
                    const double *curr_sto, const double *next_sto, const double *curr_sta, double *out) {
  SolverFlags S; fill_flags(S, fl, fd);
  int k = 0;
  if (type == 2) {
    namespace D = xb::mos1;
    D::Rec R; int j = 0;
#define GET(n) R.n = rec[j++];
    XB_MOS1_FIELDS(GET, GET)
#undef GET
    real V[D::kNodes], cs[D::kNumStore], ns[D::kNumStore], ca[D::kNumState];
    for (int i = 0; i < D::kNodes; ++i) V[i] = Vn[i];
    for (int i = 0; i < D::kNumStore; ++i) { cs[i] = curr_sto[i]; ns[i] = next_sto[i]; }
    for (int i = 0; i < D::kNumState; ++i) ca[i] = curr_sta[i];
    D::Out o;
    D::evaluate(S, R, flags, V, cs, ns, ca, o);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.F[i]);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.Q[i]);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.FL[i]);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.QL[i]);
    for (int i = 0; i < D::kSlots; ++i) out[k++] = to_double(o.JF[i]);
    for (int i = 0; i < D::kSlots; ++i) out[k++] = to_double(o.JQ[i]);
    for (int i = 0; i < D::kNumStore; ++i) out[k++] = to_double(o.store[i]);
    for (int i = 0; i < D::kNumState; ++i) out[k++] = to_double(o.state[i]);
    out[k++] = o.converged;
  } else if (type == 3) {
    namespace D = xb::bjt;
    D::Rec R; int j = 0;
#define GET(n) R.n = rec[j++];
    XB_BJT_FIELDS(GET, GET)
#undef GET
    real V[D::kNodes], cs[3], ns[3];
    for (int i = 0; i < D::kNodes; ++i) V[i] = Vn[i];
    for (int i = 0; i < 3; ++i) { cs[i] = curr_sto[i]; ns[i] = next_sto[i]; }
    D::Out o;
    D::evaluate(S, R, flags, V, cs, ns, o);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.F[i]);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.Q[i]);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.FL[i]);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.QL[i]);
    for (int i = 0; i < D::kSlots; ++i) out[k++] = to_double(o.JF[i]);
    for (int i = 0; i < D::kSlots; ++i) out[k++] = to_double(o.JQ[i]);
    for (int i = 0; i < 3; ++i) out[k++] = to_double(o.store[i]);
    for (int i = 0; i < D::kNumState; ++i) out[k++] = to_double(o.state[i]);
    out[k++] = o.origFlag;
  } else if (type == 5) {      // ADMS-generated MVS 2.0.0 ETSOI: static contributions only, no store / state, no limiting
    namespace D = xb::adms::mvs;
    D::Rec R; int j = 0;
#define GET(n) R.n = rec[j++];
    XB_MVS_FIELDS(GET)
#undef GET
    real V[D::kNodes];
    for (int i = 0; i < D::kNodes; ++i) V[i] = Vn[i];
    D::Out o;
    D::evaluate(R, V, o);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.F[i]);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.Q[i]);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.FL[i]);
    for (int i = 0; i < D::kNodes; ++i) out[k++] = to_double(o.QL[i]);
    for (int i = 0; i < D::kSlots; ++i) out[k++] = to_double(o.JF[i]);
    for (int i = 0; i < D::kSlots; ++i) out[k++] = to_double(o.JQ[i]);
    out[k++] = 1;
  }
  return k;
}

// One instance of a translated ADMS model through the host build of its generated evaluator.
// out = F, Q (nodes each), JF, JQ (slots each), store (output variables); returns the number of values or -1 for an unknown model.
int xbh_adms_gen_eval2(const char *name, const double *rec, const double *Vn, const int *fl, const double *fd, const double *curr_sto,
                       const double *next_sto, double *out);
int xbh_adms_gen_eval(const char *name, const double *rec, const double *Vn, double gmin, double *out) {
  const int fl[12] = {0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 1};
  const double fd[3] = {gmin, 1.0, 1.0};
  return xbh_adms_gen_eval2(name, rec, Vn, fl, fd, nullptr, nullptr, out);
}
// the same with the solver flags (b4_common FLAG order) and the store vectors of the instance (limited probes);
// out additionally ends with FL, QL (nodes each) and origFlag
int xbh_adms_gen_eval2(const char *name, const double *rec, const double *Vn, const int *fl, const double *fd, const double *curr_sto,
                       const double *next_sto, double *out) {
#ifdef XB_HAVE_ADMS_GEN
  const std::string nm(name);
#define XB_GEN_EVAL(i, nm_) if (nm == #nm_) { typedef xb::adms::gen_##nm_::Traits T; T::Rec R; T::Out o; real V[T::kNodes]; \
    for (int k = 0; k < T::kNumFields; ++k) R.f[k] = rec[k]; for (int k = 0; k < T::kNodes; ++k) V[k] = Vn[k]; \
    SolverFlags S; fill_flags(S, fl, fd); real cs[T::kNumStore + 1], ns[T::kNumStore + 1]; \
    for (int t = 0; t < T::kNumStore; ++t) { cs[t] = curr_sto ? curr_sto[t] : 0.0; ns[t] = next_sto ? next_sto[t] : 0.0; } \
    T::eval(S, R, V, o, cs, ns); int k = 0; \
    for (int r = 0; r < T::kNodes; ++r) out[k++] = to_double(o.F[r]); for (int r = 0; r < T::kNodes; ++r) out[k++] = to_double(o.Q[r]); \
    for (int s = 0; s < T::kSlots; ++s) out[k++] = to_double(o.JF[s]); for (int s = 0; s < T::kSlots; ++s) out[k++] = to_double(o.JQ[s]); \
    for (int s = 0; s < T::kNumStore; ++s) out[k++] = to_double(o.store[s]); \
    if (curr_sto) { for (int r = 0; r < T::kNodes; ++r) out[k++] = to_double(o.FL[r]); for (int r = 0; r < T::kNodes; ++r) out[k++] = to_double(o.QL[r]); out[k++] = o.origFlag; } \
    return k; }
  XB_ADMS_GEN_LIST(XB_GEN_EVAL)
#undef XB_GEN_EVAL
#endif
  (void)name; (void)rec; (void)Vn; (void)fl; (void)fd; (void)curr_sto; (void)next_sto; (void)out;
  return -1;
}

void xbh_b4_slot_tables(int *row, int *col) {
  for (int s = 0; s < kNumSlots; ++s) { row[s] = kSlotRow[s]; col[s] = kSlotCol[s]; }
}

}  // extern "C"
