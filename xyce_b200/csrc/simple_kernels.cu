// xyce_b200 -- kernels of the small compact models (sm_100a).  Memory-bound-ish (few hundred flops per
// instance): records are read once, coalesced; planes written coalesced.
#include "pdl.cuh"
#include "simple_kernels.cuh"
#include "diode_eval.h"
#include "adms_rlc_eval.h"
#include "adms_mvs_eval.h"
#include "bjt_eval.h"
#include "mos1_eval.h"

namespace xb {
namespace simple {

namespace {

__device__ __forceinline__ double gatherv(const double *__restrict__ x, int lid) { return lid >= 0 ? __ldg(x + lid) : 0.0; }

__global__ void __launch_bounds__(128) diode_kernel(GroupDev g, b4::LoadArgs a) {
  xb::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int n = g.n;
  diode::Rec D;
  {
    int k = 0;
#define LD(name) D.name = __ldg(g.rec + (size_t)(k++) * n + i);
    XB_DIODE_FIELDS(LD, LD)
#undef LD
  }
  real V[diode::kNodes];
#pragma unroll
  for (int t = 0; t < diode::kNodes; ++t) V[t] = gatherv(a.sol, __ldg(g.lids + (size_t)t * n + i));
  const int sto0 = __ldg(g.sto_lid0 + i), ss = g.sto_stride;
  diode::Out o;
  diode::evaluate(a.S, D, __ldg(g.flags + i), V, a.curr_sto[sto0], a.next_sto[sto0], o);
  a.next_sto[sto0] = to_double(o.Vd);
  a.next_sto[sto0 + (size_t)ss] = to_double(o.Qd);
  a.next_sto[sto0 + 2 * (size_t)ss] = to_double(o.Cd);
  g.orig_flag[i] = o.origFlag;
#pragma unroll
  for (int r = 0; r < diode::kNodes; ++r) {
    a.vec_planes[0][g.vec_base + (size_t)r * n + i] = to_double(o.F[r]);
    a.vec_planes[1][g.vec_base + (size_t)r * n + i] = to_double(o.Q[r]);
    a.vec_planes[2][g.vec_base + (size_t)r * n + i] = to_double(o.FL[r]);
    a.vec_planes[3][g.vec_base + (size_t)r * n + i] = to_double(o.QL[r]);
  }
#pragma unroll
  for (int s = 0; s < diode::kSlots; ++s) {
    a.mat_planes[0][g.mat_base + (size_t)s * n + i] = to_double(o.JF[s]);
    a.mat_planes[1][g.mat_base + (size_t)s * n + i] = to_double(o.JQ[s]);
  }
}

// shared tail of the small-device kernels: contributions -> planes (coalesced, plane[row][instance])
template <class Out, int NODES, int SLOTS>
__device__ __forceinline__ void store_planes(const GroupDev &g, const b4::LoadArgs &a, const Out &o, int i) {
  const int n = g.n;
#pragma unroll
  for (int r = 0; r < NODES; ++r) {
    a.vec_planes[0][g.vec_base + (size_t)r * n + i] = to_double(o.F[r]);
    a.vec_planes[1][g.vec_base + (size_t)r * n + i] = to_double(o.Q[r]);
    a.vec_planes[2][g.vec_base + (size_t)r * n + i] = to_double(o.FL[r]);
    a.vec_planes[3][g.vec_base + (size_t)r * n + i] = to_double(o.QL[r]);
  }
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    a.mat_planes[0][g.mat_base + (size_t)s * n + i] = to_double(o.JF[s]);
    a.mat_planes[1][g.mat_base + (size_t)s * n + i] = to_double(o.JQ[s]);
  }
}

__global__ void __launch_bounds__(128) mos1_kernel(GroupDev g, b4::LoadArgs a) {
  xb::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int n = g.n;
  namespace D = mos1;
  D::Rec R;
  {
    int k = 0;
#define LD(name) R.name = __ldg(g.rec + (size_t)(k++) * n + i);
    XB_MOS1_FIELDS(LD, LD)
#undef LD
  }
  real V[D::kNodes];
#pragma unroll
  for (int t = 0; t < D::kNodes; ++t) V[t] = gatherv(a.sol, __ldg(g.lids + (size_t)t * n + i));
  const int sto0 = __ldg(g.sto_lid0 + i), ss = g.sto_stride, sta0 = __ldg(g.sta_lid0 + i), as = g.sta_stride;
  real cs[D::kNumStore], ns[D::kNumStore], ca[D::kNumState];
#pragma unroll
  for (int t = 0; t < D::kNumStore; ++t) { cs[t] = a.curr_sto[sto0 + (size_t)t * ss]; ns[t] = a.next_sto[sto0 + (size_t)t * ss]; }
#pragma unroll
  for (int t = 0; t < D::kNumState; ++t) ca[t] = a.curr_sta[sta0 + (size_t)t * as];
  D::Out o;
  D::evaluate(a.S, R, __ldg(g.flags + i), V, cs, ns, ca, o);
#pragma unroll
  for (int t = 0; t < D::kNumStore; ++t) a.next_sto[sto0 + (size_t)t * ss] = to_double(o.store[t]);
#pragma unroll
  for (int t = 0; t < D::kNumState; ++t) a.next_sta[sta0 + (size_t)t * as] = to_double(o.state[t]);
  g.orig_flag[i] = o.converged;
  store_planes<D::Out, D::kNodes, D::kSlots>(g, a, o, i);
}

__global__ void __launch_bounds__(128) bjt_kernel(GroupDev g, b4::LoadArgs a) {
  xb::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int n = g.n;
  namespace D = bjt;
  D::Rec R;
  {
    int k = 0;
#define LD(name) R.name = __ldg(g.rec + (size_t)(k++) * n + i);
    XB_BJT_FIELDS(LD, LD)
#undef LD
  }
  real V[D::kNodes];
#pragma unroll
  for (int t = 0; t < D::kNodes; ++t) V[t] = gatherv(a.sol, __ldg(g.lids + (size_t)t * n + i));
  const int sto0 = __ldg(g.sto_lid0 + i), ss = g.sto_stride, sta0 = __ldg(g.sta_lid0 + i), as = g.sta_stride;
  real cs[3], ns[3];
#pragma unroll
  for (int t = 0; t < 3; ++t) { cs[t] = a.curr_sto[sto0 + (size_t)t * ss]; ns[t] = a.next_sto[sto0 + (size_t)t * ss]; }
  D::Out o;
  real cex_c = 0.0, cex_l = 0.0;
  const size_t cex = sto0 + (size_t)D::st_cexbc * ss;
  if (R.excessPhaseFac != 0.0 && !a.S.dcopFlag && !a.S.beginIntegrationFlag) { cex_c = a.curr_sto[cex]; cex_l = a.last_sto[cex]; }
  D::evaluate(a.S, R, __ldg(g.flags + i), V, cs, ns, o, cex_c, cex_l);
#pragma unroll
  for (int t = 0; t < 3; ++t) a.next_sto[sto0 + (size_t)t * ss] = to_double(o.store[t]);
  if (o.cexbc_mode & 1) a.next_sto[cex] = to_double(o.cexbc_next);
  if (o.cexbc_mode & 2) { a.curr_sto[cex] = to_double(o.cexbc_init); a.last_sto[cex] = to_double(o.cexbc_init); }
#pragma unroll
  for (int t = 0; t < D::kNumState; ++t) a.next_sta[sta0 + (size_t)t * as] = to_double(o.state[t]);
  // first Newton step of the first transient step: charges also go to the current state (N_DEV_BJT.C:4135-4147)
  if (!a.S.dcopFlag && a.S.initTranFlag && a.S.newtonIter == 0) {
#pragma unroll
    for (int t = 0; t < D::kNumState; ++t) a.curr_sta[sta0 + (size_t)t * as] = to_double(o.state[t]);
  }
  g.orig_flag[i] = o.origFlag;
  if (g.lead) {
#pragma unroll
    for (int t = 0; t < 4; ++t) { g.lead[(size_t)t * n + i] = to_double(o.leadF[t]); g.lead[(size_t)(4 + t) * n + i] = to_double(o.leadQ[t]); }
  }
  store_planes<D::Out, D::kNodes, D::kSlots>(g, a, o, i);
}

// Lead currents (Instance::loadLeadCurrent; what .PRINT I(D1) / IC(Q1) / P(M1) switches on).  Diode and MOSFET level 1: the
// lead quantities ARE terms the evaluation has just written to the contribution planes -- diode leadF = Id mf = -(Neg row
// term) (N_DEV_Diode.C:1889-1897); MOSFET1 leadF[id] = the Drain row term when RD != 0, else the Drain' one, and so on
// (N_DEV_MOSFET1.C:4544-4572) -- BJT: the block the evaluation kernel wrote (terminal currents are not row terms when
// RC / RB / RE != 0).  ASSIGNS, like the reference.
__global__ void __launch_bounds__(256) lead_kernel(GroupDev g, const int *__restrict__ branch0, const double *__restrict__ pF,
                                                   const double *__restrict__ pQ, const double *__restrict__ sol, double *leadF,
                                                   double *leadQ, double *junctionV, int f0, int f1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int b = branch0[i];
  if (b < 0) return;
  const size_t n = g.n;
  auto v = [&](int node) { const int l = g.lids[node * n + i]; return l >= 0 ? sol[l] : 0.0; };
  if (g.type == kDiode) {
    leadF[b] = -pF[diode::kNeg * n + i];
    if (g.rec[f0 * n + i] != 0.0) leadQ[b] = -pQ[diode::kNeg * n + i];      // model CJO != 0 (tJctCap = CJO x temperature factor)
    junctionV[b] = v(diode::kPos) - v(diode::kNeg);
  } else if (g.type == kMos1) {
    const bool rd = g.rec[f0 * n + i] != 0.0, rs = g.rec[f1 * n + i] != 0.0;      // drainConductance / sourceConductance
    leadF[b + 0] = pF[(rd ? mos1::kD : mos1::kDP) * n + i];
    if (!rd) leadQ[b + 0] = pQ[mos1::kDP * n + i];
    leadF[b + 2] = pF[(rs ? mos1::kS : mos1::kSP) * n + i];
    if (!rs) leadQ[b + 2] = pQ[mos1::kSP * n + i];
    leadF[b + 1] = pF[mos1::kG * n + i]; leadQ[b + 1] = pQ[mos1::kG * n + i];
    leadF[b + 3] = pF[mos1::kB * n + i]; leadQ[b + 3] = pQ[mos1::kB * n + i];
    junctionV[b + 0] = v(mos1::kD) - v(mos1::kS);
    junctionV[b + 1] = v(mos1::kG) - v(mos1::kS);
    junctionV[b + 2] = 0.0; junctionV[b + 3] = 0.0;
  } else if (g.type == kBjt) {
#pragma unroll
    for (int t = 0; t < 4; ++t) { leadF[b + t] = g.lead[t * n + i]; leadQ[b + t] = g.lead[(4 + t) * n + i]; }
    junctionV[b + 2] = v(bjt::kC) - v(bjt::kE);
    junctionV[b + 3] = 0.0;
    junctionV[b + 0] = v(bjt::kB) - v(bjt::kE);
    junctionV[b + 1] = 0.0;
  }
}

__global__ void __launch_bounds__(128) rlc_kernel(GroupDev g, b4::LoadArgs a) {
  xb::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int n = g.n;
  namespace R = adms::rlc;
  const double r = __ldg(g.rec + i), l = __ldg(g.rec + (size_t)n + i), c = __ldg(g.rec + 2 * (size_t)n + i);
  real V[R::kNodes];
#pragma unroll
  for (int t = 0; t < R::kNodes; ++t) V[t] = gatherv(a.sol, __ldg(g.lids + (size_t)t * n + i));
  R::Out o;
  R::evaluate(r, l, c, V, o);
  g.orig_flag[i] = 1;
#pragma unroll
  for (int t = 0; t < R::kNodes; ++t) {
    a.vec_planes[0][g.vec_base + (size_t)t * n + i] = to_double(o.F[t]);
    a.vec_planes[1][g.vec_base + (size_t)t * n + i] = to_double(o.Q[t]);
    a.vec_planes[2][g.vec_base + (size_t)t * n + i] = 0.0;
    a.vec_planes[3][g.vec_base + (size_t)t * n + i] = 0.0;
  }
#pragma unroll
  for (int s = 0; s < R::kSlots; ++s) {
    a.mat_planes[0][g.mat_base + (size_t)s * n + i] = to_double(o.JF[s]);
    a.mat_planes[1][g.mat_base + (size_t)s * n + i] = to_double(o.JQ[s]);
  }
}

// ADMS-generated MVS 2.0.0 ETSOI (N_DEV_ADMSmvs_2_0_0_etsoi.C): static contributions only, no limiting, no state
__global__ void __launch_bounds__(128) mvs_kernel(GroupDev g, b4::LoadArgs a) {
  xb::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int n = g.n;
  namespace M = adms::mvs;
  M::Rec R;
  {
    int k = 0;
#define LD(name) R.name = __ldg(g.rec + (size_t)(k++) * n + i);
    XB_MVS_FIELDS(LD)
#undef LD
  }
  real V[M::kNodes];
#pragma unroll
  for (int t = 0; t < M::kNodes; ++t) V[t] = gatherv(a.sol, __ldg(g.lids + (size_t)t * n + i));
  M::Out o;
  M::evaluate(R, V, o);
  g.orig_flag[i] = 1;
  store_planes<M::Out, M::kNodes, M::kSlots>(g, a, o, i);
}

const int kMvsRow[adms::mvs::kSlots] = {0, 0, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 6, 6, 6, 6};
const int kMvsCol[adms::mvs::kSlots] = {3, 0, 2, 4, 4, 3, 5, 0, 4, 3, 5, 2, 6, 5, 4, 1, 3};
const TypeInfo kMvsInfo = {adms::mvs::kNodes, adms::mvs::kSlots, adms::mvs::kNumFields, 0, 0, kMvsRow, kMvsCol};

const int kRlcRow[adms::rlc::kSlots] = {0, 0, 2, 2, 2, 3, 3, 3, 1, 4, 4, 4};
const int kRlcCol[adms::rlc::kSlots] = {0, 2, 0, 2, 3, 2, 3, 4, 4, 3, 1, 4};
const TypeInfo kRlcInfo = {adms::rlc::kNodes, adms::rlc::kSlots, adms::rlc::kNumFields, 0, 0, kRlcRow, kRlcCol};

const int kMos1Row[mos1::kSlots] = {0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5};
const int kMos1Col[mos1::kSlots] = {0, 4, 1, 3, 4, 5, 2, 5, 1, 3, 4, 5, 0, 1, 3, 4, 5, 1, 2, 3, 4, 5};
const TypeInfo kMos1Info = {mos1::kNodes, mos1::kSlots, mos1::kNumFields, mos1::kNumStore, mos1::kNumState, kMos1Row, kMos1Col};
const int kBjtRow[bjt::kSlots] = {0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6};
const int kBjtCol[bjt::kSlots] = {0, 4, 1, 4, 5, 6, 2, 6, 3, 4, 0, 1, 3, 4, 5, 6, 1, 4, 5, 6, 2, 4, 5, 6};
const TypeInfo kBjtInfo = {bjt::kNodes, bjt::kSlots, bjt::kNumFields, bjt::kNumStore, bjt::kNumState, kBjtRow, kBjtCol};

const int kDiodeRow[diode::kSlots] = {0, 0, 1, 1, 2, 2, 2};
const int kDiodeCol[diode::kSlots] = {0, 2, 1, 2, 0, 1, 2};
const TypeInfo kDiodeInfo = {diode::kNodes, diode::kSlots, diode::kNumFields, 3, 0, kDiodeRow, kDiodeCol};

}  // namespace

int lead_count(int type) { return type == kDiode ? 1 : (type == kMos1 || type == kBjt) ? 4 : 0; }

void launch_lead(const GroupDev &g, const int *branch0, const double *planeF, const double *planeQ, const double *sol,
                 double *leadF, double *leadQ, double *junctionV, cudaStream_t s) {
  if (g.n <= 0 || lead_count(g.type) == 0) return;
  int f0 = 0, f1 = 0;
  if (g.type == kDiode) f0 = (int)(offsetof(diode::Rec, tJctCap) / sizeof(double));
  if (g.type == kMos1) { f0 = (int)(offsetof(mos1::Rec, drainConductance) / sizeof(double)); f1 = (int)(offsetof(mos1::Rec, sourceConductance) / sizeof(double)); }
  lead_kernel<<<(g.n + 255) / 256, 256, 0, s>>>(g, branch0, planeF, planeQ, sol, leadF, leadQ, junctionV, f0, f1);
}

int bjt_excess_phase_field() { return (int)(offsetof(bjt::Rec, excessPhaseFac) / sizeof(double)); }

const TypeInfo *type_info(int type) {
  switch (type) {
    case kDiode: return &kDiodeInfo;
    case kMos1: return &kMos1Info;
    case kBjt: return &kBjtInfo;
    case kRlc: return &kRlcInfo;
    case kMvs: return &kMvsInfo;
    default: break;
  }
  return adms_gen_type_info(type);      // translated ADMS models (adms_gen_kernels.cu), nullptr when unknown
}

void launch_group(const GroupDev &g, const b4::LoadArgs &a, cudaStream_t s) {
  if (g.n <= 0) return;
  const int blocks = (g.n + 127) / 128;
  switch (g.type) {
    case kDiode: xb::launch_pdl(diode_kernel, dim3(blocks), dim3(128), 0, s, g, a); break;
    case kMos1: xb::launch_pdl(mos1_kernel, dim3(blocks), dim3(128), 0, s, g, a); break;
    case kBjt: xb::launch_pdl(bjt_kernel, dim3(blocks), dim3(128), 0, s, g, a); break;
    case kRlc: xb::launch_pdl(rlc_kernel, dim3(blocks), dim3(128), 0, s, g, a); break;
    case kMvs: xb::launch_pdl(mvs_kernel, dim3(blocks), dim3(128), 0, s, g, a); break;
    default: launch_adms_gen_group(g, a, s); break;
  }
}

}  // namespace simple
}  // namespace xb
