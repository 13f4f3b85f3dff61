// xyce_b200 -- BSIM4 group evaluation kernels (sm_100a).
//
// One thread evaluates one instance end to end (the reference's Master::updateState +
// loadDAEVectors + loadDAEMatrices for that instance, N_DEV_MOSFET_B4.C:10540-11686) and
// writes its F/Q/limiter rows and dF/dx, dQ/dx stamp values to the contribution planes.
// No atomics: the planes are reduced into the CSR system by assembly.cu in a fixed order.
//
// Memory behaviour: instance constants are structure-of-arrays (coalesced 8-byte loads,
// read exactly once), model / bin records are shared by whole warps (instances are sorted
// by bin, so these loads broadcast out of L1/L2), node voltages come through the gather map.
// The kernel is FP64-pipe bound; see DESIGN.md for the per-instance byte and flop budget.
#include "b4_kernels.cuh"

namespace xb {
namespace b4 {

namespace {

template <bool GENERAL>
struct PlaneEmitter;

// Default topology: accumulate into 4 rows / 16 slots held in registers.
template <>
struct PlaneEmitter<false> {
  double F[4], Q[4], FL[4], QL[4], JF[16], JQ[16];
  __device__ __forceinline__ PlaneEmitter() {
#pragma unroll
    for (int i = 0; i < 4; ++i) F[i] = Q[i] = FL[i] = QL[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) JF[i] = JQ[i] = 0.0;
  }
  template <int R> __device__ __forceinline__ void f(double v) { constexpr int k = default_collapse(R); F[k] += v; }
  template <int R> __device__ __forceinline__ void q(double v) { constexpr int k = default_collapse(R); Q[k] += v; }
  template <int R> __device__ __forceinline__ void fl(double v) { constexpr int k = default_collapse(R); FL[k] += v; }
  template <int R> __device__ __forceinline__ void ql(double v) { constexpr int k = default_collapse(R); QL[k] += v; }
  template <int S> __device__ __forceinline__ void jf(double v) {
    constexpr int k = 4 * default_collapse(slot_row(S)) + default_collapse(slot_col(S));
    JF[k] += v;
  }
  template <int S> __device__ __forceinline__ void jq(double v) {
    constexpr int k = 4 * default_collapse(slot_row(S)) + default_collapse(slot_col(S));
    JQ[k] += v;
  }
};

template <>
struct PlaneEmitter<true> {
  double F[kNumRows], Q[kNumRows], FL[kNumRows], QL[kNumRows], JF[kNumSlots], JQ[kNumSlots];
  __device__ __forceinline__ PlaneEmitter() {
#pragma unroll
    for (int i = 0; i < kNumRows; ++i) F[i] = Q[i] = FL[i] = QL[i] = 0.0;
#pragma unroll
    for (int i = 0; i < kNumSlots; ++i) JF[i] = JQ[i] = 0.0;
  }
  template <int R> __device__ __forceinline__ void f(double v) { F[R] += v; }
  template <int R> __device__ __forceinline__ void q(double v) { Q[R] += v; }
  template <int R> __device__ __forceinline__ void fl(double v) { FL[R] += v; }
  template <int R> __device__ __forceinline__ void ql(double v) { QL[R] += v; }
  template <int S> __device__ __forceinline__ void jf(double v) { JF[S] += v; }
  template <int S> __device__ __forceinline__ void jq(double v) { JQ[S] += v; }
};

__device__ __forceinline__ double gather(const double *__restrict__ x, int lid) {
  return lid >= 0 ? __ldg(x + lid) : 0.0;
}

template <bool GENERAL>
__global__ void __launch_bounds__(128) b4_eval_kernel(GroupDev g, LoadArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int n = g.n;

  // ---- parameter records ----
  B4Inst I;
  {
    int k = 0;
#define LD(name) I.name = __ldg(g.inst_d + (size_t)(k++) * n + i);
    XB_B4_INST_D(LD)
#undef LD
  }
  if (GENERAL) {
    unpack_topo(__ldg(g.topo + i), I);
  } else {
    I.rgateMod = 0; I.rbodyMod = 0; I.trnqsMod = 0; I.acnqsMod = 0;
    I.drainMOSFET_B4Exists = 0; I.sourceMOSFET_B4Exists = 0;
    I.OFF = (__ldg(g.topo + i) >> 6) & 1;
  }
  const B4Model &M = g.models[__ldg(g.model_idx + i)];
  const B4Size &P = g.sizes[__ldg(g.size_idx + i)];

  // ---- node voltages through the gather map ----
  double V[kNumNodes];
  if (GENERAL) {
#pragma unroll
    for (int t = 0; t < kNumNodes; ++t) V[t] = gather(a.sol, __ldg(g.lids + (size_t)t * n + i));
  } else {
    const double vd = gather(a.sol, __ldg(g.lids + 0 * (size_t)n + i));
    const double vg = gather(a.sol, __ldg(g.lids + 1 * (size_t)n + i));
    const double vs = gather(a.sol, __ldg(g.lids + 2 * (size_t)n + i));
    const double vb = gather(a.sol, __ldg(g.lids + 3 * (size_t)n + i));
    V[kD] = vd; V[kDP] = vd; V[kGE] = vg; V[kGP] = vg; V[kGM] = vg; V[kS] = vs; V[kSP] = vs;
    V[kB] = vb; V[kBP] = vb; V[kSB] = vb; V[kDB] = vb; V[kQ] = 0.0;
  }

  // ---- previous limiting voltages ----
  const int src = old_source(a.S);
  const int sto0 = __ldg(g.sto_lid0 + i);
  const int ss = g.sto_stride;
  double sto_old[13];
  if (src != kOldNone) {
    const double *sv = (src == kOldCurr) ? a.curr_sto : a.next_sto;
    if (GENERAL) {
#pragma unroll
      for (int t = 0; t < 13; ++t) sto_old[t] = sv[sto0 + (size_t)t * ss];
    } else {
      // only vbd, vbs, vgs, vds feed the limiters of a 4-terminal device
#pragma unroll
      for (int t = 0; t < 4; ++t) sto_old[t] = sv[sto0 + (size_t)t * ss];
#pragma unroll
      for (int t = 4; t < 13; ++t) sto_old[t] = 0.0;
    }
  }

  B4Mid W;
  PlaneEmitter<GENERAL> e;
  evaluate(a.S, M, P, I, V, sto_old, src != kOldNone, g.von[i], W, e);

  // ---- carried state, store and state vectors ----
  g.von[i] = W.von;
  g.orig_flag[i] = W.origFlag;
  {
    double *ns = a.next_sto;
    for_each_store(W, [&](int s, double v) { ns[sto0 + (size_t)s * ss] = v; });
  }
  {
    const int sta0 = __ldg(g.sta_lid0 + i);
    const int as = g.sta_stride;
    double *st = a.next_sta;
    st[sta0 + (size_t)sa_qb * as] = W.qb;
    st[sta0 + (size_t)sa_qg * as] = W.qg;
    st[sta0 + (size_t)sa_qd * as] = W.qd;
    if (GENERAL) {
      int k = 3;
      if (I.rgateMod == 3) st[sta0 + (size_t)(k++) * as] = W.qgmid;
      if (I.rbodyMod) { st[sta0 + (size_t)(k++) * as] = W.qbs; st[sta0 + (size_t)(k++) * as] = W.qbd; }
    }
    // first Newton step of the first transient step: charges also go to the current state
    // (N_DEV_MOSFET_B4.C:10629-10664)
    if (!a.S.dcopFlag && a.S.initTranFlag && a.S.newtonIter == 0) {
      double *cs = a.curr_sta;
      cs[sta0 + (size_t)sa_qb * as] = W.qb;
      cs[sta0 + (size_t)sa_qg * as] = W.qg;
      cs[sta0 + (size_t)sa_qd * as] = W.qd;
      if (GENERAL) {
        int k = 3;
        if (I.rgateMod == 3) cs[sta0 + (size_t)(k++) * as] = W.qgmid;
        if (I.rbodyMod) { cs[sta0 + (size_t)(k++) * as] = W.qbs; cs[sta0 + (size_t)(k++) * as] = W.qbd; }
      }
    }
  }

  // ---- contribution planes (coalesced: plane[row][instance]) ----
  constexpr int R = GENERAL ? kRowsGeneral : kRowsDefault;
  constexpr int SL = GENERAL ? kSlotsGeneral : kSlotsDefault;
  double *pf = a.vec_planes[0] + g.vec_base + i;
  double *pq = a.vec_planes[1] + g.vec_base + i;
  double *pfl = a.vec_planes[2] + g.vec_base + i;
  double *pql = a.vec_planes[3] + g.vec_base + i;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    pf[(size_t)r * n] = e.F[r];
    pq[(size_t)r * n] = e.Q[r];
    pfl[(size_t)r * n] = e.FL[r];
    pql[(size_t)r * n] = e.QL[r];
  }
  double *jf = a.mat_planes[0] + g.mat_base + i;
  double *jq = a.mat_planes[1] + g.mat_base + i;
#pragma unroll
  for (int s = 0; s < SL; ++s) {
    jf[(size_t)s * n] = e.JF[s];
    jq[(size_t)s * n] = e.JQ[s];
  }
}

}  // namespace

void launch_b4_group(const GroupDev &g, const LoadArgs &a, cudaStream_t stream) {
  if (g.n <= 0) return;
  const int threads = 128;
  const int blocks = (g.n + threads - 1) / threads;
  if (g.general) b4_eval_kernel<true><<<blocks, threads, 0, stream>>>(g, a);
  else b4_eval_kernel<false><<<blocks, threads, 0, stream>>>(g, a);
}

}  // namespace b4
}  // namespace xb
