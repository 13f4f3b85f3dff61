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
// Compiled three times (see Makefile) with XB_ARITH =
//   0 "exact": plain double, -fmad=false  -> every multiply/add rounded separately (parity build)
//   1 "fma"  : plain double, -fmad=true   -> FMA contraction
//   2 "fast" : FastReal (branch-free reciprocal division), -fmad=true
#ifndef XB_ARITH
#define XB_ARITH 0
#endif
#if XB_ARITH == 2
#include "xb_real.h"
#define XB_REAL xb::FastReal
#endif
#include "b4_kernels.cuh"
#define XB_CAT2(a, b) a##b
#define XB_CAT(a, b) XB_CAT2(a, b)

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
  template <int R> __device__ __forceinline__ void f(real v) { constexpr int k = default_collapse(R); F[k] += to_double(v); }
  template <int R> __device__ __forceinline__ void q(real v) { constexpr int k = default_collapse(R); Q[k] += to_double(v); }
  template <int R> __device__ __forceinline__ void fl(real v) { constexpr int k = default_collapse(R); FL[k] += to_double(v); }
  template <int R> __device__ __forceinline__ void ql(real v) { constexpr int k = default_collapse(R); QL[k] += to_double(v); }
  template <int S> __device__ __forceinline__ void jf(real v) {
    constexpr int k = 4 * default_collapse(slot_row(S)) + default_collapse(slot_col(S));
    JF[k] += to_double(v);
  }
  template <int S> __device__ __forceinline__ void jq(real v) {
    constexpr int k = 4 * default_collapse(slot_row(S)) + default_collapse(slot_col(S));
    JQ[k] += to_double(v);
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
  template <int R> __device__ __forceinline__ void f(real v) { F[R] += to_double(v); }
  template <int R> __device__ __forceinline__ void q(real v) { Q[R] += to_double(v); }
  template <int R> __device__ __forceinline__ void fl(real v) { FL[R] += to_double(v); }
  template <int R> __device__ __forceinline__ void ql(real v) { QL[R] += to_double(v); }
  template <int S> __device__ __forceinline__ void jf(real v) { JF[S] += to_double(v); }
  template <int S> __device__ __forceinline__ void jq(real v) { JQ[S] += to_double(v); }
};

__device__ __forceinline__ double gather(const double *__restrict__ x, int lid) {
  return lid >= 0 ? __ldg(x + lid) : 0.0;
}

template <bool GENERAL, int MINBLOCKS>
__global__ void __launch_bounds__(128, MINBLOCKS) b4_eval_kernel(GroupDev g, LoadArgs a) {
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
  real V[kNumNodes];
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
  real sto_old[13];
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
  g.von[i] = to_double(W.von);
  g.orig_flag[i] = W.origFlag;
  {
    double *ns = a.next_sto;
    for_each_store(W, [&](int s, real v) { ns[sto0 + (size_t)s * ss] = to_double(v); });
  }
  {
    const int sta0 = __ldg(g.sta_lid0 + i);
    const int as = g.sta_stride;
    double *st = a.next_sta;
    st[sta0 + (size_t)sa_qb * as] = to_double(W.qb);
    st[sta0 + (size_t)sa_qg * as] = to_double(W.qg);
    st[sta0 + (size_t)sa_qd * as] = to_double(W.qd);
    if (GENERAL) {
      int k = 3;
      if (I.rgateMod == 3) st[sta0 + (size_t)(k++) * as] = to_double(W.qgmid);
      if (I.rbodyMod) { st[sta0 + (size_t)(k++) * as] = to_double(W.qbs); st[sta0 + (size_t)(k++) * as] = to_double(W.qbd); }
    }
    // first Newton step of the first transient step: charges also go to the current state
    // (N_DEV_MOSFET_B4.C:10629-10664)
    if (!a.S.dcopFlag && a.S.initTranFlag && a.S.newtonIter == 0) {
      double *cs = a.curr_sta;
      cs[sta0 + (size_t)sa_qb * as] = to_double(W.qb);
      cs[sta0 + (size_t)sa_qg * as] = to_double(W.qg);
      cs[sta0 + (size_t)sa_qd * as] = to_double(W.qd);
      if (GENERAL) {
        int k = 3;
        if (I.rgateMod == 3) cs[sta0 + (size_t)(k++) * as] = to_double(W.qgmid);
        if (I.rbodyMod) { cs[sta0 + (size_t)(k++) * as] = to_double(W.qbs); cs[sta0 + (size_t)(k++) * as] = to_double(W.qbd); }
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

void XB_CAT(launch_b4_group_a, XB_ARITH)(const GroupDev &g, const LoadArgs &a, int minblocks, cudaStream_t stream) {
  if (g.n <= 0) return;
  const int threads = 128;
  const int blocks = (g.n + threads - 1) / threads;
  if (g.general) {
    b4_eval_kernel<true, 2><<<blocks, threads, 0, stream>>>(g, a);
  } else {
    switch (minblocks) {
      case 3: b4_eval_kernel<false, 3><<<blocks, threads, 0, stream>>>(g, a); break;
      case 4: b4_eval_kernel<false, 4><<<blocks, threads, 0, stream>>>(g, a); break;
      case 5: b4_eval_kernel<false, 5><<<blocks, threads, 0, stream>>>(g, a); break;
      case 6: b4_eval_kernel<false, 6><<<blocks, threads, 0, stream>>>(g, a); break;
      default: b4_eval_kernel<false, 2><<<blocks, threads, 0, stream>>>(g, a); break;
    }
  }
}

}  // namespace b4
}  // namespace xb
