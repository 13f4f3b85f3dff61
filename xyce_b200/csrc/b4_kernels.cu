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

// One instance, end to end.  `valid` = false marks a padding thread of a lock-step block: it runs the
// arithmetic on the last instance's data (so every thread reaches the block barriers) and stores nothing.
template <bool GENERAL>
__device__ __forceinline__ void eval_instance(const GroupDev &g, const LoadArgs &a, const B4Model &M, const B4Size &P,
                                              const int i, const bool valid) {
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

  if (!valid) return;
  if (GENERAL) {        // lead currents of devices with internal nodes (the 4-terminal fast path gets them from the planes)
    if (g.lead) {
      real lf[4], lq[4];
      emit_lead(M, I, W, lf, lq);
#pragma unroll
      for (int k = 0; k < 4; ++k) { g.lead[(size_t)k * n + i] = to_double(lf[k]); g.lead[(size_t)(4 + k) * n + i] = to_double(lq[k]); }
    }
  }
  // ---- carried state, store and state vectors ----
  g.von[i] = to_double(W.von);
  g.orig_flag[i] = !W.limitedFlag;      // Instance::isConverged() (N_DEV_MOSFET_B4.h:2328-2331): only pnjlim invalidates convergence
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


// Variant A ("per-thread records"): every thread fetches its own model / bin record through the index
// arrays.  Works for any mix of bins inside a block; the loads broadcast out of L1 when neighbours agree.
template <bool GENERAL, int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) b4_eval_kernel(const __grid_constant__ GroupDev g,
                                                                     const __grid_constant__ LoadArgs a) {
  int i = blockIdx.x * THREADS + threadIdx.x;
#if XB_LOCKSTEP
  const bool valid = i < g.n;
  if (!valid) i = g.n - 1;
#else
  if (i >= g.n) return;
  const bool valid = true;
#endif
  const B4Model &M = g.models[__ldg(g.model_idx + i)];
  const B4Size &P = g.sizes[__ldg(g.size_idx + i)];
  eval_instance<GENERAL>(g, a, M, P, i, valid);
}

// Variant B ("uniform records"): instances are sorted by (model, bin); blockIdx.y selects one run of equal
// records whose model card and bin live in the kernel parameter block (constant bank), so every parameter
// is a uniform operand: no per-thread loads, no vector registers, and the model's mode switches
// (capMod, mobMod, igcMod, ...) become uniform branches.
template <bool GENERAL, int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) b4_eval_uniform_kernel(const __grid_constant__ GroupDev g,
                                                                             const __grid_constant__ LoadArgs a,
                                                                             const __grid_constant__ BinPack bp) {
  const BinRun &run = bp.run[blockIdx.y];
  int k = blockIdx.x * THREADS + threadIdx.x;
#if XB_LOCKSTEP
  if (blockIdx.x * THREADS >= run.count) return;       // whole block outside this run
  const bool valid = k < run.count;
  if (!valid) k = run.count - 1;
#else
  if (k >= run.count) return;
  const bool valid = true;
#endif
  eval_instance<GENERAL>(g, a, run.M, run.P, run.start + k, valid);
}

template <bool GENERAL, int THREADS, int MINBLOCKS>
void launch_variant(const GroupDev &g, const LoadArgs &a, const BinPack *packs, int npacks, cudaStream_t stream) {
  if (packs && npacks > 0) {
    for (int p = 0; p < npacks; ++p) {
      int most = 0;
      for (int r = 0; r < packs[p].nruns; ++r) most = packs[p].run[r].count > most ? packs[p].run[r].count : most;
      const dim3 grid((most + THREADS - 1) / THREADS, packs[p].nruns);
      b4_eval_uniform_kernel<GENERAL, THREADS, MINBLOCKS><<<grid, THREADS, 0, stream>>>(g, a, packs[p]);
    }
  } else {
#if !XB_LOCKSTEP && !XB_SPEC
    b4_eval_kernel<GENERAL, THREADS, MINBLOCKS><<<(g.n + THREADS - 1) / THREADS, THREADS, 0, stream>>>(g, a);
#endif
  }
}

}  // namespace

#ifndef XB_SPEC
#define XB_SPEC 0
#endif
#ifndef XB_SPEC_ID
#define XB_SPEC_ID 0
#endif
#if XB_SPEC && XB_SPEC_ID > 0
#define XB_LAUNCH_NAME XB_CAT(XB_CAT(XB_CAT(launch_b4_group_a, XB_ARITH), x), XB_SPEC_ID)
#elif XB_SPEC
#define XB_LAUNCH_NAME XB_CAT(XB_CAT(launch_b4_group_a, XB_ARITH), x)
#elif XB_LOCKSTEP
#define XB_LAUNCH_NAME XB_CAT(XB_CAT(launch_b4_group_a, XB_ARITH), s)
#else
#define XB_LAUNCH_NAME XB_CAT(launch_b4_group_a, XB_ARITH)
#endif

int XB_LAUNCH_NAME(const GroupDev &g, const LoadArgs &a, int threads, int minblocks, const BinPack *packs, int npacks,
                   cudaStream_t stream) {
  if (g.n <= 0) return 0;
  const int launches = (packs && npacks > 0) ? npacks : 1;
  if (g.general) {
    launch_variant<true, 128, 2>(g, a, packs, npacks, stream);
    return launches;
  }
#define XB_CASE(T, B) if (threads == T && minblocks == B) { launch_variant<false, T, B>(g, a, packs, npacks, stream); return launches; }
  XB_B4_LAUNCH_SHAPES(XB_CASE)
#undef XB_CASE
  return -1;   // unsupported (threads, minblocks) pair
}

}  // namespace b4
}  // namespace xb
