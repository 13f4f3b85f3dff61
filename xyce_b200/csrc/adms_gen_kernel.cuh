// xyce_b200 -- the generic kernel of the models written by the ADMS translator (xyce_b200/adms/translate.py).
//
// Any model that admsXml emits with Xyce's `_nosac` templates has the same shape as the hand-restated MVS of
// simple_kernels.cu: flat per-instance record, node voltages through the gather map, static + dynamic contributions and
// their probe derivatives copied onto rows and Jacobian stamp slots; no limiting, no store / state.  One kernel template
// serves them all; the translator writes one small translation unit per model (gen_adms/kernel_<model>.cu) that
// instantiates it, so the library build compiles the models in parallel.
// Fast arithmetic variant (FastReal: shared-reciprocal division, lean exp / log, inlined sqrt, FMA contraction; <= 2 ulp
// per operation, tests hold 1e-12 against the reference's generated classes), independent of the strict small-device
// kernels of simple_kernels.cu.
#pragma once
#if !defined(XB_ADMS_STRICT)      // -DXB_ADMS_STRICT: plain double, IEEE division, libdevice math (diagnostics)
#include "xb_real.h"
#define XB_REAL xb::FastReal
#endif
#include "pdl.cuh"
#include "simple_kernels.cuh"
#include "xb_common.h"

namespace xb {
namespace simple {

// field k of the record is loaded (coalesced, read-only path) where the analog block reads it: a 600-field record
// (PSP 103) does not sit in registers for the whole evaluation
struct LazyRec {
  struct Fields {
    const double *p; size_t n;
    __device__ __forceinline__ real operator[](int k) const { return real(__ldg(p + (size_t)k * n)); }
  } f;
};

constexpr int kAdmsPrefetchMaxFields = 232;

template <class T>
__global__ void __launch_bounds__(128) adms_gen_kernel(GroupDev g, b4::LoadArgs a) {
  xb::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = g.n;
  if (T::kNumFields <= kAdmsPrefetchMaxFields) {
    // the warp asks L2 for its 256 B of every record field up front: the lazy loads below then wait for an L2 hit, not a
    // DRAM miss.  Measured (200 000 instances, profiles/r02_simple_kernels_timing_v3.json): JUNCAP200 -17 %, EKV -10 %,
    // HICUM/L0 -7 %, VBIC -3 %; records of which one evaluation reads a small part (PSP 103, BSIM-CMG: +6..11 %) are
    // left alone.
    const int lane = threadIdx.x & 31, i0 = i - lane;
    for (int k = lane; k < 2 * T::kNumFields; k += 32) {
      const int ii = i0 + (k & 1) * 16;
      if (ii < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(g.rec + (size_t)(k >> 1) * n + ii));
    }
  }
  if (i >= n) return;
  const LazyRec R{{g.rec + i, (size_t)n}};
  real V[T::kNodes];
#pragma unroll
  for (int t = 0; t < T::kNodes; ++t) {
    const int lid = __ldg(g.lids + (size_t)t * n + i);
    V[t] = real(lid >= 0 ? __ldg(a.sol + lid) : 0.0);
  }
  typename T::Out o;
  const int sto0 = (T::kNumStore > 0) ? __ldg(g.sto_lid0 + i) : 0, ss = g.sto_stride;
  if (T::kHasLimit) {          // $limit: previous-iterate values of the limited probes come from the store vectors
    real cs[T::kNumStore > 0 ? T::kNumStore : 1], ns[T::kNumStore > 0 ? T::kNumStore : 1];
#pragma unroll
    for (int t = 0; t < T::kNumStore; ++t) { cs[t] = a.curr_sto[sto0 + (size_t)t * ss]; ns[t] = a.next_sto[sto0 + (size_t)t * ss]; }
    T::eval(a.S, R, V, o, cs, ns);
  } else {
    T::eval(a.S, R, V, o);
  }
  g.orig_flag[i] = o.origFlag;
  if (T::kNumStore > 0) {      // limited probes and output variables -> the store vector, like updatePrimaryState
#pragma unroll
    for (int t = 0; t < T::kNumStore; ++t) a.next_sto[sto0 + (size_t)t * ss] = to_double(o.store[t]);
  }
#pragma unroll
  for (int r = 0; r < T::kNodes; ++r) {
    a.vec_planes[0][g.vec_base + (size_t)r * n + i] = to_double(o.F[r]);
    a.vec_planes[1][g.vec_base + (size_t)r * n + i] = to_double(o.Q[r]);
    a.vec_planes[2][g.vec_base + (size_t)r * n + i] = to_double(o.FL[r]);
    a.vec_planes[3][g.vec_base + (size_t)r * n + i] = to_double(o.QL[r]);
  }
#pragma unroll
  for (int s = 0; s < T::kSlots; ++s) {
    a.mat_planes[0][g.mat_base + (size_t)s * n + i] = to_double(o.JF[s]);
    a.mat_planes[1][g.mat_base + (size_t)s * n + i] = to_double(o.JQ[s]);
  }
}

template <class T>
inline void launch_adms_gen(const GroupDev &g, const b4::LoadArgs &a, cudaStream_t s) {
  if (g.n <= 0) return;
  xb::launch_pdl(adms_gen_kernel<T>, dim3((g.n + 127) / 128), dim3(128), 0, s, g, a);
}

}  // namespace simple
}  // namespace xb
