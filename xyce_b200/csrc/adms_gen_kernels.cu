// xyce_b200 -- kernels of the models written by the ADMS translator (xyce_b200/adms/translate.py -> gen_adms/).
//
// Any model that admsXml emits with Xyce's `_nosac` templates has the same shape as the hand-restated MVS of
// simple_kernels.cu: flat per-instance record, node voltages through the gather map, static + dynamic contributions and
// their probe derivatives copied onto rows and Jacobian stamp slots; no limiting, no store / state.  One kernel template
// serves them all; the registry (gen_adms/registry.h, written at build time) instantiates it per model.
// Its own translation unit so that it can use the fast arithmetic variant (FastReal: shared-reciprocal division,
// constant-bank exp / log, inlined sqrt, FMA contraction; <= 2 ulp per operation, tests hold 1e-12 against the
// reference's generated classes) independently of the strict small-device kernels.
#include "xb_real.h"
#define XB_REAL xb::FastReal
#include "pdl.cuh"
#include "simple_kernels.cuh"
#include "xb_common.h"
#if defined(__has_include)
#if __has_include("gen_adms/registry.h")
#include "gen_adms/registry.h"
#define XB_HAVE_ADMS_GEN 1
#endif
#endif

namespace xb {
namespace simple {

namespace {

// field k of the record is loaded (coalesced, read-only path) where the analog block reads it: a 76-field record
// (EKV) does not sit in registers for the whole evaluation
struct LazyRec {
  struct Fields {
    const double *p; size_t n;
    __device__ __forceinline__ real operator[](int k) const { return real(__ldg(p + (size_t)k * n)); }
  } f;
};

template <class T>
__global__ void __launch_bounds__(128) adms_gen_kernel(GroupDev g, b4::LoadArgs a) {
  xb::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.n) return;
  const int n = g.n;
  const LazyRec R{{g.rec + i, (size_t)n}};
  real V[T::kNodes];
#pragma unroll
  for (int t = 0; t < T::kNodes; ++t) {
    const int lid = __ldg(g.lids + (size_t)t * n + i);
    V[t] = real(lid >= 0 ? __ldg(a.sol + lid) : 0.0);
  }
  typename T::Out o;
  T::eval(a.S, R, V, o);
  g.orig_flag[i] = 1;
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

#ifdef XB_HAVE_ADMS_GEN
#define XB_GEN_INFO(idx_, nm_) {adms::gen_##nm_::Traits::kNodes, adms::gen_##nm_::Traits::kSlots, adms::gen_##nm_::Traits::kNumFields, 0, 0, \
                                adms::gen_##nm_::Traits::slot_row(), adms::gen_##nm_::Traits::slot_col()},
const TypeInfo kGenInfo[XB_ADMS_GEN_COUNT] = {XB_ADMS_GEN_LIST(XB_GEN_INFO)};
#undef XB_GEN_INFO
#endif

}  // namespace

const TypeInfo *adms_gen_type_info(int type) {
#ifdef XB_HAVE_ADMS_GEN
  if (type >= kAdmsGenBase && type < kAdmsGenBase + XB_ADMS_GEN_COUNT) return &kGenInfo[type - kAdmsGenBase];
#endif
  (void)type;
  return nullptr;
}

int adms_gen_count() {
#ifdef XB_HAVE_ADMS_GEN
  return XB_ADMS_GEN_COUNT;
#else
  return 0;
#endif
}
const char *adms_gen_name(int idx) {
#ifdef XB_HAVE_ADMS_GEN
#define XB_GEN_NAME(i, nm_) if (idx == i) return adms::gen_##nm_::Traits::name();
  XB_ADMS_GEN_LIST(XB_GEN_NAME)
#undef XB_GEN_NAME
#endif
  (void)idx;
  return nullptr;
}
const char *adms_gen_fields(int idx) {
#ifdef XB_HAVE_ADMS_GEN
#define XB_GEN_FIELDS(i, nm_) if (idx == i) return adms::gen_##nm_::Traits::fields();
  XB_ADMS_GEN_LIST(XB_GEN_FIELDS)
#undef XB_GEN_FIELDS
#endif
  (void)idx;
  return nullptr;
}
int adms_gen_ext(int idx) {
#ifdef XB_HAVE_ADMS_GEN
#define XB_GEN_EXT(i, nm_) if (idx == i) return adms::gen_##nm_::Traits::kExt;
  XB_ADMS_GEN_LIST(XB_GEN_EXT)
#undef XB_GEN_EXT
#endif
  (void)idx;
  return -1;
}

void launch_adms_gen_group(const GroupDev &g, const b4::LoadArgs &a, cudaStream_t s) {
  if (g.n <= 0) return;
  const int blocks = (g.n + 127) / 128;
#ifdef XB_HAVE_ADMS_GEN
#define XB_GEN_LAUNCH(i, nm_) if (g.type == kAdmsGenBase + i) xb::launch_pdl(adms_gen_kernel<adms::gen_##nm_::Traits>, dim3(blocks), dim3(128), 0, s, g, a);
  XB_ADMS_GEN_LIST(XB_GEN_LAUNCH)
#undef XB_GEN_LAUNCH
#endif
  (void)blocks; (void)a; (void)s;
}

}  // namespace simple
}  // namespace xb
