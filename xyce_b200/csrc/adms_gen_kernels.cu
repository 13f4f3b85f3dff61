// xyce_b200 -- registry and dispatch of the models written by the ADMS translator (xyce_b200/adms/translate.py).
// The kernels themselves are one translation unit per model (gen_adms/kernel_<model>.cu, instantiating
// adms_gen_kernel.cuh); this file only holds the tables (gen_adms/registry_info.h, written at build time) behind
// xgpu_adms_gen_count / xgpu_adms_gen_info and routes a group's launch to its model.
#include "simple_kernels.cuh"
#if defined(__has_include)
#if __has_include("gen_adms/registry_info.h")
#include "gen_adms/registry_info.h"
#define XB_HAVE_ADMS_GEN 1
#endif
#endif

namespace xb {
namespace simple {

#ifdef XB_HAVE_ADMS_GEN
#define XB_GEN_DECL(i_, nm_, nodes_, ext_, slots_, nf_, nsto_) void launch_adms_gen_##nm_(const GroupDev &g, const b4::LoadArgs &a, cudaStream_t s);
XB_ADMS_GEN_LIST(XB_GEN_DECL)
#undef XB_GEN_DECL
namespace {
#define XB_GEN_INFO(i_, nm_, nodes_, ext_, slots_, nf_, nsto_) {nodes_, slots_, nf_, nsto_, 0, kAdmsRow_##nm_, kAdmsCol_##nm_},
const TypeInfo kGenInfo[XB_ADMS_GEN_COUNT] = {XB_ADMS_GEN_LIST(XB_GEN_INFO)};
#undef XB_GEN_INFO
}  // namespace
#endif

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
#define XB_GEN_NAME(i_, nm_, nodes_, ext_, slots_, nf_, nsto_) if (idx == i_) return #nm_;
  XB_ADMS_GEN_LIST(XB_GEN_NAME)
#undef XB_GEN_NAME
#endif
  (void)idx;
  return nullptr;
}
const char *adms_gen_fields(int idx) {
#ifdef XB_HAVE_ADMS_GEN
#define XB_GEN_FIELDS(i_, nm_, nodes_, ext_, slots_, nf_, nsto_) if (idx == i_) return kAdmsFields_##nm_;
  XB_ADMS_GEN_LIST(XB_GEN_FIELDS)
#undef XB_GEN_FIELDS
#endif
  (void)idx;
  return nullptr;
}
int adms_gen_ext(int idx) {
#ifdef XB_HAVE_ADMS_GEN
#define XB_GEN_EXT(i_, nm_, nodes_, ext_, slots_, nf_, nsto_) if (idx == i_) return ext_;
  XB_ADMS_GEN_LIST(XB_GEN_EXT)
#undef XB_GEN_EXT
#endif
  (void)idx;
  return -1;
}

void launch_adms_gen_group(const GroupDev &g, const b4::LoadArgs &a, cudaStream_t s) {
#ifdef XB_HAVE_ADMS_GEN
#define XB_GEN_LAUNCH(i_, nm_, nodes_, ext_, slots_, nf_, nsto_) if (g.type == kAdmsGenBase + i_) { launch_adms_gen_##nm_(g, a, s); return; }
  XB_ADMS_GEN_LIST(XB_GEN_LAUNCH)
#undef XB_GEN_LAUNCH
#endif
  (void)g; (void)a; (void)s;
}

}  // namespace simple
}  // namespace xb
