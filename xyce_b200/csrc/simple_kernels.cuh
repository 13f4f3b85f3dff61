// xyce_b200 -- device groups of the small compact models: one thread per instance, flat SoA records,
// contributions written to the same assembly planes as BSIM4 (assembly.cuh).
#pragma once
#include <cuda_runtime.h>
#include "b4_kernels.cuh"   // LoadArgs

namespace xb {
namespace simple {

enum Type { kDiode = 1, kMos1 = 2, kBjt = 3, kRlc = 4, kMvs = 5 };
int bjt_excess_phase_field();      // index of excessPhaseFac in the BJT record
int lead_count(int type);           // branch-data entries per instance (0 = the type has no lead currents here)
// Models produced by the ADMS translator (xyce_b200/adms/translate.py -> gen_adms/registry.h at build time) take the type
// ids kAdmsGenBase + position in the registry; adms_gen_* describe them to the callers of xgpu_simple_group_add.
constexpr int kAdmsGenBase = 100;
int adms_gen_count();
const char *adms_gen_name(int idx);        // nullptr when idx is out of range
const char *adms_gen_fields(int idx);      // space-separated record fields: "M:x" model member, "I:x" instance member
int adms_gen_ext(int idx);                 // number of external nodes

struct GroupDev {
  int type, n;
  const double *rec;      // [nfields][n]
  const int *flags;       // [n]
  const int *lids;        // [nodes][n], -1 = ground
  const int *sto_lid0, *sta_lid0;
  int sto_stride, sta_stride;
  int *orig_flag;
  long long vec_base, mat_base;
  double *lead;           // BJT with lead currents requested: [8][n] block (F ib ie ic is, Q ib ie ic is), else null
};

// static description of a device type (host side)
struct TypeInfo { int nodes, slots, nfields, nstore, nstate; const int *slot_row, *slot_col; };
const TypeInfo *type_info(int type);

void launch_group(const GroupDev &g, const b4::LoadArgs &a, cudaStream_t s);
// copies / derives leadF, leadQ, junctionV of one group at its branch-data LIDs (assign): after the evaluation launch
void launch_lead(const GroupDev &g, const int *branch0, const double *planeF, const double *planeQ, const double *sol,
                 double *leadF, double *leadQ, double *junctionV, cudaStream_t s);
// adms_gen_kernels.cu (its own translation unit: fast arithmetic variant)
const TypeInfo *adms_gen_type_info(int type);
void launch_adms_gen_group(const GroupDev &g, const b4::LoadArgs &a, cudaStream_t s);

}  // namespace simple
}  // namespace xb
