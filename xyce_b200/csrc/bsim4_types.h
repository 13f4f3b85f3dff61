// xyce_b200 -- BSIM4 record types, general-stamp node and Jacobian-slot ids.
#pragma once
#include "xb_common.h"
#include "bsim4_fields.def"

namespace xb {
namespace b4 {

#define XB_DECL_D(n) double n;
#define XB_DECL_I(n) int n;

struct B4Model { XB_B4_MODEL_D(XB_DECL_D) XB_B4_MODEL_I(XB_DECL_I) };
struct B4Size  { XB_B4_SIZE_D(XB_DECL_D) };
struct B4Inst  { XB_B4_INST_D(XB_DECL_D) XB_B4_INST_I(XB_DECL_I) };
// Every intermediate that outlives one evaluation stage.  Names follow the
// reference's Instance members so host tests can compare them one by one.
#define XB_DECL_R(n) real n;
struct B4Mid   { XB_B4_MID_D(XB_DECL_R) XB_B4_MID_EXTRA_D(XB_DECL_R) XB_B4_MID_I(XB_DECL_I) };

#define XB_COUNT(n) +1
constexpr int kNumModelD = 0 XB_B4_MODEL_D(XB_COUNT);
constexpr int kNumModelI = 0 XB_B4_MODEL_I(XB_COUNT);
constexpr int kNumSizeD  = 0 XB_B4_SIZE_D(XB_COUNT);
constexpr int kNumInstD  = 0 XB_B4_INST_D(XB_COUNT);
constexpr int kNumInstI  = 0 XB_B4_INST_I(XB_COUNT);
#undef XB_COUNT

// General-stamp node order: the reference's 11-row jacStamp
// (N_DEV_MOSFET_B4.C:5829-5907) plus the NQS charge node.
enum B4Node {
  kD = 0, kGE = 1, kS = 2, kB = 3, kDP = 4, kSP = 5, kGP = 6, kGM = 7, kBP = 8, kSB = 9, kDB = 10,
  kQ = 11, kNumNodes = 12, kNumRows = 11
};

// Jacobian slots of the general stamp, row-major in jacStamp order (62 entries).
// Name = s<ROW><col>, matching the reference's offset names (Dd, Ddp, ... DBdb).
#define XB_B4_SLOTS(X) \
  X(Dd, kD, kD) X(Ddp, kD, kDP) X(Dsp, kD, kSP) X(Dgp, kD, kGP) X(Dbp, kD, kBP) \
  X(GEge, kGE, kGE) X(GEdp, kGE, kDP) X(GEsp, kGE, kSP) X(GEgp, kGE, kGP) X(GEgm, kGE, kGM) X(GEbp, kGE, kBP) \
  X(Ss, kS, kS) X(Sdp, kS, kDP) X(Ssp, kS, kSP) X(Sgp, kS, kGP) X(Sbp, kS, kBP) \
  X(Bb, kB, kB) X(Bbp, kB, kBP) X(Bsb, kB, kSB) X(Bdb, kB, kDB) \
  X(DPd, kDP, kD) X(DPdp, kDP, kDP) X(DPsp, kDP, kSP) X(DPgp, kDP, kGP) X(DPgm, kDP, kGM) X(DPbp, kDP, kBP) X(DPdb, kDP, kDB) \
  X(SPs, kSP, kS) X(SPdp, kSP, kDP) X(SPsp, kSP, kSP) X(SPgp, kSP, kGP) X(SPgm, kSP, kGM) X(SPbp, kSP, kBP) X(SPsb, kSP, kSB) \
  X(GPge, kGP, kGE) X(GPdp, kGP, kDP) X(GPsp, kGP, kSP) X(GPgp, kGP, kGP) X(GPgm, kGP, kGM) X(GPbp, kGP, kBP) \
  X(GMge, kGM, kGE) X(GMdp, kGM, kDP) X(GMsp, kGM, kSP) X(GMgp, kGM, kGP) X(GMgm, kGM, kGM) X(GMbp, kGM, kBP) \
  X(BPb, kBP, kB) X(BPdp, kBP, kDP) X(BPsp, kBP, kSP) X(BPgp, kBP, kGP) X(BPgm, kBP, kGM) X(BPbp, kBP, kBP) X(BPsb, kBP, kSB) X(BPdb, kBP, kDB) \
  X(SBb, kSB, kB) X(SBsp, kSB, kSP) X(SBbp, kSB, kBP) X(SBsb, kSB, kSB) \
  X(DBb, kDB, kB) X(DBdp, kDB, kDP) X(DBbp, kDB, kBP) X(DBdb, kDB, kDB)

enum B4Slot {
#define XB_SLOT_ENUM(name, r, c) s##name,
  XB_B4_SLOTS(XB_SLOT_ENUM)
#undef XB_SLOT_ENUM
  kNumSlots
};
static_assert(kNumSlots == 62, "general BSIM4 stamp has 62 entries");

#define XB_SLOT_ROW2(name, r, c) r,
#define XB_SLOT_COL2(name, r, c) c,
#define XB_SLOT_ROW(name, r, c) r,
#define XB_SLOT_COL(name, r, c) c,
// row / column node of each slot (host + device constexpr tables)
constexpr int kSlotRow[kNumSlots] = { XB_B4_SLOTS(XB_SLOT_ROW) };
constexpr int kSlotCol[kNumSlots] = { XB_B4_SLOTS(XB_SLOT_COL) };
#undef XB_SLOT_ROW
#undef XB_SLOT_COL

// Default topology (rgateMod = rbodyMod = 0, no S/D resistor nodes, no NQS):
// general node -> one of the 4 external terminals {D,G,S,B} = {0,1,2,3}.
constexpr int kDefaultCollapse[kNumRows] = {0, 1, 2, 3, 0, 2, 1, 1, 3, 3, 3};

// constexpr accessors usable in device code (namespace-scope constexpr arrays are host-only)
XB_HD constexpr int slot_row(int s) { constexpr int t[kNumSlots] = { XB_B4_SLOTS(XB_SLOT_ROW2) }; return t[s]; }
XB_HD constexpr int slot_col(int s) { constexpr int t[kNumSlots] = { XB_B4_SLOTS(XB_SLOT_COL2) }; return t[s]; }
XB_HD constexpr int default_collapse(int r) { constexpr int t[kNumRows] = {0, 1, 2, 3, 0, 2, 1, 1, 3, 3, 3}; return t[r]; }

// Store-vector slot order (Instance::registerStoreLIDs, N_DEV_MOSFET_B4.C:6445-6487).
enum B4Store {
  st_vbd, st_vbs, st_vgs, st_vds, st_vges, st_vgms, st_vdes, st_vses, st_vdbs, st_vsbs, st_vdbd,
  st_vged, st_vgmd, st_gm, st_Vds, st_Vgs, st_Vbs, st_Vdsat, st_Vth, st_Gds, st_Cgs, st_Cgd,
  kNumStore
};
static_assert(kNumStore == 22, "BSIM4 has 22 store variables");
// State-vector slot order (registerStateLIDs, :6380-6435) for the default topology.
enum B4State { sa_qb, sa_qg, sa_qd, kNumStateDefault };

}  // namespace b4
}  // namespace xb
