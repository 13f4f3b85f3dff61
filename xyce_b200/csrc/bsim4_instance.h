// xyce_b200 -- BSIM4: one complete instance evaluation = the work the reference does in
// Master::updateState + loadDAEVectors + loadDAEMatrices for one instance
// (N_DEV_MOSFET_B4.C:10540-11686), expressed as one inlinable call.
#pragma once
#include "bsim4_load.h"

namespace xb {
namespace b4 {

// Which store vector holds the "old" limiting voltages (B4p82.C:3096-3148):
//   newtonIter == 0 and (!dcop || (loca && dcop))  -> currSto
//   newtonIter == 0 otherwise                      -> no history (old = present)
//   newtonIter != 0                                -> nextSto
enum OldSource { kOldCurr = 0, kOldNone = 1, kOldNext = 2 };
XB_HD int old_source(const SolverFlags &S) {
  if (S.newtonIter == 0) return (!S.dcopFlag || (S.locaEnabledFlag && S.dcopFlag)) ? kOldCurr : kOldNone;
  return kOldNext;
}

template <class E>
XB_HD void evaluate(const SolverFlags &S, const B4Model &M, const B4Size &P, const B4Inst &I,
                    const real *V, const real *sto_old, bool have_old, real von_prev,
                    B4Mid &W, E &e) {
  DcCarry C;
  stage_voltages(S, M, I, V, sto_old, have_old, von_prev, W);
  XB_SYNC_POINT(1);
  stage_dc(S, M, P, I, W, C);
  XB_SYNC_POINT(1);
  stage_cv(S, M, P, I, W, C);
  XB_SYNC_POINT(1);
  stage_caps(M, P, I, W);
  XB_SYNC_POINT(1);
  stage_fvars(M, I, W);
  XB_SYNC_POINT(1);
  emit_vectors(S, M, I, W, e);
  XB_SYNC_POINT(1);
  emit_matrices(M, I, W, e);
  XB_SYNC_POINT(2);
}

// The 22 store-vector values published by Master::updateState (N_DEV_MOSFET_B4.C:10552-10580).
// vged / vgmd slots are never written by the reference; callers must leave them untouched.
template <class F>
XB_HD void for_each_store(const B4Mid &W, F put) {
  put(st_vbd, W.vbd); put(st_vbs, W.vbs); put(st_vgs, W.vgs); put(st_vds, W.vds);
  put(st_vges, W.vges); put(st_vgms, W.vgms); put(st_vdes, W.vdes); put(st_vses, W.vses);
  put(st_vdbs, W.vdbs); put(st_vsbs, W.vsbs); put(st_vdbd, W.vdbd);
  put(st_gm, (W.mode >= 0) ? W.gm : -W.gm);
  put(st_Vds, W.Vds_s); put(st_Vgs, W.Vgs_s); put(st_Vbs, W.Vbs_s);
  put(st_Vdsat, W.Vdsat); put(st_Vth, W.Vth);
  put(st_Gds, W.gds); put(st_Cgs, W.cgsb); put(st_Cgd, W.cgdb);
}

}  // namespace b4
}  // namespace xb
