// xyce_b200 -- MOSFET level 1 (Shichman-Hodges + Meyer capacitances): one instance evaluation =
//   Instance::updateIntermediateVars                     (src/DeviceModelPKG/OpenModels/N_DEV_MOSFET1.C:2689-3485)
//   Master::updateState / loadDAEVectors / loadDAEMatrices  (N_DEV_MOSFET1.C:4241-4331, :4343-4585, :4587-4876)
//   DeviceSupport::qmeyer                                (src/DeviceModelPKG/Core/N_DEV_DeviceSupport.C:507-579)
// restated for a one-thread-per-instance SoA kernel.  The default (back-averaging) Meyer formulation only:
// DeviceOptions::newMeyerFlag (extra "dot" unknowns) is rejected by the C ABI.
// Nodes: 0 D, 1 G, 2 S, 3 B, 4 D' (aliases D when RD = 0), 5 S' (aliases S when RS = 0).
// Store: vbd vbs vgs vds von gm.  State: qgs qgd qgb capgs capgd capgb qbd qbs.
#pragma once
#include "xb_common.h"
#include "simple_fields.def"

namespace xb {
namespace mos1 {

constexpr double kMaxExpArg = 100.0;    // CONSTMAX_EXP_ARG
enum { kD = 0, kG, kS, kB, kDP, kSP, kNodes };
// jacStamp_DC_SC without the "dot" unknowns (N_DEV_MOSFET1.C:1294-1325), row-major
enum { sDd = 0, sDdp, sGg, sGb, sGdp, sGsp, sSs, sSsp, sBg, sBb, sBdp, sBsp, sDPd, sDPg, sDPb, sDPdp, sDPsp,
       sSPg, sSPs, sSPb, sSPdp, sSPsp, kSlots };
XB_HD constexpr int slot_row(int s) {
  constexpr int t[kSlots] = {0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5};
  return t[s];
}
XB_HD constexpr int slot_col(int s) {
  constexpr int t[kSlots] = {0, 4, 1, 3, 4, 5, 2, 5, 1, 3, 4, 5, 0, 1, 3, 4, 5, 1, 2, 3, 4, 5};
  return t[s];
}
enum { fIcGiven = 1, fOff = 2 };
enum { st_vbd = 0, st_vbs, st_vgs, st_vds, st_von, st_gm, kNumStore };
enum { sa_qgs = 0, sa_qgd, sa_qgb, sa_capgs, sa_capgd, sa_capgb, sa_qbd, sa_qbs, kNumState };

#define XB_M_DECL(n) double n;
struct Rec { XB_MOS1_FIELDS(XB_M_DECL, XB_M_DECL) };
#undef XB_M_DECL
#define XB_CNT(n) +1
constexpr int kNumFields = 0 XB_MOS1_FIELDS(XB_CNT, XB_CNT);
#undef XB_CNT

struct Out {
  real F[kNodes], Q[kNodes], FL[kNodes], QL[kNodes], JF[kSlots], JQ[kSlots];
  real store[kNumStore], state[kNumState];
  int origFlag;       // 0 when any limited voltage differs from the solution (switches the Jdxp terms on)
  int converged;      // Instance::isConverged() = !limitedFlag: only pnjlim invalidates convergence (N_DEV_MOSFET1.h:771-774)
};

// Meyer gate capacitances (half of the non-constant part; N_DEV_DeviceSupport.C:507-579)
XB_HD void qmeyer(real vgs, real vgd, real von, real vdsat, real &capgs, real &capgd, real &capgb, real phi, real cox) {
  const real vgst = vgs - von;
  if (vgst <= -phi) {
    capgb = cox / 2; capgs = 0; capgd = 0;
  } else if (vgst <= -phi / 2) {
    capgb = -vgst * cox / (2 * phi); capgs = 0; capgd = 0;
  } else if (vgst <= 0) {
    capgb = -vgst * cox / (2 * phi);
    capgs = vgst * cox / (1.5 * phi) + cox / 3;
    capgd = 0;
  } else {
    const real vds = vgs - vgd;
    if (vdsat <= vds) {
      capgs = cox / 3; capgd = 0; capgb = 0;
    } else {
      const real vddif = 2.0 * vdsat - vds;
      const real vddif1 = vdsat - vds;
      const real vddif2 = vddif * vddif;
      capgd = cox * (1.0 - vdsat * vdsat / vddif2) / 3;
      capgs = cox * (1.0 - vddif1 * vddif1 / vddif2) / 3;
      capgb = 0;
    }
  }
}

XB_HD real grading(real arg, real mj) {      // arg^-mj with the sqrt special case of the reference
  if (mj == .5) return 1 / sqrt(arg);
  return exp(-mj * log(arg));
}

// curr_sto / next_sto: the 6 store entries of the current / next store vector;
// curr_sta: the 8 state entries of the current state vector (Meyer back-averaging history)
XB_HD void evaluate(const SolverFlags &S, const Rec &M, int flags, const real *V, const real *curr_sto,
                    const real *next_sto, const real *curr_sta, Out &o) {
  const real ty = M.dtype;
  real DrainSatCur, SourceSatCur;
  if ((M.tSatCurDens == 0) || (M.drainArea == 0) || (M.sourceArea == 0)) {
    DrainSatCur = M.tSatCur; SourceSatCur = M.tSatCur;
  } else {
    DrainSatCur = M.tSatCurDens * M.drainArea; SourceSatCur = M.tSatCurDens * M.sourceArea;
  }
  const real Beta = M.tTransconductance * M.w / M.EffectiveLength;
  const real Vd = V[kD], Vg = V[kG], Vs = V[kS], Vb = V[kB], Vsp = V[kSP], Vdp = V[kDP];
  const real Vddp = Vd - Vdp, Vssp = Vs - Vsp;
  real vbs = ty * (Vb - Vsp), vgs = ty * (Vg - Vsp), vds = ty * (Vdp - Vsp);
  real vbd = vbs - vds, vgd = vgs - vds;
  int origFlag = 1;
  const real vgs_orig = vgs, vds_orig = vds, vbs_orig = vbs, vbd_orig = vbd, vgd_orig = vgd;
  const bool OFF = (flags & fOff) != 0;

  if (S.initJctFlag && !OFF && S.voltageLimiterFlag) {
    if (flags & fIcGiven) {
      vds = ty * M.icVDS; vgs = ty * M.icVGS; vbs = ty * M.icVBS;
      vbd = vbs - vds; vgd = vgs - vds;
      origFlag = 0;
    } else {     // (inputOPFlag is rejected by the C ABI)
      vbs = -1; vgs = ty * M.tVto; vds = 0;
      vbd = vbs - vds; vgd = vgs - vds;
    }
  } else if ((S.initFixFlag || S.initJctFlag) && OFF) {
    vbs = vgs = vds = 0;
    vbd = vgd = 0;
  }

  real vbs_old, vbd_old, vgs_old, vds_old, vgd_old, Von;
  if (S.newtonIter == 0) {
    if (!S.dcopFlag || (S.locaEnabledFlag && S.dcopFlag)) {
      vbs_old = curr_sto[st_vbs]; vbd_old = curr_sto[st_vbd]; vgs_old = curr_sto[st_vgs]; vds_old = curr_sto[st_vds];
      Von = ty * curr_sto[st_von];
    } else {
      vbs_old = vbs; vbd_old = vbd; vgs_old = vgs; vds_old = vds;
      Von = 0.0;
    }
    vgd_old = vgs_old - vds_old;
  } else {
    vbs_old = next_sto[st_vbs]; vbd_old = next_sto[st_vbd]; vgs_old = next_sto[st_vgs]; vds_old = next_sto[st_vds];
    Von = ty * next_sto[st_von];
    vgd_old = vgs_old - vds_old;
  }

  int limited = 0;
  if (S.voltageLimiterFlag) {
    if (!(S.initFixFlag && OFF)) {
      int Check = 1;
      if (vds_old >= 0) {
        vgs = fetlim(vgs, vgs_old, Von);
        vds = vgs - vgd;
        vds = limvds(vds, vds_old);
        vgd = vgs - vds;
      } else {
        vgd = fetlim(vgd, vgd_old, Von);
        vds = vgs - vgd;
        vds = -limvds(-vds, -vds_old);
        vgs = vgd + vds;
      }
      if (vds >= 0.0) {
        vbs = pnjlim(vbs, vbs_old, M.vt, M.sourceVcrit, Check);
        vbd = vbs - vds;
      } else {
        vbd = pnjlim(vbd, vbd_old, M.vt, M.drainVcrit, Check);
        vbs = vbd + vds;
      }
      if (Check == 1) limited = 1;
    }
  }
  vbd = vbs - vds;
  vgd = vgs - vds;
  const real Vgb = vgs - vbs;
  if (vgs_orig != vgs || vds_orig != vds || vbs_orig != vbs || vbd_orig != vbd || vgd_orig != vgd) origFlag = 0;

  // bulk-source / bulk-drain diodes
  real gbs, cbs, gbd, cbd;
  if (vbs <= 0) {
    gbs = SourceSatCur / M.vt;
    gbs += S.gmin;
    cbs = gbs * vbs;
  } else {
    const real evbs = exp(dmin(kMaxExpArg, vbs / M.vt));
    gbs = (SourceSatCur * evbs / M.vt + S.gmin);
    cbs = (SourceSatCur * (evbs - 1) + S.gmin * vbs);
  }
  if (vbd <= 0) {
    gbd = DrainSatCur / M.vt;
    gbd += S.gmin;
    cbd = gbd * vbd;
  } else {
    const real evbd = exp(dmin(kMaxExpArg, vbd / M.vt));
    gbd = (DrainSatCur * evbd / M.vt + S.gmin);
    cbd = (DrainSatCur * (evbd - 1) + S.gmin * vbd);
  }
  const int mode = (vds >= 0) ? 1 : -1;

  // DCOP continuation ("MOSFET homotopy", N_DEV_MOSFET1.C:2978-3001): the drain-current and Meyer-capacitance code below
  // sees vds scaled by nltermScale and the controlling gate voltage blended towards vgstConst by gainScale
  // (DeviceSupport::contVds / contVgst, Core/N_DEV_DeviceSupport.C:779-835); the loads use the saved values (:3410-3412)
  const real vds_save = vds, vgs_save = vgs, vgd_save = vgd;
  if (S.artParameterFlag) {
    real mn = S.vdsScaleMin; if (mn <= 0.0) mn = 0.3;
    vds = vds * (S.nltermScale * (1.0 - mn) + mn);
    if (mode == 1) vgs = S.gainScale * vgs + (1.0 - S.gainScale) * S.vgstConst;
    else vgd = S.gainScale * vgd + (1.0 - S.gainScale) * S.vgstConst;
  }

  // Shichman-Hodges drain current
  real cdrain, gm, gds, gmbs, Vdsat;
  {
    const real vbx = (mode == 1 ? vbs : vbd);
    real sarg;
    if (vbx <= 0) {
      sarg = sqrt(M.tPhi - vbx);
    } else {
      sarg = sqrt(M.tPhi);
      sarg = sarg - vbx / (sarg + sarg);
      sarg = dmax(0.0, sarg);
    }
    Von = (M.tVbi * ty) + M.gamma * sarg;
    const real vgst = (mode == 1 ? vgs : vgd) - Von;
    Vdsat = dmax(vgst, 0.0);
    real arg;
    if (sarg <= 0) arg = 0;
    else arg = M.gamma / (sarg + sarg);
    if (vgst <= 0) {
      cdrain = 0; gm = 0; gds = 0; gmbs = 0;
    } else {
      const real betap = Beta * (1 + M.lambda * (vds * mode));
      if (vgst <= (vds * mode)) {
        cdrain = betap * vgst * vgst * .5;
        gm = betap * vgst;
        gds = M.lambda * Beta * vgst * vgst * .5;
        gmbs = gm * arg;
      } else {
        cdrain = betap * (vds * mode) * (vgst - .5 * (vds * mode));
        gm = betap * (vds * mode);
        gds = betap * (vgst - (vds * mode)) + M.lambda * Beta * (vds * mode) * (vgst - .5 * (vds * mode));
        gmbs = gm * arg;
      }
    }
  }
  const real von = ty * Von;

  // depletion charges of the bulk junctions
  real qbs, capbs, qbd, capbd;
  if (M.Cbs != 0 || M.Cbssw != 0) {
    if (vbs < M.tDepCap) {
      const real arg = 1 - vbs / M.tBulkPot;
      real sarg, sargsw;
      if (M.bulkJctBotGradingCoeff == M.bulkJctSideGradingCoeff) {
        sarg = sargsw = grading(arg, M.bulkJctBotGradingCoeff);
      } else {
        sarg = grading(arg, M.bulkJctBotGradingCoeff);
        sargsw = grading(arg, M.bulkJctSideGradingCoeff);
      }
      qbs = M.tBulkPot * (M.Cbs * (1 - arg * sarg) / (1 - M.bulkJctBotGradingCoeff) +
                          M.Cbssw * (1 - arg * sargsw) / (1 - M.bulkJctSideGradingCoeff));
      capbs = M.Cbs * sarg + M.Cbssw * sargsw;
    } else {
      qbs = M.f4s + vbs * (M.f2s + vbs * (M.f3s / 2));
      capbs = M.f2s + M.f3s * vbs;
    }
  } else {
    qbs = 0; capbs = 0;
  }
  if (M.Cbd != 0 || M.Cbdsw != 0) {
    if (vbd < M.tDepCap) {
      const real arg = 1 - vbd / M.tBulkPot;
      real sarg, sargsw;
      if (M.bulkJctBotGradingCoeff == .5 && M.bulkJctSideGradingCoeff == .5) {
        sarg = sargsw = 1 / sqrt(arg);
      } else {
        sarg = grading(arg, M.bulkJctBotGradingCoeff);
        sargsw = grading(arg, M.bulkJctSideGradingCoeff);
      }
      qbd = M.tBulkPot * (M.Cbd * (1 - arg * sarg) / (1 - M.bulkJctBotGradingCoeff) +
                          M.Cbdsw * (1 - arg * sargsw) / (1 - M.bulkJctSideGradingCoeff));
      capbd = M.Cbd * sarg + M.Cbdsw * sargsw;
    } else {
      qbd = M.f4d + vbd * (M.f2d + vbd * M.f3d / 2);
      capbd = M.f2d + vbd * M.f3d;
    }
  } else {
    qbd = 0; capbd = 0;
  }

  // Meyer capacitances, averaged with the previous time point during transient
  real capgs, capgd, capgb;
  if (mode > 0) qmeyer(vgs, vgd, Von, Vdsat, capgs, capgd, capgb, M.tPhi, M.OxideCap);
  else qmeyer(vgd, vgs, Von, Vdsat, capgd, capgs, capgb, M.tPhi, M.OxideCap);
  real Capgs, Capgd, Capgb;
  if (S.dcopFlag) {
    Capgs = 2.0 * capgs + M.GateSourceOverlapCap;
    Capgd = 2.0 * capgd + M.GateDrainOverlapCap;
    Capgb = 2.0 * capgb + M.GateBulkOverlapCap;
  } else {
    Capgs = (capgs + curr_sta[sa_capgs] + M.GateSourceOverlapCap);
    Capgd = (capgd + curr_sta[sa_capgd] + M.GateDrainOverlapCap);
    Capgb = (capgb + curr_sta[sa_capgb] + M.GateBulkOverlapCap);
  }
  Capgs *= (Capgs < 0.0) ? -1.0 : 1.0;
  Capgd *= (Capgd < 0.0) ? -1.0 : 1.0;
  Capgb *= (Capgb < 0.0) ? -1.0 : 1.0;

  const real Idrain = M.drainConductance * Vddp;
  const real Isource = M.sourceConductance * Vssp;
  real Gm, Gmbs, nrmsum, revsum, cdreq;
  if (mode >= 0) {
    Gm = gm; Gmbs = gmbs; nrmsum = Gm + Gmbs; revsum = 0; cdreq = ty * cdrain;
  } else {
    Gm = -gm; Gmbs = -gmbs; nrmsum = 0; revsum = -(Gm + Gmbs); cdreq = -(ty)*cdrain;
  }
  vds = vds_save; vgs = vgs_save; vgd = vgd_save;

  // ---- Master::updateState: store, Meyer charges ----
  o.store[st_vbd] = vbd; o.store[st_vbs] = vbs; o.store[st_vgs] = vgs; o.store[st_vds] = vds;
  o.store[st_von] = von; o.store[st_gm] = Gm;
  real qgs, qgd, qgb;
  if (S.dcopFlag) {
    qgs = Capgs * vgs; qgd = Capgd * vgd; qgb = Capgb * Vgb;
  } else {
    const real vgs1 = curr_sto[st_vgs], vbs1 = curr_sto[st_vbs], vds1 = curr_sto[st_vds];
    const real vgb1 = vgs1 - vbs1, vgd1 = vgs1 - vds1;
    qgs = curr_sta[sa_qgs]; qgd = curr_sta[sa_qgd]; qgb = curr_sta[sa_qgb];
    qgs += Capgs * (vgs - vgs1);
    qgd += Capgd * (vgd - vgd1);
    qgb += Capgb * ((vgs - vbs) - vgb1);
  }
  o.state[sa_qgs] = qgs; o.state[sa_qgd] = qgd; o.state[sa_qgb] = qgb;
  o.state[sa_capgs] = capgs; o.state[sa_capgd] = capgd; o.state[sa_capgb] = capgb;
  o.state[sa_qbd] = qbd; o.state[sa_qbs] = qbs;
  o.origFlag = origFlag;
  o.converged = !limited;

  // ---- Master::loadDAEVectors ----
  const real np = M.numberParallel;
  for (int i = 0; i < kNodes; ++i) o.F[i] = o.Q[i] = o.FL[i] = o.QL[i] = 0.0;
  const real ceqbs = ty * (cbs), ceqbd = ty * (cbd);
  const real ceqgb = 0.0, ceqgs = 0.0, ceqgd = 0.0;
  if (M.drainConductance != 0.0) o.F[kD] += Idrain * np;
  o.F[kG] += (ceqgs + ceqgd + ceqgb) * np;
  if (M.sourceConductance != 0.0) o.F[kS] += Isource * np;
  o.F[kB] += (ceqbs + ceqbd - ceqgb) * np;
  o.F[kDP] += (-Idrain - (ceqbd - cdreq + ceqgd)) * np;
  o.F[kSP] += (-Isource - (ceqbs + cdreq + ceqgs)) * np;
  {
    const real Qeqbs = ty * (qbs), Qeqbd = ty * (qbd), Qeqgb = ty * (qgb), Qeqgs = ty * (qgs), Qeqgd = ty * (qgd);
    o.Q[kG] += (Qeqgs + Qeqgd + Qeqgb) * np;
    o.Q[kB] += (Qeqbs + Qeqbd - Qeqgb) * np;
    o.Q[kDP] += (-(Qeqbd + Qeqgd)) * np;
    o.Q[kSP] += (-(Qeqbs + Qeqgs)) * np;
  }
  const bool caps_active = S.tranopFlag || S.acopFlag || S.transientFlag;
  real gcgd = 0.0, gcgs = 0.0, gcgb = 0.0, gcbs = 0.0, gcbd = 0.0;
  if (caps_active) { gcgd = Capgd; gcgs = Capgs; gcgb = Capgb; gcbs = capbs; gcbd = capbd; }
  if (!origFlag) {
    const real gmin1 = S.gmin;
    const real dvg = (mode > 0) ? (vgs - vgs_orig) : (vgd - vgd_orig);
    const real dvb = (mode > 0) ? (vbs - vbs_orig) : (vbd - vbd_orig);
    const real j4 = ty * (+((gbd - gmin1)) * (vbd - vbd_orig) + ((gbs - gmin1)) * (vbs - vbs_orig));
    const real j5 = ty * (-((gbd - gmin1)) * (vbd - vbd_orig) + gds * (vds - vds_orig) + Gm * dvg + Gmbs * dvb);
    const real j6 = ty * (-((gbs - gmin1)) * (vbs - vbs_orig) - gds * (vds - vds_orig) - Gm * dvg - Gmbs * dvb);
    o.FL[kB] += j4 * np;
    o.FL[kDP] += j5 * np;
    o.FL[kSP] += j6 * np;
    const real q2 = ty * (gcgd * (vgd - vgd_orig) + gcgs * (vgs - vgs_orig) + gcgb * (vgs - vgs_orig - vbs + vbs_orig));
    const real q4 = ty * (-(gcgb) * (vgs - vgs_orig - vbs + vbs_orig) + (gcgb) * (vbd - vbd_orig) + (gcbs) * (vbs - vbs_orig));
    const real q5 = ty * (-(gcgd) * (vgd - vgd_orig) - (gcbd) * (vbd - vbd_orig));
    const real q6 = ty * (-gcgs * (vgs - vgs_orig) - (gcbs) * (vbs - vbs_orig));
    o.QL[kG] += q2 * np;
    o.QL[kB] += q4 * np;
    o.QL[kDP] += q5 * np;
    o.QL[kSP] += q6 * np;
  }

  // ---- Master::loadDAEMatrices ----
  for (int s = 0; s < kSlots; ++s) o.JF[s] = o.JQ[s] = 0.0;
  const real gd = M.drainConductance, gs = M.sourceConductance;
  o.JF[sDd] += gd * np;
  o.JF[sDdp] -= gd * np;
  o.JF[sSs] += gs * np;
  o.JF[sSsp] -= gs * np;
  o.JF[sBb] += (gbs + gbd) * np;
  o.JF[sBdp] -= gbd * np;
  o.JF[sBsp] -= gbs * np;
  o.JF[sDPd] -= gd * np;
  o.JF[sDPg] += (Gm)*np;
  o.JF[sDPb] += (-gbd + Gmbs) * np;
  o.JF[sDPdp] += (gd + gds + gbd + revsum) * np;
  o.JF[sDPsp] += (-gds - nrmsum) * np;
  o.JF[sSPg] -= (Gm)*np;
  o.JF[sSPs] -= gs * np;
  o.JF[sSPb] -= (gbs + Gmbs) * np;
  o.JF[sSPdp] -= (gds + revsum) * np;
  o.JF[sSPsp] += (gs + gds + gbs + nrmsum) * np;
  o.JQ[sGg] += (gcgd + gcgs + gcgb) * np;
  o.JQ[sGb] -= gcgb * np;
  o.JQ[sGdp] -= gcgd * np;
  o.JQ[sGsp] -= gcgs * np;
  o.JQ[sBg] -= gcgb * np;
  o.JQ[sBb] += (+gcbs + gcbd + gcgb) * np;
  o.JQ[sBdp] -= +gcbd * np;
  o.JQ[sBsp] -= +gcbs * np;
  o.JQ[sDPg] += -gcgd * np;
  o.JQ[sDPb] += -gcbd * np;
  o.JQ[sDPdp] += (+gcbd + gcgd) * np;
  o.JQ[sSPg] -= gcgs * np;
  o.JQ[sSPb] -= +gcbs * np;
  o.JQ[sSPsp] += (+gcbs + gcgs) * np;
}

}  // namespace mos1
}  // namespace xb
