// xyce_b200 -- Gummel-Poon BJT: one instance evaluation =
//   Instance::updateIntermediateVars + auxDAECalculations   (src/DeviceModelPKG/OpenModels/N_DEV_BJT.C:2810-3434, :2139-2198)
//   Master::updateState / loadDAEVectors / loadDAEMatrices   (N_DEV_BJT.C:4112-4160, :4199-4398, :4400-4520)
//   Instance::oldDAEExcessPhaseCalculation1 / 2              (N_DEV_BJT.C:2706-2799: Weil's approximation of the excess
//                                                             phase, model PTF != 0, history in the store entry CEXBC)
// restated for a one-thread-per-instance SoA kernel.  Scope: DeviceOptions::newExcessPhase = false (the build
// default, N_DEV_DeviceOptions.C:55-59; the other formulation adds two unknowns per instance).
// Nodes: 0 Coll, 1 Base, 2 Emit, 3 Subst, 4 Coll' , 5 Base', 6 Emit' (primed nodes alias the external ones
// when RC / RB / RE = 0).  Store: vBE vBC capeqCB cexbc.  State: qBEdiff qBEdep qCS qBCdiff qBCdep qBX.
#pragma once
#include "xb_common.h"
#include "simple_fields.def"

namespace xb {
namespace bjt {

constexpr double kMaxExpArg = 100.0;    // CONSTMAX_EXP_ARG
enum { kC = 0, kB, kE, kS, kCP, kBP, kEP, kNodes };
// jacStamp_RB_RC_RE_ (N_DEV_BJT.C:1146-1182), row-major
enum { sCc = 0, sCcp, sBb, sBcp, sBbp, sBep, sEe, sEep, sSs, sScp, sCPc, sCPb, sCPs, sCPcp, sCPbp, sCPep,
       sBPb, sBPcp, sBPbp, sBPep, sEPe, sEPcp, sEPbp, sEPep, kSlots };
XB_HD constexpr int slot_row(int s) {
  constexpr int t[kSlots] = {0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6};
  return t[s];
}
XB_HD constexpr int slot_col(int s) {
  constexpr int t[kSlots] = {0, 4, 1, 4, 5, 6, 2, 6, 3, 4, 0, 1, 3, 4, 5, 6, 1, 4, 5, 6, 2, 4, 5, 6};
  return t[s];
}
enum { fIcGiven = 1, fOff = 2 };
enum { st_vBE = 0, st_vBC, st_capeqCB, st_cexbc, kNumStore };
enum { sa_qBEdiff = 0, sa_qBEdep, sa_qCS, sa_qBCdiff, sa_qBCdep, sa_qBX, kNumState };

#define XB_Q_DECL(n) double n;
struct Rec { XB_BJT_FIELDS(XB_Q_DECL, XB_Q_DECL) };
#undef XB_Q_DECL
#define XB_CNT(n) +1
constexpr int kNumFields = 0 XB_BJT_FIELDS(XB_CNT, XB_CNT);
#undef XB_CNT

struct Out {
  real F[kNodes], Q[kNodes], FL[kNodes], QL[kNodes], JF[kSlots], JQ[kSlots];
  real store[3], state[kNumState];
  int origFlag;
  // excess phase: bit 0 = cexbc_next goes to the next store, bit 1 = cexbc_init goes to the current AND the last store
  // (no history yet: first step out of a break point, N_DEV_BJT.C:2716-2721)
  int cexbc_mode;
  real cexbc_next, cexbc_init;
  // lead currents (Master::loadDAEVectors, N_DEV_BJT.C:4358-4378), branch order ib, ie, ic, is (registerBranchDataLIDs :1548-1551)
  real leadF[4], leadQ[4];
};

// depletion charge / capacitance of one junction (the four copies at N_DEV_BJT.C:3223-3330)
XB_HD void depletion(real v, real cz, real fcp, real pot, real mj, real f1, real fa, real fb, real &q, real &cap) {
  if (cz == 0.0) { q = 0.0; cap = 0.0; return; }
  if (v < fcp) {
    const real arg = 1.0 - v / pot;
    const real sarg = exp(-mj * log(arg));
    q = pot * cz * (1.0 - arg * sarg) / (1.0 - mj);
    cap = cz * sarg;
  } else {
    const real czf = cz / fa;
    q = cz * f1 + czf * (fb * (v - fcp) + (mj / (2.0 * pot)) * (v * v - fcp * fcp));
    cap = czf * (fb + mj * v / pot);
  }
}

XB_HD void junction(real v, real csat, real vte, real ileak, real vtleak, real gmin, real &i, real &g, real &il, real &gl) {
  if (v > -5.0 * vte) {
    const real ev = exp(dmin(kMaxExpArg, v / vte));
    i = csat * (ev - 1.0) + gmin * v;
    g = csat * ev / vte + gmin;
    if (ileak == 0.0) { il = gl = 0.0; }
    else {
      const real evl = exp(dmin(kMaxExpArg, v / vtleak));
      il = ileak * (evl - 1.0);
      gl = ileak * evl / vtleak;
    }
  } else {
    g = -csat / v + gmin;
    i = g * v;
    gl = -ileak / v;
    il = gl * v;
  }
}

XB_HD void evaluate(const SolverFlags &S, const Rec &M, int flags, const real *V, const real *curr_sto,
                    const real *next_sto, Out &o, real cexbc_curr = 0.0, real cexbc_last = 0.0) {
  const real ty = M.TYPE;
  const real AREA = M.AREA, vt = M.vt;
  const real vEEp = V[kE] - V[kEP], vBBp = V[kB] - V[kBP], vCCp = V[kC] - V[kCP];
  real vBE = ty * (V[kBP] - V[kEP]);
  real vBC = ty * (V[kBP] - V[kCP]);
  real vBX = ty * (V[kB] - V[kCP]);
  const real vCS = ty * (V[kS] - V[kCP]);
  const real vBE_orig = vBE, vBC_orig = vBC;
  int origFlag = 1, offFlag = 0;
  const bool OFF = (flags & fOff) != 0;
  if (S.initJctFlag && !OFF && S.voltageLimiterFlag) {
    if (flags & fIcGiven) {
      vBE = ty * M.icVBE;
      const real vCE = ty * M.icVCE;
      vBC = vBE - vCE;
      vBX = vBC;
      origFlag = 0;
    } else {       // (inputOPFlag is rejected by the C ABI)
      vBE = M.tVCrit;
      vBX = vBC;
      origFlag = 0;
    }
  } else if ((S.initFixFlag || S.initJctFlag) && OFF) {
    vBX = vBC = vBE = 0.0;
    offFlag = 1;
  }
  real vBE_old, vBC_old, capeqCB;
  if (S.newtonIter == 0) {
    if (!S.dcopFlag || (S.locaEnabledFlag && S.dcopFlag)) {
      vBE_old = curr_sto[st_vBE]; vBC_old = curr_sto[st_vBC]; capeqCB = curr_sto[st_capeqCB];
    } else {
      vBE_old = vBE; vBC_old = vBC; capeqCB = 0.0;
    }
  } else {
    vBE_old = next_sto[st_vBE]; vBC_old = next_sto[st_vBC]; capeqCB = next_sto[st_capeqCB];
  }
  if (S.voltageLimiterFlag && !(S.initFixFlag && OFF)) {
    if (S.newtonIter >= 0) {
      int icheck = 0, ichk1 = 1;
      vBE = pnjlim(vBE, vBE_old, vt, M.tVCrit, icheck);
      vBC = pnjlim(vBC, vBC_old, vt, M.tVCrit, ichk1);
      if (ichk1 == 1) icheck = 1;
      if (icheck == 1) origFlag = 0;
    }
  }

  // junction currents
  const real csat = M.tSatCur * AREA;
  const real vtF = vt * M.emissionCoeffF, vtR = vt * M.emissionCoeffR;
  const real vtE = vt * M.tleakBEEmissionCoeff, vtC = vt * M.tleakBCEmissionCoeff;
  const real iLeakBE = M.tBELeakCur * AREA, iLeakBC = M.tBCLeakCur * AREA;
  real iBE, gBE, iBEleak, gBEleak, iBC, gBC, iBCleak, gBCleak;
  junction(vBE, csat, vtF, iLeakBE, vtE, S.gmin, iBE, gBE, iBEleak, gBEleak);
  junction(vBC, csat, vtR, iLeakBC, vtC, S.gmin, iBC, gBC, iBCleak, gBCleak);

  // base charge
  const real rofF = M.tInvRollOffF / AREA, rofR = M.tInvRollOffR / AREA;
  const real q1 = 1.0 / (1.0 - M.tInvEarlyVoltF * vBC - M.tInvEarlyVoltR * vBE);
  real qB, dqBdvEp, dqBdvCp;
  if (rofF == 0.0 && rofR == 0.0) {
    qB = q1;
    const real q1_qB = q1 * qB;
    dqBdvEp = q1_qB * M.tInvEarlyVoltR;
    dqBdvCp = q1_qB * M.tInvEarlyVoltF;
  } else {
    const real q2 = rofF * iBE + rofR * iBC;
    if (q2 >= 0) {
      const real arg = 1.0 + 4.0 * q2;
      real sqarg = 1.0;
      if (fabs(arg) > 0.0) sqarg = rpow(arg, M.tRollOffExp);
      real rofF_gBE_invSqarg = 0.0, rofR_gBC_invSqarg = 0.0;
      if (arg != 0) rofF_gBE_invSqarg = rofF * gBE * 2 * M.tRollOffExp * sqarg / arg;
      if (arg != 0) rofR_gBC_invSqarg = rofR * gBC * 2 * M.tRollOffExp * sqarg / arg;
      qB = 0.5 * q1 * (1.0 + sqarg);
      dqBdvEp = q1 * (qB * M.tInvEarlyVoltR + rofF_gBE_invSqarg);
      dqBdvCp = q1 * (qB * M.tInvEarlyVoltF + rofR_gBC_invSqarg);
    } else {
      qB = q1;
      dqBdvEp = q1 * qB * M.tInvEarlyVoltR;
      dqBdvCp = q1 * qB * M.tInvEarlyVoltF;
    }
  }
  const real dqBdvBp = -(dqBdvEp + dqBdvCp);
  const real invqB = 1.0 / qB;

  // charges
  const real ctot = M.tBCCap * AREA;
  const real czBC = ctot * M.baseFracBCCap;
  const real czBX = ctot - czBC;
  const real czBE = M.tBECap * AREA;
  const real czCS = M.CJS * AREA;
  const real fcpc = M.depCapCoeff * M.potBC;
  const real fcpe = M.tDepCap;
  real iBEhighCurr = iBE, gBEhighCurr = gBE;
  if (M.transTimeF != 0.0 && vBE > 0.0 && ((!S.dcopFlag) || S.tranopFlag || S.acopFlag)) {
    real argtf = 0.0, arg2 = 0.0, arg3 = 0.0;
    if (M.transTimeBiasCoeffF != 0.0) {
      argtf = M.transTimeBiasCoeffF;
      if (M.transTimeVBCFac != 0.0) argtf *= exp(vBC * M.transTimeVBCFac);
      arg2 = argtf;
      if (M.transTimeHighCurrF != 0.0) {
        const real tmp = iBEhighCurr / (iBEhighCurr + M.transTimeHighCurrF * AREA);
        argtf *= tmp * tmp;
        arg2 = argtf * (3.0 - 2.0 * tmp);
      }
      arg3 = iBEhighCurr * argtf * M.transTimeVBCFac;
    }
    iBEhighCurr *= (1.0 + argtf) / qB;
    gBEhighCurr = (gBEhighCurr * (1.0 + arg2) - iBEhighCurr * dqBdvEp) / qB;
    capeqCB = M.transTimeF * (arg3 - iBEhighCurr * dqBdvCp) / qB;
  }
  real qBEdep, capBEdep, qBCdep, capBCdep, qBX, capBX;
  depletion(vBE, czBE, fcpe, M.tBEPot, M.juncExpBE, M.tF1, M.f2, M.f3, qBEdep, capBEdep);
  real qBEdiff, capBEdiff;
  if (M.transTimeF == 0.0) { qBEdiff = capBEdiff = 0.0; }
  else { qBEdiff = M.transTimeF * iBEhighCurr; capBEdiff = M.transTimeF * gBEhighCurr; }
  depletion(vBC, czBC, fcpc, M.tBCPot, M.juncExpBC, M.tF5, M.f6, M.f7, qBCdep, capBCdep);
  real qBCdiff, capBCdiff;
  if (M.transTimeR == 0.0) { qBCdiff = capBCdiff = 0.0; }
  else { qBCdiff = M.transTimeR * iBC; capBCdiff = M.transTimeR * gBC; }
  depletion(vBX, czBX, fcpc, M.tBCPot, M.juncExpBC, M.tF5, M.f6, M.f7, qBX, capBX);
  real qCS, capCS;
  if (czCS == 0.0) { qCS = 0.0; capCS = 0.0; }
  else if (vCS < 0.0) {
    const real arg = 1.0 - vCS / M.potSubst;
    const real sarg = exp(-M.expSubst * log(arg));
    qCS = M.potSubst * czCS * (1.0 - arg * sarg) / (1.0 - M.expSubst);
    capCS = czCS * sarg;
  } else {
    qCS = vCS * czCS * (1.0 + M.expSubst * vCS / (2.0 * M.potSubst));
    capCS = czCS * (1.0 + M.expSubst * vCS / M.potSubst);
  }

  // terminal currents; excess phase (oldDAEExcessPhaseCalculation1 / 2): in transient with td != 0 the collector current
  // follows iBE / qB through a second-order Bessel response integrated with backward Euler on the step history
  real iEX = iBE, gEX = gBE, iC_local = 0.0;
  o.cexbc_mode = 0; o.cexbc_next = o.cexbc_init = 0.0;
  const real td = M.excessPhaseFac;
  if (!S.dcopFlag && td != 0.0) {
    const real dt0 = S.currTimeStep, dt1 = S.lastTimeStep;
    real arg1 = dt0 / td;
    const real arg2 = 3.0 * arg1;
    arg1 = arg2 * arg1;
    const real denom = 1.0 + arg1 + arg2;
    const real phaseScalar = arg1 / denom;
    real currCexbc = cexbc_curr, lastCexbc = cexbc_last;
    if (S.beginIntegrationFlag) {
      currCexbc = lastCexbc = iBE / qB;
      o.cexbc_init = currCexbc; o.cexbc_mode |= 2;
    }
    iC_local = ((currCexbc) * (1 + dt0 / dt1 + arg2) - (lastCexbc) * dt0 / dt1) / denom;
    iEX = iBE * phaseScalar;
    gEX = gBE * phaseScalar;
    o.cexbc_next = iC_local + iEX / qB; o.cexbc_mode |= 1;
  }
  const real iCE = (iEX - iBC) / qB;
  const real iC = iC_local + (iCE - iBC / M.tBetaR - iBCleak);
  const real iB = iBE / M.tBetaF + iBEleak + iBC / M.tBetaR + iBCleak;
  const real iE = -iC - iB;
  const real diCEdvEp = invqB * (iCE * dqBdvEp - gEX);
  const real diCEdvCp = invqB * (iCE * dqBdvCp + gBC);
  const real diCEdvBp = invqB * (iCE * dqBdvBp + gEX - gBC);
  const real gEpr = M.emitterConduct * AREA, gCpr = M.collectorConduct * AREA;
  const real rBpr = M.minBaseResist / AREA;
  const real rBpi = M.baseResist / AREA - rBpr;
  const real xjrB = M.baseCurrHalfResist * AREA;
  real gX = rBpr + rBpi / qB;
  if (fabs(xjrB) > 0.0) {
    real arg1 = dmax(iB / xjrB, 1.0e-09);
    const real arg2 = (-1.0 + sqrt(1.0 + 14.59025 * arg1)) / 2.4317 / sqrt(arg1);
    arg1 = rtan(arg2);
    gX = rBpr + 3.0 * rBpi * (arg1 - arg2) / arg2 / arg1 / arg1;
  }
  if (fabs(gX) > 0.0) gX = 1.0 / gX;
  const real diBrdvB = gX, diBrdvCp = 0.0, diBrdvEp = 0.0, diBrdvBp = -gX;
  const real gBEtot = gBE / M.tBetaF + gBEleak;
  const real gBCtot = gBC / M.tBetaR + gBCleak;

  // ---- Master::updateState ----
  o.store[st_vBE] = vBE; o.store[st_vBC] = vBC; o.store[st_capeqCB] = capeqCB;
  o.state[sa_qBEdiff] = qBEdiff; o.state[sa_qBEdep] = qBEdep; o.state[sa_qCS] = qCS;
  o.state[sa_qBCdiff] = qBCdiff; o.state[sa_qBCdep] = qBCdep; o.state[sa_qBX] = qBX;
  o.origFlag = origFlag;

  // ---- Master::loadDAEVectors ----
  const real mf = M.multiplicityFactor;
  for (int i = 0; i < kNodes; ++i) o.F[i] = o.Q[i] = o.FL[i] = o.QL[i] = 0.0;
  const real vbe_diff = vBE - vBE_orig, vbc_diff = vBC - vBC_orig;
  const real vce_diff = vbe_diff - vbc_diff;
  o.F[kC] -= -vCCp * gCpr * mf;
  o.F[kB] -= -vBBp * gX * mf;
  o.F[kE] -= -vEEp * gEpr * mf;
  o.F[kCP] -= (vCCp * gCpr + ty * (-iC)) * mf;
  o.F[kBP] -= (vBBp * gX - ty * (iB)) * mf;
  o.F[kEP] -= (vEEp * gEpr + ty * (-iE)) * mf;
  o.Q[kB] -= -ty * qBX * mf;
  o.Q[kS] -= -ty * qCS * mf;
  o.Q[kCP] -= ty * (qCS + qBX + qBCdep + qBCdiff) * mf;
  o.Q[kBP] -= -ty * (qBEdep + qBEdiff + qBCdep + qBCdiff) * mf;
  o.Q[kEP] -= ty * (qBEdep + qBEdiff) * mf;
  if (S.voltageLimiterFlag && (!origFlag || offFlag)) {
    real c = +diCEdvBp * vbe_diff + diCEdvCp * vce_diff - gBCtot * vbc_diff;
    real b = gBEtot * vbe_diff + gBCtot * vbc_diff;
    real e = -diCEdvCp * vce_diff - (diCEdvBp + gBEtot) * vbe_diff;
    o.FL[kCP] += c * ty * mf;
    o.FL[kBP] += b * ty * mf;
    o.FL[kEP] += e * ty * mf;
    c = -(capBCdep + capBCdiff) * vbc_diff;
    b = (capBEdep + capBEdiff) * vbe_diff + (capBCdiff + capBCdep + capeqCB) * vbc_diff;
    e = -capeqCB * vbc_diff - (capBEdiff + capBEdep) * vbe_diff;
    o.QL[kCP] += c * ty * mf;
    o.QL[kBP] += b * ty * mf;
    o.QL[kEP] += e * ty * mf;
  }

  o.leadQ[2] = -ty * (qCS + qBX + qBCdep + qBCdiff) * mf;
  o.leadQ[0] = ty * (qBX + qBEdep + qBEdiff + qBCdep + qBCdiff) * mf;
  o.leadQ[1] = -ty * (qBEdep + qBEdiff) * mf;
  o.leadQ[3] = ty * qCS * mf;
  o.leadF[2] = ty * (iC) * mf;
  o.leadF[3] = 0.0;
  o.leadF[1] = ty * (iE) * mf;
  o.leadF[0] = ty * (iB) * mf;

  // ---- Master::loadDAEMatrices ----
  for (int s = 0; s < kSlots; ++s) o.JF[s] = o.JQ[s] = 0.0;
  o.JF[sCc] += gCpr * mf;
  o.JF[sCcp] -= gCpr * mf;
  o.JF[sBb] += diBrdvB * mf;
  o.JF[sBcp] += diBrdvCp * mf;
  o.JF[sBbp] += diBrdvBp * mf;
  o.JF[sBep] += diBrdvEp * mf;
  o.JF[sEe] += gEpr * mf;
  o.JF[sEep] -= gEpr * mf;
  o.JF[sCPc] -= gCpr * mf;
  o.JF[sCPcp] += (diCEdvCp + gBCtot + gCpr) * mf;
  o.JF[sCPbp] += (diCEdvBp - gBCtot) * mf;
  o.JF[sCPep] += diCEdvEp * mf;
  o.JF[sBPb] -= diBrdvB * mf;
  o.JF[sBPcp] += (-diBrdvCp - gBCtot) * mf;
  o.JF[sBPbp] += (-diBrdvBp + gBEtot + gBCtot) * mf;
  o.JF[sBPep] += (-diBrdvEp - gBEtot) * mf;
  o.JF[sEPe] -= gEpr * mf;
  o.JF[sEPcp] += -diCEdvCp * mf;
  o.JF[sEPbp] += (-diCEdvBp - gBEtot) * mf;
  o.JF[sEPep] += (gBEtot + gEpr + diCEdvBp + diCEdvCp) * mf;
  o.JQ[sBb] += capBX * mf;
  o.JQ[sBcp] += -capBX * mf;
  o.JQ[sSs] += capCS * mf;
  o.JQ[sScp] -= capCS * mf;
  o.JQ[sCPb] -= capBX * mf;
  o.JQ[sCPs] -= capCS * mf;
  o.JQ[sCPcp] += (capCS + capBX + capBCdep + capBCdiff) * mf;
  o.JQ[sCPbp] += (-capBCdep - capBCdiff) * mf;
  o.JQ[sBPcp] += (-capBCdiff - capBCdep - capeqCB) * mf;
  o.JQ[sBPbp] += (capBEdiff + capBEdep + capBCdiff + capBCdep + capeqCB) * mf;
  o.JQ[sBPep] += (-capBEdiff - capBEdep) * mf;
  o.JQ[sEPcp] += capeqCB * mf;
  o.JQ[sEPbp] += (-capBEdiff - capBEdep - capeqCB) * mf;
  o.JQ[sEPep] += (capBEdiff + capBEdep) * mf;
}

}  // namespace bjt
}  // namespace xb
