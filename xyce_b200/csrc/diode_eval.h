// xyce_b200 -- junction diode (Xyce level 1/2): one instance evaluation =
//   Instance::updateIntermediateVars   (src/DeviceModelPKG/OpenModels/N_DEV_Diode.C:1072-1420)
//   Master::updateState/loadDAEVectors/loadDAEMatrices   (N_DEV_Diode.C:1830-1950)
// restated for a one-thread-per-instance SoA kernel.  Nodes: 0 Pos, 1 Neg, 2 Pri (internal; aliases Pos
// when RS = 0).  Stamp slots (row, col): see kDiodeSlotRow/Col.  Store: vd, qd, cd.
#pragma once
#include "xb_common.h"
#include "simple_fields.def"

namespace xb {
namespace diode {

constexpr double kMaxExpArg = 100.0;                       // CONSTMAX_EXP_ARG
constexpr double kE = 2.718281828459045;                   // CONSTe = exp(1.0)
enum { kPos = 0, kNeg = 1, kPri = 2, kNodes = 3 };
// slots: PosPos, PosPri, NegNeg, NegPri, PriPos, PriNeg, PriPri
enum { sPP = 0, sPI, sNN, sNI, sIP, sIN, sII, kSlots };
XB_HD constexpr int slot_row(int s) { constexpr int t[kSlots] = {0, 0, 1, 1, 2, 2, 2}; return t[s]; }
XB_HD constexpr int slot_col(int s) { constexpr int t[kSlots] = {0, 2, 1, 2, 0, 1, 2}; return t[s]; }
enum { fBVGiven = 1, fJSWGiven = 2, fNSGiven = 4, fInitCondGiven = 8, fOff = 16 };

#define XB_D_DECL(n) double n;
struct Rec { XB_DIODE_FIELDS(XB_D_DECL, XB_D_DECL) };
#undef XB_D_DECL
#define XB_CNT(n) +1
constexpr int kNumFields = 0 XB_DIODE_FIELDS(XB_CNT, XB_CNT);
#undef XB_CNT

struct Out {
  real F[kNodes], Q[kNodes], FL[kNodes], QL[kNodes], JF[kSlots], JQ[kSlots];
  real Vd, Qd, Cd;      // store values
  int origFlag;
};

// vd_curr / vd_next: the store-vector entries li_storevd of the current and next store vectors
XB_HD void evaluate(const SolverFlags &S, const Rec &D, int flags, const real *V, real vd_curr, real vd_next, Out &o) {
  const real Vp = V[kPos], Vn = V[kNeg], Vpp = V[kPri];
  real Vd = Vpp - Vn;
  real Isat = D.tSatCur * D.Area;
  const real IsatSW = D.tSatSWCur * D.PJ;
  const real IsatR = D.tSatCurR * D.Area;
  const real Vt = kKoverQ * D.Temp;
  const real Vte = D.N * Vt, VteR = D.NR * Vt, VteSW = D.NS * Vt, VteBRK = D.NBV * Vt;
  real IdSW = 0.0, GdSW = 0.0;
  const real Gspr = D.tCOND * D.Area;
  const real Vd_orig = Vd;
  int origFlag = 1;
  real Vd_old;
  if (S.newtonIter == 0) {
    if (S.initJctFlag && S.voltageLimiterFlag) {
      if (flags & fInitCondGiven) { Vd = D.InitCond; origFlag = 0; }
      else if (flags & fOff) { Vd = 0.0; origFlag = 0; }
      else { Vd = D.tVcrit; origFlag = 0; }
    }
    Vd_old = Vd;
    if (!S.dcopFlag || (S.locaEnabledFlag && S.dcopFlag)) Vd_old = vd_curr;
  } else {
    Vd_old = vd_next;
  }
  if (S.voltageLimiterFlag) {
    int ichk = 0;
    if (S.newtonIter >= 0) {
      if ((flags & fBVGiven) && (Vd < dmin(0.0, -D.BV + 10.0 * VteBRK))) {
        real Vdtmp = -(D.BV + Vd);
        Vdtmp = pnjlim(Vdtmp, -(Vd_old + D.BV), VteBRK, D.tVcrit, ichk);
        Vd = -(Vdtmp + D.BV);
      } else {
        Vd = pnjlim(Vd, Vd_old, Vte, D.tVcrit, ichk);
      }
      if (ichk) origFlag = 0;
    }
  }
  if (flags & fJSWGiven) {
    if (flags & fNSGiven) {
      if (Vd >= -3 * VteSW) {
        real arg1 = dmin(kMaxExpArg, Vd / VteSW);
        const real evd = exp(arg1);
        IdSW = IsatSW * (evd - 1.0);
        GdSW = IsatSW * evd / VteSW;
      } else if (!(D.tBrkdwnV != 0.0) || (Vd >= -D.tBrkdwnV)) {
        real argsw = 3 * VteSW / (Vd * kE);
        argsw = argsw * argsw * argsw;
        IdSW = -IsatSW * (1 + argsw);
        GdSW = IsatSW * 3 * argsw / Vd;
      } else {
        real arg1 = dmin(kMaxExpArg, -(D.tBrkdwnV + Vd) / VteBRK);
        const real evrev = exp(arg1);
        IdSW = -IsatSW * evrev;
        GdSW = IsatSW * evrev / VteBRK;
      }
    } else {
      Isat = Isat + IsatSW;
    }
  }
  real Id, Gd;
  if (Vd >= -3.0 * Vte) {
    real arg1 = dmin(kMaxExpArg, Vd / Vte);
    real evd = exp(arg1);
    const real Inorm = Isat * (evd - 1.0) + S.gmin * Vd;
    const real Gd1 = Isat * evd / Vte + S.gmin;
    arg1 = dmin(kMaxExpArg, Vd / VteR);
    evd = exp(arg1);
    const real Irec = IsatR * (evd - 1.0);
    const real Gd2 = IsatR * evd / VteR;
    real Khi = 1, DKhi = 0;
    if (D.tIKF > 0) {
      Khi = sqrt(D.tIKF / (D.tIKF + Inorm));
      DKhi = -0.5 * Khi * Gd1 / (D.tIKF + Inorm);
    }
    real Kgen = 0, DKgen = 0;
    if (Irec != 0) {
      const real a = 1 - Vd / D.tJctPot;
      Kgen = sqrt(rpow(a * a + 0.005, D.M));
      DKgen = -D.M * a * Kgen / (D.tJctPot * (a * a + 0.005));
    }
    Id = Inorm * Khi + Irec * Kgen + IdSW;
    Gd = Gd1 * Khi + Inorm * DKhi + Gd2 * Kgen + Irec * DKgen + GdSW;
  } else if (!(D.tBrkdwnV != 0.0) || (Vd >= -D.tBrkdwnV)) {
    real arg = 3.0 * Vte / (Vd * kE);
    arg = arg * arg * arg;
    Id = -Isat * (1.0 + arg) + IdSW + S.gmin * Vd;
    Gd = Isat * 3.0 * arg / Vd + GdSW + S.gmin;
  } else {
    real arg1 = dmin(kMaxExpArg, -(D.tBrkdwnV + Vd) / VteBRK);
    const real evrev = exp(arg1);
    Id = -Isat * evrev + IdSW + S.gmin * Vd;
    Gd = Isat * evrev / VteBRK + GdSW + S.gmin;
  }
  const real Vc = Vd;
  real Qd, Cd;
  if (D.tJctCap != 0.0) {
    const real Czero = D.tJctCap * D.Area;
    if (Vc < D.tDepCap) {
      const real arg = 1.0 - Vc / D.tJctPot;
      real arg1 = dmin(kMaxExpArg, -D.M * log(arg));
      const real sarg = exp(arg1);
      Qd = D.TT * Id + D.tJctPot * Czero * (1.0 - arg * sarg) / (1.0 - D.M);
      Cd = D.TT * Gd + Czero * sarg;
    } else {
      const real Czof2 = Czero / D.F2;
      const real MotJctPot = D.M / D.tJctPot;
      Qd = D.TT * Id + Czero * D.tF1 + Czof2 * (D.F3 * (Vc - D.tDepCap) + (0.5 * MotJctPot) * (Vc * Vc - D.tDepCap * D.tDepCap));
      Cd = D.TT * Gd + Czof2 * (D.F3 + MotJctPot * Vc);
    }
    const real CzeroSW = D.tJctSWCap * D.PJ;
    if (Vc < D.tDepSWCap) {
      const real argSW = 1.0 - Vc / D.tJctSWPot;
      real argSW1 = dmin(kMaxExpArg, -D.MJSW * log(argSW));
      const real sargSW = exp(argSW1);
      Qd += D.tJctSWPot * CzeroSW * (1.0 - argSW * sargSW) / (1.0 - D.MJSW);
      Cd += CzeroSW * sargSW;
    } else {
      const real Czof2SW = CzeroSW / D.F2SW;
      const real MotJctSWPot = D.MJSW / D.tJctSWPot;
      Qd += CzeroSW * D.tF1 + Czof2SW * (D.F3SW * (Vc - D.tDepSWCap) + (0.5 * MotJctSWPot) * (Vc * Vc - D.tDepSWCap * D.tDepSWCap));
      Cd += Czof2SW * (D.F3SW + MotJctSWPot * Vc);
    }
  } else {
    Qd = 0.0;
    Cd = 0.0;
  }

  // ---- loads (Master::loadDAEVectors / loadDAEMatrices) ----
  const real mf = D.multiplicityFactor;
  const real Ir = Gspr * (Vp - Vpp);
  for (int i = 0; i < kNodes; ++i) o.F[i] = o.Q[i] = o.FL[i] = o.QL[i] = 0.0;
  o.F[kPos] -= -Ir * mf;
  o.F[kNeg] -= Id * mf;
  o.F[kPri] -= (-Id + Ir) * mf;
  o.Q[kNeg] -= Qd * mf;
  o.Q[kPri] -= -Qd * mf;
  if (S.voltageLimiterFlag) {
    const real Vd_diff = Vd - Vd_orig;
    const real Cd_Jdxp = -(Cd)*Vd_diff * mf;
    const real Gd_Jdxp = -(Gd)*Vd_diff * mf;
    o.FL[kNeg] += Gd_Jdxp * mf;     // (the reference applies the multiplicity twice here)
    o.FL[kPri] -= Gd_Jdxp * mf;
    o.QL[kNeg] += Cd_Jdxp * mf;
    o.QL[kPri] -= Cd_Jdxp * mf;
  }
  for (int s = 0; s < kSlots; ++s) o.JF[s] = o.JQ[s] = 0.0;
  o.JF[sPP] += Gspr * mf;
  o.JF[sPI] -= Gspr * mf;
  o.JF[sNN] += Gd * mf;
  o.JF[sNI] -= Gd * mf;
  o.JF[sIP] -= Gspr * mf;
  o.JF[sIN] -= Gd * mf;
  o.JF[sII] += (Gspr + Gd) * mf;
  o.JQ[sNN] += Cd * mf;
  o.JQ[sNI] -= Cd * mf;
  o.JQ[sIN] -= Cd * mf;
  o.JQ[sII] += Cd * mf;
  o.Vd = Vd; o.Qd = Qd; o.Cd = Cd;
  o.origFlag = origFlag;
}

}  // namespace diode
}  // namespace xb
