// xyce_b200 -- BSIM4 v4.8.2 charge / capacitance stage: intrinsic C-V for
// capMod 0, 1, 2 (incl. charge-thickness model), NQS conductances, junction
// and overlap capacitances.  See bsim4_eval.h for the contract.
// Behavioural specification: N_DEV_MOSFET_B4p82.C:5843-7035.
#pragma once

namespace xb {
namespace b4 {


// Depletion charge / capacitance of one bulk junction: bottom, sidewall and
// gate-side sidewall components (B4p82.C:6800-6928).
XB_HD void jct_comp(real v, real cz, real mj, real phib, real &q, real &cap, bool first) {
  if (cz > 0.0) {
    const real arg = 1.0 - v / phib;
    const real sarg = (mj == 0.5) ? 1.0 / sqrt(arg) : exp(-mj * log(arg));
    const real dq = phib * cz * (1.0 - arg * sarg) / (1.0 - mj);
    if (first) { q = dq; cap = cz * sarg; }
    else { q += dq; cap += cz * sarg; }
  } else if (first) {
    q = 0.0;
    cap = 0.0;
  }
}
XB_HD void junction_charge(real v, real cz, real czsw, real czswg,
                           real mj, real mjsw, real mjswg,
                           real phib, real phibsw, real phibswg,
                           real &q, real &cap) {
  if (v == 0.0) {
    q = 0.0;
    cap = cz + czsw + czswg;
  } else if (v < 0.0) {
    jct_comp(v, cz, mj, phib, q, cap, true);
    jct_comp(v, czsw, mjsw, phibsw, q, cap, false);
    jct_comp(v, czswg, mjswg, phibswg, q, cap, false);
  } else {
    const real T0 = cz + czsw + czswg;
    const real T1 = v * (cz * mj / phib + czsw * mjsw / phibsw + czswg * mjswg / phibswg);
    q = v * (T0 + 0.5 * T1);
    cap = T0 + T1;
  }
}

XB_HELPER Real2 junction_charge_v(real v, real cz, real czsw, real czswg, real mj, real mjsw, real mjswg,
                                  real phib, real phibsw, real phibswg) {
  Real2 r; junction_charge(v, cz, czsw, czswg, mj, mjsw, mjswg, phib, phibsw, phibswg, r.a, r.b); return r;
}

XB_HD void stage_cv(const SolverFlags &S, const B4Model &M, const B4Size &P,
                    const B4Inst &I, B4Mid &W, DcCarry &C) {
  real T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, tmp, tmp1;
  const bool charge_needed = S.tranopFlag || S.acopFlag || S.transientFlag || S.dcsweepFlag;
  W.ChargeComputationNeeded = charge_needed ? 1 : 0;
  const bool kV47 = M.versionDouble >= 4.7;      // 4.6.1 differences of the capMod = 0 branch

  const real Vds = C.Vds, Vbs = C.Vbs, Vbseff = C.Vbseff, dVbseff_dVb = C.dVbseff_dVb;
  const real Phis = C.Phis, dPhis_dVb = C.dPhis_dVb, sqrtPhis = C.sqrtPhis, dsqrtPhis_dVb = C.dsqrtPhis_dVb;
  const real Vgs_eff = C.Vgs_eff, dVgs_eff_dVg = C.dVgs_eff_dVg;
  const real n = C.n, dn_dVb = C.dn_dVb, dn_dVd = C.dn_dVd, Vtm = C.Vtm;
  const real Abulk0 = C.Abulk0, dAbulk0_dVb = C.dAbulk0_dVb;
  const real epssub = C.epssub, toxe = C.toxe;
  real Vth = C.Vth, dVth_dVb = C.dVth_dVb, dVth_dVd = C.dVth_dVd, Vgst = C.Vgst;
  real Vgsteff = C.Vgsteff, dVgsteff_dVg = C.dVgsteff_dVg, dVgsteff_dVd = C.dVgsteff_dVd,
         dVgsteff_dVb = C.dVgsteff_dVb;
  real Vdsat = C.Vdsat;
  real VbseffCV, dVbseffCV_dVb, Vfb, CoxWL;
  real qgate = 0, qbulk = 0, qdrn = 0, qsrc = 0;

  if ((M.xpart < 0) || (!charge_needed)) {
    qgate = qdrn = qsrc = qbulk = 0.0;
    W.cggb = W.cgsb = W.cgdb = 0.0;
    W.cdgb = W.cdsb = W.cddb = 0.0;
    W.cbgb = W.cbsb = W.cbdb = 0.0;
    W.csgb = W.cssb = W.csdb = 0.0;
    W.cgbb = W.csbb = W.cdbb = W.cbbb = 0.0;
    W.cqdb = W.cqsb = W.cqgb = W.cqbb = 0.0;
    W.gtau = 0.0;
  } else if (M.capMod == 0) {
    if (Vbseff < 0.0) { VbseffCV = Vbs; dVbseffCV_dVb = 1.0; }
    else { VbseffCV = P.phi - Phis; dVbseffCV_dVb = kV47 ? -dPhis_dVb * dVbseff_dVb : -dPhis_dVb; }
    Vfb = P.vfbcv;
    Vth = Vfb + P.phi + P.k1ox * sqrtPhis;
    Vgst = Vgs_eff - Vth;
    dVth_dVb = kV47 ? P.k1ox * dsqrtPhis_dVb * dVbseff_dVb : P.k1ox * dsqrtPhis_dVb;
    CoxWL = M.coxe * P.weffCV * P.leffCV * I.nf;
    const real Arg1 = Vgs_eff - VbseffCV - Vfb;
    if (Arg1 <= 0.0) {            // accumulation
      qgate = CoxWL * Arg1;
      qbulk = -qgate;
      qdrn = 0.0;
      W.cggb = CoxWL * dVgs_eff_dVg;
      W.cgdb = 0.0;
      W.cgsb = CoxWL * (dVbseffCV_dVb - dVgs_eff_dVg);
      W.cdgb = 0.0; W.cddb = 0.0; W.cdsb = 0.0;
      W.cbgb = -CoxWL * dVgs_eff_dVg;
      W.cbdb = 0.0;
      W.cbsb = -W.cgsb;
    } else if (Vgst <= 0.0) {     // depletion
      T1 = 0.5 * P.k1ox;
      T2 = sqrt(T1 * T1 + Arg1);
      qgate = CoxWL * P.k1ox * (T2 - T1);
      qbulk = -qgate;
      qdrn = 0.0;
      T0 = CoxWL * T1 / T2;
      W.cggb = T0 * dVgs_eff_dVg;
      W.cgdb = 0.0;
      W.cgsb = T0 * (dVbseffCV_dVb - dVgs_eff_dVg);
      W.cdgb = 0.0; W.cddb = 0.0; W.cdsb = 0.0;
      W.cbgb = -W.cggb;
      W.cbdb = 0.0;
      W.cbsb = -W.cgsb;
    } else {                      // inversion
      const real One_Third_CoxWL = CoxWL / 3.0;
      const real Two_Third_CoxWL = 2.0 * One_Third_CoxWL;
      const real AbulkCV = Abulk0 * P.abulkCVfactor;
      real dAbulkCV_dVb, dVdsat_dVg, dVdsat_dVb;
      if (kV47) {
        dAbulkCV_dVb = P.abulkCVfactor * dAbulk0_dVb * dVbseff_dVb;
        dVdsat_dVg = 1.0 / AbulkCV;
        Vdsat = Vgst * dVdsat_dVg;
        dVdsat_dVb = -(Vdsat * dAbulkCV_dVb + dVth_dVb) * dVdsat_dVg;
      } else {      // N_DEV_MOSFET_B4p61.C capMod = 0 inversion branch
        dAbulkCV_dVb = P.abulkCVfactor * dAbulk0_dVb;
        Vdsat = Vgst / AbulkCV;
        dVdsat_dVg = dVgs_eff_dVg / AbulkCV;
        dVdsat_dVb = -(Vdsat * dAbulkCV_dVb + dVth_dVb) / AbulkCV;
      }
      real Alphaz, dAlphaz_dVg, dAlphaz_dVb;
      if (M.xpart > 0.5) {        // 0/100 partition
        if (Vdsat <= Vds) {
          T1 = Vdsat / 3.0;
          qgate = CoxWL * (Vgs_eff - Vfb - P.phi - T1);
          T2 = -Two_Third_CoxWL * Vgst;
          qbulk = -(qgate + T2);
          qdrn = 0.0;
          W.cggb = One_Third_CoxWL * (3.0 - dVdsat_dVg) * dVgs_eff_dVg;
          T2 = -One_Third_CoxWL * dVdsat_dVb;
          W.cgsb = -(W.cggb + T2);
          W.cgdb = 0.0;
          W.cdgb = 0.0; W.cddb = 0.0; W.cdsb = 0.0;
          W.cbgb = -(W.cggb - Two_Third_CoxWL * dVgs_eff_dVg);
          T3 = -(T2 + Two_Third_CoxWL * dVth_dVb);
          W.cbsb = -(W.cbgb + T3);
          W.cbdb = 0.0;
        } else {
          Alphaz = Vgst / Vdsat;
          T1 = 2.0 * Vdsat - Vds;
          T2 = Vds / (3.0 * T1);
          T3 = T2 * Vds;
          T9 = 0.25 * CoxWL;
          T4 = T9 * Alphaz;
          T7 = 2.0 * Vds - T1 - 3.0 * T3;
          T8 = T3 - T1 - 2.0 * Vds;
          qgate = CoxWL * (Vgs_eff - Vfb - P.phi - 0.5 * (Vds - T3));
          T10 = T4 * T8;
          qdrn = T4 * T7;
          qbulk = -(qgate + qdrn + T10);
          T5 = T3 / T1;
          W.cggb = CoxWL * (1.0 - T5 * dVdsat_dVg) * dVgs_eff_dVg;
          T11 = -CoxWL * T5 * dVdsat_dVb;
          W.cgdb = CoxWL * (T2 - 0.5 + 0.5 * T5);
          W.cgsb = -(W.cggb + T11 + W.cgdb);
          T6 = 1.0 / Vdsat;
          dAlphaz_dVg = T6 * (1.0 - Alphaz * dVdsat_dVg);
          dAlphaz_dVb = -T6 * (dVth_dVb + Alphaz * dVdsat_dVb);
          T7 = T9 * T7;
          T8 = T9 * T8;
          T9 = 2.0 * T4 * (1.0 - 3.0 * T5);
          W.cdgb = (T7 * dAlphaz_dVg - T9 * dVdsat_dVg) * dVgs_eff_dVg;
          T12 = T7 * dAlphaz_dVb - T9 * dVdsat_dVb;
          W.cddb = T4 * (3.0 - 6.0 * T2 - 3.0 * T5);
          W.cdsb = -(W.cdgb + T12 + W.cddb);
          T9 = 2.0 * T4 * (1.0 + T5);
          T10 = (T8 * dAlphaz_dVg - T9 * dVdsat_dVg) * dVgs_eff_dVg;
          T11 = T8 * dAlphaz_dVb - T9 * dVdsat_dVb;
          T12 = T4 * (2.0 * T2 + T5 - 1.0);
          T0 = -(T10 + T11 + T12);
          W.cbgb = -(W.cggb + W.cdgb + T10);
          W.cbdb = -(W.cgdb + W.cddb + T12);
          W.cbsb = -(W.cgsb + W.cdsb + T0);
        }
      } else if (M.xpart < 0.5) { // 40/60 partition
        if (Vds >= Vdsat) {
          T1 = Vdsat / 3.0;
          qgate = CoxWL * (Vgs_eff - Vfb - P.phi - T1);
          T2 = -Two_Third_CoxWL * Vgst;
          qbulk = -(qgate + T2);
          qdrn = 0.4 * T2;
          W.cggb = One_Third_CoxWL * (3.0 - dVdsat_dVg) * dVgs_eff_dVg;
          T2 = -One_Third_CoxWL * dVdsat_dVb;
          W.cgsb = -(W.cggb + T2);
          W.cgdb = 0.0;
          T3 = 0.4 * Two_Third_CoxWL;
          W.cdgb = -T3 * dVgs_eff_dVg;
          W.cddb = 0.0;
          T4 = T3 * dVth_dVb;
          W.cdsb = -(T4 + W.cdgb);
          W.cbgb = -(W.cggb - Two_Third_CoxWL * dVgs_eff_dVg);
          T3 = -(T2 + Two_Third_CoxWL * dVth_dVb);
          W.cbsb = -(W.cbgb + T3);
          W.cbdb = 0.0;
        } else {
          Alphaz = Vgst / Vdsat;
          T1 = 2.0 * Vdsat - Vds;
          T2 = Vds / (3.0 * T1);
          T3 = T2 * Vds;
          T9 = 0.25 * CoxWL;
          T4 = T9 * Alphaz;
          qgate = CoxWL * (Vgs_eff - Vfb - P.phi - 0.5 * (Vds - T3));
          T5 = T3 / T1;
          W.cggb = CoxWL * (1.0 - T5 * dVdsat_dVg) * dVgs_eff_dVg;
          tmp = -CoxWL * T5 * dVdsat_dVb;
          W.cgdb = CoxWL * (T2 - 0.5 + 0.5 * T5);
          W.cgsb = -(W.cggb + W.cgdb + tmp);
          T6 = 1.0 / Vdsat;
          dAlphaz_dVg = T6 * (1.0 - Alphaz * dVdsat_dVg);
          dAlphaz_dVb = -T6 * (dVth_dVb + Alphaz * dVdsat_dVb);
          T6 = 8.0 * Vdsat * Vdsat - 6.0 * Vdsat * Vds + 1.2 * Vds * Vds;
          T8 = T2 / T1;
          T7 = Vds - T1 - T8 * T6;
          qdrn = T4 * T7;
          T7 *= T9;
          tmp = T8 / T1;
          tmp1 = T4 * (2.0 - 4.0 * tmp * T6 + T8 * (16.0 * Vdsat - 6.0 * Vds));
          W.cdgb = (T7 * dAlphaz_dVg - tmp1 * dVdsat_dVg) * dVgs_eff_dVg;
          T10 = T7 * dAlphaz_dVb - tmp1 * dVdsat_dVb;
          W.cddb = T4 * (2.0 - (1.0 / (3.0 * T1 * T1) + 2.0 * tmp) * T6 + T8 * (6.0 * Vdsat - 2.4 * Vds));
          W.cdsb = -(W.cdgb + T10 + W.cddb);
          T7 = 2.0 * (T1 + T3);
          qbulk = -(qgate - T4 * T7);
          T7 *= T9;
          T0 = 4.0 * T4 * (1.0 - T5);
          T12 = kV47 ? (-T7 * dAlphaz_dVg - T0 * dVdsat_dVg) * dVgs_eff_dVg - W.cdgb
                     : (-T7 * dAlphaz_dVg - W.cdgb - T0 * dVdsat_dVg) * dVgs_eff_dVg;
          T11 = -T7 * dAlphaz_dVb - T10 - T0 * dVdsat_dVb;
          T10 = -4.0 * T4 * (T2 - 0.5 + 0.5 * T5) - W.cddb;
          tmp = -(T10 + T11 + T12);
          W.cbgb = -(W.cggb + W.cdgb + T12);
          W.cbdb = -(W.cgdb + W.cddb + T10);
          W.cbsb = -(W.cgsb + W.cdsb + tmp);
        }
      } else {                    // 50/50 partition
        if (Vds >= Vdsat) {
          T1 = Vdsat / 3.0;
          qgate = CoxWL * (Vgs_eff - Vfb - P.phi - T1);
          T2 = -Two_Third_CoxWL * Vgst;
          qbulk = -(qgate + T2);
          qdrn = 0.5 * T2;
          W.cggb = One_Third_CoxWL * (3.0 - dVdsat_dVg) * dVgs_eff_dVg;
          T2 = -One_Third_CoxWL * dVdsat_dVb;
          W.cgsb = -(W.cggb + T2);
          W.cgdb = 0.0;
          W.cdgb = -One_Third_CoxWL * dVgs_eff_dVg;
          W.cddb = 0.0;
          T4 = One_Third_CoxWL * dVth_dVb;
          W.cdsb = -(T4 + W.cdgb);
          W.cbgb = -(W.cggb - Two_Third_CoxWL * dVgs_eff_dVg);
          T3 = -(T2 + Two_Third_CoxWL * dVth_dVb);
          W.cbsb = -(W.cbgb + T3);
          W.cbdb = 0.0;
        } else {
          Alphaz = Vgst / Vdsat;
          T1 = 2.0 * Vdsat - Vds;
          T2 = Vds / (3.0 * T1);
          T3 = T2 * Vds;
          T9 = 0.25 * CoxWL;
          T4 = T9 * Alphaz;
          qgate = CoxWL * (Vgs_eff - Vfb - P.phi - 0.5 * (Vds - T3));
          T5 = T3 / T1;
          W.cggb = CoxWL * (1.0 - T5 * dVdsat_dVg) * dVgs_eff_dVg;
          tmp = -CoxWL * T5 * dVdsat_dVb;
          W.cgdb = CoxWL * (T2 - 0.5 + 0.5 * T5);
          W.cgsb = -(W.cggb + W.cgdb + tmp);
          T6 = 1.0 / Vdsat;
          dAlphaz_dVg = T6 * (1.0 - Alphaz * dVdsat_dVg);
          dAlphaz_dVb = -T6 * (dVth_dVb + Alphaz * dVdsat_dVb);
          T7 = T1 + T3;
          qdrn = -T4 * T7;
          qbulk = -(qgate + qdrn + qdrn);
          T7 *= T9;
          T0 = T4 * (2.0 * T5 - 2.0);
          W.cdgb = (T0 * dVdsat_dVg - T7 * dAlphaz_dVg) * dVgs_eff_dVg;
          T12 = T0 * dVdsat_dVb - T7 * dAlphaz_dVb;
          W.cddb = T4 * (1.0 - 2.0 * T2 - T5);
          W.cdsb = -(W.cdgb + T12 + W.cddb);
          W.cbgb = -(W.cggb + 2.0 * W.cdgb);
          W.cbdb = -(W.cgdb + 2.0 * W.cddb);
          W.cbsb = -(W.cgsb + 2.0 * W.cdsb);
        }
      }
    }
  } else {
    // ---- capMod 1 / 2 -----------------------------------------------------------------------------
    real dT0_dVg, dT0_dVd, dT0_dVb, dT1_dVg, dT1_dVd, dT1_dVb;
    real dT9_dVg, dT9_dVd, dT9_dVb, dT10_dVg, dT10_dVd, dT10_dVb;
    real Csg, Csd, Csb, Cgg, Cgd, Cgb, Cbg, Cbd, Cbb;
    if (Vbseff < 0.0) { VbseffCV = Vbseff; dVbseffCV_dVb = 1.0; }
    else { VbseffCV = P.phi - Phis; dVbseffCV_dVb = -dPhis_dVb; }
    CoxWL = M.coxe * P.weffCV * P.leffCV * I.nf;

    if (M.cvchargeMod == 0) {
      const real noff = n * P.noff;
      const real dnoff_dVd = P.noff * dn_dVd;
      const real dnoff_dVb = P.noff * dn_dVb;
      T0 = Vtm * noff;
      const real voffcv = P.voffcv;
      const real VgstNVt = (Vgst - voffcv) / T0;
      if (VgstNVt > kExpThr) {
        Vgsteff = Vgst - voffcv;
        dVgsteff_dVg = dVgs_eff_dVg;
        dVgsteff_dVd = -dVth_dVd;
        dVgsteff_dVb = -dVth_dVb;
      } else if (VgstNVt < -kExpThr) {
        Vgsteff = T0 * log(1.0 + kMinExp);
        dVgsteff_dVg = 0.0;
        dVgsteff_dVd = Vgsteff / noff;
        dVgsteff_dVb = dVgsteff_dVd * dnoff_dVb;
        dVgsteff_dVd *= dnoff_dVd;
      } else {
        const real ExpVgst = exp(VgstNVt);
        Vgsteff = T0 * log(1.0 + ExpVgst);
        dVgsteff_dVg = ExpVgst / (1.0 + ExpVgst);
        dVgsteff_dVd = -dVgsteff_dVg * (dVth_dVd + (Vgst - voffcv) / noff * dnoff_dVd) + Vgsteff / noff * dnoff_dVd;
        dVgsteff_dVb = -dVgsteff_dVg * (dVth_dVb + (Vgst - voffcv) / noff * dnoff_dVb) + Vgsteff / noff * dnoff_dVb;
        dVgsteff_dVg *= dVgs_eff_dVg;
      }
    } else {
      T0 = n * Vtm;
      T1 = P.mstarcv * Vgst;
      T2 = T1 / T0;
      if (T2 > kExpThr) {
        T10 = T1;
        dT10_dVg = P.mstarcv * dVgs_eff_dVg;
        dT10_dVd = -dVth_dVd * P.mstarcv;
        dT10_dVb = -dVth_dVb * P.mstarcv;
      } else if (T2 < -kExpThr) {
        T10 = Vtm * log(1.0 + kMinExp);
        dT10_dVg = 0.0;
        dT10_dVd = T10 * dn_dVd;
        dT10_dVb = T10 * dn_dVb;
        T10 *= n;
      } else {
        const real ExpVgst = exp(T2);
        T3 = Vtm * log(1.0 + ExpVgst);
        T10 = n * T3;
        dT10_dVg = P.mstarcv * ExpVgst / (1.0 + ExpVgst);
        dT10_dVb = T3 * dn_dVb - dT10_dVg * (dVth_dVb + Vgst * dn_dVb / n);
        dT10_dVd = T3 * dn_dVd - dT10_dVg * (dVth_dVd + Vgst * dn_dVd / n);
        dT10_dVg *= dVgs_eff_dVg;
      }
      T1 = P.voffcbncv - (1.0 - P.mstarcv) * Vgst;
      T2 = T1 / T0;
      if (T2 < -kExpThr) {
        T3 = M.coxe * kMinExp / P.cdep0;
        T9 = P.mstarcv + T3 * n;
        dT9_dVg = 0.0; dT9_dVd = dn_dVd * T3; dT9_dVb = dn_dVb * T3;
      } else if (T2 > kExpThr) {
        T3 = M.coxe * kMaxExp / P.cdep0;
        T9 = P.mstarcv + T3 * n;
        dT9_dVg = 0.0; dT9_dVd = dn_dVd * T3; dT9_dVb = dn_dVb * T3;
      } else {
        const real ExpVgst = exp(T2);
        T3 = M.coxe / P.cdep0;
        T4 = T3 * ExpVgst;
        T5 = T1 * T4 / T0;
        T9 = P.mstarcv + n * T4;
        dT9_dVg = T3 * (P.mstarcv - 1.0) * ExpVgst / Vtm;
        dT9_dVb = T4 * dn_dVb - dT9_dVg * dVth_dVb - T5 * dn_dVb;
        dT9_dVd = T4 * dn_dVd - dT9_dVg * dVth_dVd - T5 * dn_dVd;
        dT9_dVg *= dVgs_eff_dVg;
      }
      Vgsteff = T10 / T9;
      T11 = T9 * T9;
      dVgsteff_dVg = (T9 * dT10_dVg - T10 * dT9_dVg) / T11;
      dVgsteff_dVd = (T9 * dT10_dVd - T10 * dT9_dVd) / T11;
      dVgsteff_dVb = (T9 * dT10_dVb - T10 * dT9_dVb) / T11;
    }
    W.Vgsteff = Vgsteff;

    XB_SYNC_POINT_U(1);
    if (M.capMod == 1) {
      Vfb = I.vfbzb;
      const real V3 = Vfb - Vgs_eff + VbseffCV - kDelta3;
      if (Vfb <= 0.0) T0 = sqrt(V3 * V3 - 4.0 * kDelta3 * Vfb);
      else T0 = sqrt(V3 * V3 + 4.0 * kDelta3 * Vfb);
      T1 = 0.5 * (1.0 + V3 / T0);
      const real Vfbeff = Vfb - 0.5 * (V3 + T0);
      const real dVfbeff_dVg = T1 * dVgs_eff_dVg;
      const real dVfbeff_dVb = -T1 * dVbseffCV_dVb;
      const real Qac0 = CoxWL * (Vfbeff - Vfb);
      const real dQac0_dVg = CoxWL * dVfbeff_dVg;
      const real dQac0_dVb = CoxWL * dVfbeff_dVb;

      T0 = 0.5 * P.k1ox;
      T3 = Vgs_eff - Vfbeff - VbseffCV - Vgsteff;
      if (P.k1ox == 0.0) { T1 = 0.0; T2 = 0.0; }
      else if (T3 < 0.0) { T1 = T0 + T3 / P.k1ox; T2 = CoxWL; }
      else { T1 = sqrt(T0 * T0 + T3); T2 = CoxWL * T0 / T1; }
      const real Qsub0 = CoxWL * P.k1ox * (T1 - T0);
      const real dQsub0_dVg = T2 * (dVgs_eff_dVg - dVfbeff_dVg - dVgsteff_dVg);
      const real dQsub0_dVd = -T2 * dVgsteff_dVd;
      const real dQsub0_dVb = -T2 * (dVfbeff_dVb + dVbseffCV_dVb + dVgsteff_dVb);

      const real AbulkCV = Abulk0 * P.abulkCVfactor;
      const real dAbulkCV_dVb = P.abulkCVfactor * dAbulk0_dVb;
      const real VdsatCV = Vgsteff / AbulkCV;
      T0 = VdsatCV - Vds - kDelta4;
      dT0_dVg = 1.0 / AbulkCV;
      dT0_dVb = -VdsatCV * dAbulkCV_dVb / AbulkCV;
      T1 = sqrt(T0 * T0 + 4.0 * kDelta4 * VdsatCV);
      dT1_dVg = (T0 + kDelta4 + kDelta4) / T1;
      dT1_dVd = -T0 / T1;
      dT1_dVb = dT1_dVg * dT0_dVb;
      dT1_dVg *= dT0_dVg;
      real VdseffCV, dVdseffCV_dVg, dVdseffCV_dVd, dVdseffCV_dVb;
      if (T0 >= 0.0) {
        VdseffCV = VdsatCV - 0.5 * (T0 + T1);
        dVdseffCV_dVg = 0.5 * (dT0_dVg - dT1_dVg);
        dVdseffCV_dVd = 0.5 * (1.0 - dT1_dVd);
        dVdseffCV_dVb = 0.5 * (dT0_dVb - dT1_dVb);
      } else {
        T3 = (kDelta4 + kDelta4) / (T1 - T0);
        T4 = 1.0 - T3;
        T5 = VdsatCV * T3 / (T1 - T0);
        VdseffCV = VdsatCV * T4;
        dVdseffCV_dVg = dT0_dVg * T4 + T5 * (dT1_dVg - dT0_dVg);
        dVdseffCV_dVd = T5 * (dT1_dVd + 1.0);
        dVdseffCV_dVb = dT0_dVb * (T4 - T5) + T5 * dT1_dVb;
      }
      if (Vds == 0.0) { VdseffCV = 0.0; dVdseffCV_dVg = 0.0; dVdseffCV_dVb = 0.0; }

      T0 = AbulkCV * VdseffCV;
      T1 = 12.0 * (Vgsteff - 0.5 * T0 + 1.0e-20);
      T2 = VdseffCV / T1;
      T3 = T0 * T2;
      T4 = (1.0 - 12.0 * T2 * T2 * AbulkCV);
      T5 = (6.0 * T0 * (4.0 * Vgsteff - T0) / (T1 * T1) - 0.5);
      T6 = 12.0 * T2 * T2 * Vgsteff;
      qgate = CoxWL * (Vgsteff - 0.5 * VdseffCV + T3);
      real Cgg1 = CoxWL * (T4 + T5 * dVdseffCV_dVg);
      const real Cgd1 = CoxWL * T5 * dVdseffCV_dVd + Cgg1 * dVgsteff_dVd;
      const real Cgb1 = CoxWL * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cgg1 * dVgsteff_dVb;
      Cgg1 *= dVgsteff_dVg;
      T7 = 1.0 - AbulkCV;
      qbulk = CoxWL * T7 * (0.5 * VdseffCV - T3);
      T4 = -T7 * (T4 - 1.0);
      T5 = -T7 * T5;
      T6 = -(T7 * T6 + (0.5 * VdseffCV - T3));
      real Cbg1 = CoxWL * (T4 + T5 * dVdseffCV_dVg);
      const real Cbd1 = CoxWL * T5 * dVdseffCV_dVd + Cbg1 * dVgsteff_dVd;
      const real Cbb1 = CoxWL * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cbg1 * dVgsteff_dVb;
      Cbg1 *= dVgsteff_dVg;

      if (M.xpart > 0.5) {
        T1 = T1 + T1;
        qsrc = -CoxWL * (0.5 * Vgsteff + 0.25 * T0 - T0 * T0 / T1);
        T7 = (4.0 * Vgsteff - T0) / (T1 * T1);
        T4 = -(0.5 + 24.0 * T0 * T0 / (T1 * T1));
        T5 = -(0.25 * AbulkCV - 12.0 * AbulkCV * T0 * T7);
        T6 = -(0.25 * VdseffCV - 12.0 * T0 * VdseffCV * T7);
        Csg = CoxWL * (T4 + T5 * dVdseffCV_dVg);
        Csd = CoxWL * T5 * dVdseffCV_dVd + Csg * dVgsteff_dVd;
        Csb = CoxWL * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb;
        Csg *= dVgsteff_dVg;
      } else if (M.xpart < 0.5) {
        T1 = T1 / 12.0;
        T2 = 0.5 * CoxWL / (T1 * T1);
        T3 = Vgsteff * (2.0 * T0 * T0 / 3.0 + Vgsteff * (Vgsteff - 4.0 * T0 / 3.0)) - 2.0 * T0 * T0 * T0 / 15.0;
        qsrc = -T2 * T3;
        T7 = 4.0 / 3.0 * Vgsteff * (Vgsteff - T0) + 0.4 * T0 * T0;
        T4 = -2.0 * qsrc / T1 - T2 * (Vgsteff * (3.0 * Vgsteff - 8.0 * T0 / 3.0) + 2.0 * T0 * T0 / 3.0);
        T5 = (qsrc / T1 + T2 * T7) * AbulkCV;
        T6 = (qsrc / T1 * VdseffCV + T2 * T7 * VdseffCV);
        Csg = (T4 + T5 * dVdseffCV_dVg);
        Csd = T5 * dVdseffCV_dVd + Csg * dVgsteff_dVd;
        Csb = (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb;
        Csg *= dVgsteff_dVg;
      } else {
        qsrc = -0.5 * (qgate + qbulk);
        Csg = -0.5 * (Cgg1 + Cbg1);
        Csb = -0.5 * (Cgb1 + Cbb1);
        Csd = -0.5 * (Cgd1 + Cbd1);
      }
      qgate += Qac0 + Qsub0;
      qbulk -= (Qac0 + Qsub0);
      qdrn = -(qgate + qbulk + qsrc);
      Cgg = dQac0_dVg + dQsub0_dVg + Cgg1;
      Cgd = dQsub0_dVd + Cgd1;
      Cgb = dQac0_dVb + dQsub0_dVb + Cgb1;
      Cbg = Cbg1 - dQac0_dVg - dQsub0_dVg;
      Cbd = Cbd1 - dQsub0_dVd;
      Cbb = Cbb1 - dQac0_dVb - dQsub0_dVb;
      Cgb *= dVbseff_dVb;
      Cbb *= dVbseff_dVb;
      Csb *= dVbseff_dVb;
      W.cggb = Cgg;
      W.cgsb = -(Cgg + Cgd + Cgb);
      W.cgdb = Cgd;
      W.cdgb = -(Cgg + Cbg + Csg);
      W.cdsb = (Cgg + Cgd + Cgb + Cbg + Cbd + Cbb + Csg + Csd + Csb);
      W.cddb = -(Cgd + Cbd + Csd);
      W.cbgb = Cbg;
      W.cbsb = -(Cbg + Cbd + Cbb);
      W.cbdb = Cbd;
    } else if (M.capMod == 2) {
      // charge-thickness model
      const real vfbzb = I.vfbzb;
      real V3 = vfbzb - Vgs_eff + VbseffCV - kDelta3;
      if (vfbzb <= 0.0) T0 = sqrt(V3 * V3 - 4.0 * kDelta3 * vfbzb);
      else T0 = sqrt(V3 * V3 + 4.0 * kDelta3 * vfbzb);
      T1 = 0.5 * (1.0 + V3 / T0);
      const real Vfbeff = vfbzb - 0.5 * (V3 + T0);
      const real dVfbeff_dVg = T1 * dVgs_eff_dVg;
      const real dVfbeff_dVb = -T1 * dVbseffCV_dVb;

      const real Cox = I.coxp;
      real Tox = 1.0e8 * I.toxp;
      T0 = (Vgs_eff - VbseffCV - vfbzb) / Tox;
      dT0_dVg = dVgs_eff_dVg / Tox;
      dT0_dVb = -dVbseffCV_dVb / Tox;
      real Tcen, dTcen_dVg, dTcen_dVd, dTcen_dVb;
      tmp = T0 * P.acde;
      if ((-kExpThr < tmp) && (tmp < kExpThr)) {
        Tcen = P.ldeb * exp(tmp);
        dTcen_dVg = P.acde * Tcen;
        dTcen_dVb = dTcen_dVg * dT0_dVb;
        dTcen_dVg *= dT0_dVg;
      } else if (tmp <= -kExpThr) {
        Tcen = P.ldeb * kMinExp;
        dTcen_dVg = dTcen_dVb = 0.0;
      } else {
        Tcen = P.ldeb * kMaxExp;
        dTcen_dVg = dTcen_dVb = 0.0;
      }
      const real LINK = 1.0e-3 * I.toxp;
      V3 = P.ldeb - Tcen - LINK;
      const real V4 = sqrt(V3 * V3 + 4.0 * LINK * P.ldeb);
      Tcen = P.ldeb - 0.5 * (V3 + V4);
      T1 = 0.5 * (1.0 + V3 / V4);
      dTcen_dVg *= T1;
      dTcen_dVb *= T1;

      real Ccen = epssub / Tcen;
      T2 = Cox / (Cox + Ccen);
      real Coxeff = T2 * Ccen;
      T3 = -Ccen / Tcen;
      real dCoxeff_dVg = T2 * T2 * T3;
      real dCoxeff_dVb = dCoxeff_dVg * dTcen_dVb;
      real dCoxeff_dVd;
      dCoxeff_dVg *= dTcen_dVg;
      real CoxWLcen = CoxWL * Coxeff / M.coxe;

      const real Qac0 = CoxWLcen * (Vfbeff - vfbzb);
      real QovCox = Qac0 / Coxeff;
      const real dQac0_dVg = CoxWLcen * dVfbeff_dVg + QovCox * dCoxeff_dVg;
      const real dQac0_dVb = CoxWLcen * dVfbeff_dVb + QovCox * dCoxeff_dVb;

      T0 = 0.5 * P.k1ox;
      T3 = Vgs_eff - Vfbeff - VbseffCV - Vgsteff;
      if (P.k1ox == 0.0) { T1 = 0.0; T2 = 0.0; }
      else if (T3 < 0.0) { T1 = T0 + T3 / P.k1ox; T2 = CoxWLcen; }
      else { T1 = sqrt(T0 * T0 + T3); T2 = CoxWLcen * T0 / T1; }
      const real Qsub0 = CoxWLcen * P.k1ox * (T1 - T0);
      QovCox = Qsub0 / Coxeff;
      const real dQsub0_dVg = T2 * (dVgs_eff_dVg - dVfbeff_dVg - dVgsteff_dVg) + QovCox * dCoxeff_dVg;
      const real dQsub0_dVd = -T2 * dVgsteff_dVd;
      const real dQsub0_dVb = -T2 * (dVfbeff_dVb + dVbseffCV_dVb + dVgsteff_dVb) + QovCox * dCoxeff_dVb;

      XB_SYNC_POINT_U(2);
      // gate-bias dependent delta Phis (inversion charge centroid)
      real Denomi;
      if (P.k1ox <= 0.0) { Denomi = 0.25 * P.moin * Vtm; T0 = 0.5 * P.sqrtPhi; }
      else { Denomi = P.moin * Vtm * P.k1ox * P.k1ox; T0 = P.k1ox * P.sqrtPhi; }
      T1 = 2.0 * T0 + Vgsteff;
      const real DeltaPhi = Vtm * log(1.0 + T1 * Vgsteff / Denomi);
      const real dDeltaPhi_dVg = 2.0 * Vtm * (T1 - T0) / (Denomi + T1 * Vgsteff);

      T0 = Vgsteff - DeltaPhi - 0.001;
      dT0_dVg = 1.0 - dDeltaPhi_dVg;
      T1 = sqrt(T0 * T0 + Vgsteff * 0.004);
      const real VgDP = 0.5 * (T0 + T1);
      const real dVgDP_dVg = 0.5 * (dT0_dVg + (T0 * dT0_dVg + 0.002) / T1);

      Tox += Tox;
      T0 = (Vgsteff + I.vtfbphi2) / Tox;
      tmp = exp(M.bdos * 0.7 * log(T0));
      T1 = 1.0 + tmp;
      T2 = M.bdos * 0.7 * tmp / (T0 * Tox);
      Tcen = M.ados * 1.9e-9 / T1;
      dTcen_dVg = -Tcen * T2 / T1;
      dTcen_dVd = dTcen_dVg * dVgsteff_dVd;
      dTcen_dVb = dTcen_dVg * dVgsteff_dVb;
      dTcen_dVg *= dVgsteff_dVg;

      Ccen = epssub / Tcen;
      T0 = Cox / (Cox + Ccen);
      Coxeff = T0 * Ccen;
      T1 = -Ccen / Tcen;
      dCoxeff_dVg = T0 * T0 * T1;
      dCoxeff_dVd = dCoxeff_dVg * dTcen_dVd;
      dCoxeff_dVb = dCoxeff_dVg * dTcen_dVb;
      dCoxeff_dVg *= dTcen_dVg;
      CoxWLcen = CoxWL * Coxeff / M.coxe;
      W.Coxeff = Coxeff;

      const real AbulkCV = Abulk0 * P.abulkCVfactor;
      const real dAbulkCV_dVb = P.abulkCVfactor * dAbulk0_dVb;
      const real VdsatCV = VgDP / AbulkCV;
      T0 = VdsatCV - Vds - kDelta4;
      dT0_dVg = dVgDP_dVg / AbulkCV;
      dT0_dVb = -VdsatCV * dAbulkCV_dVb / AbulkCV;
      T1 = sqrt(T0 * T0 + 4.0 * kDelta4 * VdsatCV);
      dT1_dVg = (T0 + kDelta4 + kDelta4) / T1;
      dT1_dVd = -T0 / T1;
      dT1_dVb = dT1_dVg * dT0_dVb;
      dT1_dVg *= dT0_dVg;
      real VdseffCV, dVdseffCV_dVg, dVdseffCV_dVd, dVdseffCV_dVb;
      if (T0 >= 0.0) {
        VdseffCV = VdsatCV - 0.5 * (T0 + T1);
        dVdseffCV_dVg = 0.5 * (dT0_dVg - dT1_dVg);
        dVdseffCV_dVd = 0.5 * (1.0 - dT1_dVd);
        dVdseffCV_dVb = 0.5 * (dT0_dVb - dT1_dVb);
      } else {
        T3 = (kDelta4 + kDelta4) / (T1 - T0);
        T4 = 1.0 - T3;
        T5 = VdsatCV * T3 / (T1 - T0);
        VdseffCV = VdsatCV * T4;
        dVdseffCV_dVg = dT0_dVg * T4 + T5 * (dT1_dVg - dT0_dVg);
        dVdseffCV_dVd = T5 * (dT1_dVd + 1.0);
        dVdseffCV_dVb = dT0_dVb * (T4 - T5) + T5 * dT1_dVb;
      }
      if (Vds == 0.0) { VdseffCV = 0.0; dVdseffCV_dVg = 0.0; dVdseffCV_dVb = 0.0; }

      T0 = AbulkCV * VdseffCV;
      T1 = VgDP;
      T2 = 12.0 * (T1 - 0.5 * T0 + 1.0e-20);
      T3 = T0 / T2;
      T4 = 1.0 - 12.0 * T3 * T3;
      T5 = AbulkCV * (6.0 * T0 * (4.0 * T1 - T0) / (T2 * T2) - 0.5);
      T6 = T5 * VdseffCV / AbulkCV;
      qgate = CoxWLcen * (T1 - T0 * (0.5 - T3));
      QovCox = qgate / Coxeff;
      real Cgg1 = CoxWLcen * (T4 * dVgDP_dVg + T5 * dVdseffCV_dVg);
      const real Cgd1 = CoxWLcen * T5 * dVdseffCV_dVd + Cgg1 * dVgsteff_dVd + QovCox * dCoxeff_dVd;
      const real Cgb1 = CoxWLcen * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cgg1 * dVgsteff_dVb + QovCox * dCoxeff_dVb;
      Cgg1 = Cgg1 * dVgsteff_dVg + QovCox * dCoxeff_dVg;

      T7 = 1.0 - AbulkCV;
      T8 = T2 * T2;
      T9 = 12.0 * T7 * T0 * T0 / (T8 * AbulkCV);
      T10 = T9 * dVgDP_dVg;
      T11 = -T7 * T5 / AbulkCV;
      T12 = -(T9 * T1 / AbulkCV + VdseffCV * (0.5 - T0 / T2));
      qbulk = CoxWLcen * T7 * (0.5 * VdseffCV - T0 * VdseffCV / T2);
      QovCox = qbulk / Coxeff;
      real Cbg1 = CoxWLcen * (T10 + T11 * dVdseffCV_dVg);
      const real Cbd1 = CoxWLcen * T11 * dVdseffCV_dVd + Cbg1 * dVgsteff_dVd + QovCox * dCoxeff_dVd;
      const real Cbb1 = CoxWLcen * (T11 * dVdseffCV_dVb + T12 * dAbulkCV_dVb) + Cbg1 * dVgsteff_dVb + QovCox * dCoxeff_dVb;
      Cbg1 = Cbg1 * dVgsteff_dVg + QovCox * dCoxeff_dVg;

      if (M.xpart > 0.5) {
        qsrc = -CoxWLcen * (T1 / 2.0 + T0 / 4.0 - 0.5 * T0 * T0 / T2);
        QovCox = qsrc / Coxeff;
        T2 += T2;
        T3 = T2 * T2;
        T7 = -(0.25 - 12.0 * T0 * (4.0 * T1 - T0) / T3);
        T4 = -(0.5 + 24.0 * T0 * T0 / T3) * dVgDP_dVg;
        T5 = T7 * AbulkCV;
        T6 = T7 * VdseffCV;
        Csg = CoxWLcen * (T4 + T5 * dVdseffCV_dVg);
        Csd = CoxWLcen * T5 * dVdseffCV_dVd + Csg * dVgsteff_dVd + QovCox * dCoxeff_dVd;
        Csb = CoxWLcen * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb + QovCox * dCoxeff_dVb;
        Csg = Csg * dVgsteff_dVg + QovCox * dCoxeff_dVg;
      } else if (M.xpart < 0.5) {
        T2 = T2 / 12.0;
        T3 = 0.5 * CoxWLcen / (T2 * T2);
        T4 = T1 * (2.0 * T0 * T0 / 3.0 + T1 * (T1 - 4.0 * T0 / 3.0)) - 2.0 * T0 * T0 * T0 / 15.0;
        qsrc = -T3 * T4;
        QovCox = qsrc / Coxeff;
        T8 = 4.0 / 3.0 * T1 * (T1 - T0) + 0.4 * T0 * T0;
        T5 = -2.0 * qsrc / T2 - T3 * (T1 * (3.0 * T1 - 8.0 * T0 / 3.0) + 2.0 * T0 * T0 / 3.0);
        T6 = AbulkCV * (qsrc / T2 + T3 * T8);
        T7 = T6 * VdseffCV / AbulkCV;
        Csg = T5 * dVgDP_dVg + T6 * dVdseffCV_dVg;
        Csd = Csg * dVgsteff_dVd + T6 * dVdseffCV_dVd + QovCox * dCoxeff_dVd;
        Csb = Csg * dVgsteff_dVb + T6 * dVdseffCV_dVb + T7 * dAbulkCV_dVb + QovCox * dCoxeff_dVb;
        Csg = Csg * dVgsteff_dVg + QovCox * dCoxeff_dVg;
      } else {
        qsrc = -0.5 * qgate;
        Csg = -0.5 * Cgg1;
        Csd = -0.5 * Cgd1;
        Csb = -0.5 * Cgb1;
      }
      qgate += Qac0 + Qsub0 - qbulk;
      qbulk -= (Qac0 + Qsub0);
      qdrn = -(qgate + qbulk + qsrc);
      Cbg = Cbg1 - dQac0_dVg - dQsub0_dVg;
      Cbd = Cbd1 - dQsub0_dVd;
      Cbb = Cbb1 - dQac0_dVb - dQsub0_dVb;
      Cgg = Cgg1 - Cbg;
      Cgd = Cgd1 - Cbd;
      Cgb = Cgb1 - Cbb;
      Cgb *= dVbseff_dVb;
      Cbb *= dVbseff_dVb;
      Csb *= dVbseff_dVb;
      W.cggb = Cgg;
      W.cgsb = -(Cgg + Cgd + Cgb);
      W.cgdb = Cgd;
      W.cdgb = -(Cgg + Cbg + Csg);
      W.cdsb = (Cgg + Cgd + Cgb + Cbg + Cbd + Cbb + Csg + Csd + Csb);
      W.cddb = -(Cgd + Cbd + Csd);
      W.cbgb = Cbg;
      W.cbsb = -(Cbg + Cbd + Cbb);
      W.cbdb = Cbd;
    }
  }
  W.Vth = Vth;
  W.Vdsat = Vdsat;
  W.qgate = qgate; W.qbulk = qbulk; W.qdrn = qdrn; W.qsrc = qsrc;

  W.csgb = -W.cggb - W.cdgb - W.cbgb;
  W.csdb = -W.cgdb - W.cddb - W.cbdb;
  W.cssb = -W.cgsb - W.cdsb - W.cbsb;
  W.cgbb = -W.cgdb - W.cggb - W.cgsb;
  W.cdbb = -W.cddb - W.cdgb - W.cdsb;
  W.cbbb = -W.cbgb - W.cbdb - W.cbsb;
  W.csbb = -W.cgbb - W.cdbb - W.cbbb;

  XB_SYNC_POINT(1);
  // NQS: relaxation-time conductance
  if (I.trnqsMod || I.acnqsMod) {
    W.qchqs = W.qcheq = -(qbulk + qgate);
    W.cqgb = -(W.cggb + W.cbgb);
    W.cqdb = -(W.cgdb + W.cbdb);
    W.cqsb = -(W.cgsb + W.cbsb);
    W.cqbb = -(W.cqgb + W.cqdb + W.cqsb);
    CoxWL = M.coxe * P.weffCV * I.nf * P.leffCV;
    T1 = W.gcrg / CoxWL;
    W.gtau = T1 * 1.0e-9;   // ScalingFactor
    if (I.acnqsMod) W.taunet = 1.0 / T1;
  }

  // ---- S/D bulk junction depletion charge and capacitance ------------------------------------------
  if (charge_needed) {
    const real czbd = M.DunitAreaTempJctCap * I.Adeff;
    const real czbs = M.SunitAreaTempJctCap * I.Aseff;
    const real czbdsw = M.DunitLengthSidewallTempJctCap * I.Pdeff;
    const real czbdswg = M.DunitLengthGateSidewallTempJctCap * P.weffCJ * I.nf;
    const real czbssw = M.SunitLengthSidewallTempJctCap * I.Pseff;
    const real czbsswg = M.SunitLengthGateSidewallTempJctCap * P.weffCJ * I.nf;
    Real2 qc = junction_charge_v(W.vbs_jct, czbs, czbssw, czbsswg,
                    M.SbulkJctBotGradingCoeff, M.SbulkJctSideGradingCoeff, M.SbulkJctGateSideGradingCoeff,
                    M.PhiBS, M.PhiBSWS, M.PhiBSWGS);
    W.qbs = qc.a; W.capbs = qc.b;
    qc = junction_charge_v(W.vbd_jct, czbd, czbdsw, czbdswg,
                    M.DbulkJctBotGradingCoeff, M.DbulkJctSideGradingCoeff, M.DbulkJctGateSideGradingCoeff,
                    M.PhiBD, M.PhiBSWD, M.PhiBSWGD);
    W.qbd = qc.a; W.capbd = qc.b;
  } else {
    W.qbs = W.qbd = W.capbs = W.capbd = 0.0;
  }

  XB_SYNC_POINT(2);
  // ---- gate electrode resistance currents & overlap capacitances -----------------------------------------
  real vgdx, vgsx;
  if (I.rgateMod == 3) { vgdx = W.vgmd; vgsx = W.vgms; }
  else { vgdx = W.vgd; vgsx = W.vgs; }
  W.Igate = W.IgateMid = 0.0;
  if (I.rgateMod == 1) W.Igate = I.grgeltd * W.Vgegp;
  else if (I.rgateMod == 2) W.Igate = W.gcrg * W.Vgegp;
  else if (I.rgateMod == 3) { W.Igate = I.grgeltd * W.Vgegm; W.IgateMid = W.gcrg * W.Vgmgp; }

  if (M.capMod == 0) {
    W.cgdo = P.cgdo;
    W.qgdo = P.cgdo * vgdx;
    W.cgso = P.cgso;
    W.qgso = P.cgso * vgsx;
  } else {
    T0 = vgdx + kDelta1;
    T1 = sqrt(T0 * T0 + 4.0 * kDelta1);
    T2 = 0.5 * (T0 - T1);
    T3 = P.weffCV * P.cgdl;
    T4 = sqrt(1.0 - 4.0 * T2 / P.ckappad);
    W.cgdo = P.cgdo + T3 - T3 * (1.0 - 1.0 / T4) * (0.5 - 0.5 * T0 / T1);
    W.qgdo = (P.cgdo + T3) * vgdx - T3 * (T2 + 0.5 * P.ckappad * (T4 - 1.0));
    T0 = vgsx + kDelta1;
    T1 = sqrt(T0 * T0 + 4.0 * kDelta1);
    T2 = 0.5 * (T0 - T1);
    T3 = P.weffCV * P.cgsl;
    T4 = sqrt(1.0 - 4.0 * T2 / P.ckappas);
    W.cgso = P.cgso + T3 - T3 * (1.0 - 1.0 / T4) * (0.5 - 0.5 * T0 / T1);
    W.qgso = (P.cgso + T3) * vgsx - T3 * (T2 + 0.5 * P.ckappas * (T4 - 1.0));
  }
  if (I.nf != 1.0) { W.cgdo *= I.nf; W.cgso *= I.nf; W.qgdo *= I.nf; W.qgso *= I.nf; }
  W.vgdx = vgdx; W.vgsx = vgsx;
  (void)T6; (void)T7; (void)T8; (void)T9; (void)T10; (void)T11; (void)T12; (void)tmp1; (void)toxe;
}

}  // namespace b4
}  // namespace xb
