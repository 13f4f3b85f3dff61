// xyce_b200 -- BSIM4 v4.8.2 DC core: junction diodes, Vth, mobility, Vdsat,
// Ids, output conductance, substrate current, GIDL/GISL, gate tunnelling and
// bias-dependent S/D resistance.  See bsim4_eval.h for the overall contract.
// Behavioural specification: N_DEV_MOSFET_B4p82.C:3296-5842.
#pragma once

namespace xb {
namespace b4 {


// Gate-induced drain/source leakage, gidlMod = 0 (B4p82.C:5072-5155).  `T1` is the
// normalised field term, `vbx` the body-to-(drain|source) voltage.
XB_HD void gidl_mod0(real T0, real T1, real dveff_dvg, real a, real b, real c,
                     real weffCJ, real vbx, real &Ig, real &Gd, real &Gg, real &Gb) {
  if ((a <= 0.0) || (b <= 0.0) || (T1 <= 0.0) || (c <= 0.0) || (vbx > 0.0)) {
    Ig = Gd = Gg = Gb = 0.0;
    return;
  }
  const real dT1_dVd = 1.0 / T0;
  const real dT1_dVg = -dveff_dvg * dT1_dVd;
  const real T2 = b / T1;
  if (T2 < 100.0) {
    Ig = a * weffCJ * T1 * exp(-T2);
    const real T3 = Ig * (1.0 + T2) / T1;
    Gd = T3 * dT1_dVd;
    Gg = T3 * dT1_dVg;
  } else {
    Ig = a * weffCJ * 3.720075976e-44;
    Gd = Ig * dT1_dVd;
    Gg = Ig * dT1_dVg;
    Ig *= T1;
  }
  const real T4 = vbx * vbx;
  const real T5 = -vbx * T4;
  const real T6 = c + T5;
  const real T7 = T5 / T6;
  const real T8 = 3.0 * c * T4 / T6 / T6;
  Gd = Gd * T7 + Ig * T8;
  Gg = Gg * T7;
  Gb = -Ig * T8;
  Ig *= T7;
}

// gidlMod = 1 (B4p82.C:5157-5290).
XB_HD void gidl_mod1(real T0, real T1, real rg, real dveff_dvg, real a, real b, real c,
                     real k, real f, real weffCJ, real vbx, real clamp,
                     real &Ig, real &Gd, real &Gg, real &Gb) {
  if ((a <= 0.0) || (b <= 0.0) || (T1 <= 0.0) || (c < 0.0)) {
    Ig = Gd = Gg = Gb = 0.0;
    return;
  }
  const real dT1_dVd = 1 / T0;
  const real dT1_dVg = -rg * dT1_dVd * dveff_dvg;
  const real T2 = b / T1;
  real T3;
  if (T2 < kExpLThr) {
    Ig = weffCJ * a * T1 * exp(-T2);
    T3 = Ig / T1 * (T2 + 1);
    Gd = T3 * dT1_dVd;
    Gg = T3 * dT1_dVg;
  } else {
    T3 = weffCJ * a * kMinExpL;
    Ig = T3 * T1;
    Gd = T3 * dT1_dVd;
    Gg = T3 * dT1_dVg;
  }
  real T4 = vbx - f;
  if (T4 > clamp) T4 = clamp;
  const real T5 = (T4 == 0) ? kExpLThr : k / T4;
  real T6;
  if (T5 < kExpLThr) {
    T6 = exp(T5);
    Gb = -Ig * T6 * T5 / T4;
  } else {
    T6 = kMaxExpL;
    Gb = 0.0;
  }
  Gd *= T6;
  Gg *= T6;
  Ig *= T6;
}

// Everything the C-V stage needs from the DC stage besides B4Mid.
XB_HELPER Real4 gidl_mod0_v(real T0, real T1, real dveff_dvg, real a, real b, real c, real weffCJ, real vbx) {
  Real4 r; gidl_mod0(T0, T1, dveff_dvg, a, b, c, weffCJ, vbx, r.a, r.b, r.c, r.d); return r;
}
XB_HELPER Real4 gidl_mod1_v(real T0, real T1, real rg, real dveff_dvg, real a, real b, real c, real k, real f, real weffCJ,
                            real vbx, real clamp) {
  Real4 r; gidl_mod1(T0, T1, rg, dveff_dvg, a, b, c, k, f, weffCJ, vbx, clamp, r.a, r.b, r.c, r.d); return r;
}

struct DcCarry {
  real Vds, Vgs, Vbs, Vdb;
  real Vbseff, dVbseff_dVb, Phis, dPhis_dVb, sqrtPhis, dsqrtPhis_dVb;
  real Vth, dVth_dVb, dVth_dVd;
  real Vgs_eff, dVgs_eff_dVg, Vgst;
  real n, dn_dVb, dn_dVd, Vtm, Vtm0;
  real Vgsteff, dVgsteff_dVg, dVgsteff_dVd, dVgsteff_dVb;
  real Vdseff, dVdseff_dVg, dVdseff_dVd, dVdseff_dVb;
  real Abulk, dAbulk_dVb, dAbulk_dVg, Abulk0, dAbulk0_dVb;
  real Weff, Leff, epsrox, toxe, epssub;
  real Vfb;      // flat-band of the gate-current section (0 when igcMod == igbMod == 0)
  real dCoxeff_dVg;
  real Vdsat;    // as left by the tnoiMod block
};

XB_HD void stage_dc(const SolverFlags &S, const B4Model &M, const B4Size &P,
                    const B4Inst &I, B4Mid &W, DcCarry &C) {
  real T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, T13, T14;
  real dT0_dVg, dT0_dVd, dT0_dVb, dT1_dVg, dT1_dVd, dT1_dVb;
  real dT2_dVg, dT2_dVd, dT2_dVb, dT3_dVg, dT3_dVd, dT3_dVb;
  real dT4_dVd, dT4_dVb, dT5_dVg, dT5_dVd, dT5_dVb;
  real dT6_dVg, dT6_dVd, dT6_dVb, dT7_dVg, dT7_dVd, dT7_dVb;
  real dT8_dVg, dT8_dVd, dT8_dVb, dT9_dVg, dT9_dVd, dT9_dVb;
  real dT10_dVg, dT10_dVd, dT10_dVb;
  const real gmin = S.gmin;
  const bool kV48 = M.versionDouble >= 4.8;
  const bool kV47 = M.versionDouble >= 4.7;      // 4.6.1 evaluator differences (N_DEV_MOSFET_B4p61.C vs B4p70.C)

  XB_SYNC_POINT(2);
  // ---- source / drain bulk junction diodes --------------------------------
  {
    JctPar js;
    js.Nvtm = M.vtm * M.SjctEmissionCoeff;
    if ((I.Aseff <= 0.0) && (I.Pseff <= 0.0)) js.Isat = kV47 ? 0.0 : 1.0e-14;      // B4p61.C: 1.0e-14
    else js.Isat = I.Aseff * M.SjctTempSatCurDensity + I.Pseff * M.SjctSidewallTempSatCurDensity
                 + P.weffCJ * I.nf * M.SjctGateSidewallTempSatCurDensity;
    js.xjbv = M.xjbvs; js.bv = M.bvs; js.XExpBV = I.XExpBVS;
    js.vjmFwd = I.vjsmFwd; js.vjmRev = I.vjsmRev; js.IVjmFwd = I.IVjsmFwd; js.IVjmRev = I.IVjsmRev;
    js.slpFwd = I.SslpFwd; js.slpRev = I.SslpRev;
    { const Real2 r = junction_diode_v(M.dioMod, js, W.vbs_jct, gmin); W.gbs = r.a; W.cbs = r.b; }

    JctPar jd;
    jd.Nvtm = M.vtm * M.DjctEmissionCoeff;
    if ((I.Adeff <= 0.0) && (I.Pdeff <= 0.0)) jd.Isat = kV47 ? 0.0 : 1.0e-14;
    else jd.Isat = I.Adeff * M.DjctTempSatCurDensity + I.Pdeff * M.DjctSidewallTempSatCurDensity
                 + P.weffCJ * I.nf * M.DjctGateSidewallTempSatCurDensity;
    jd.xjbv = M.xjbvd; jd.bv = M.bvd; jd.XExpBV = I.XExpBVD;
    jd.vjmFwd = I.vjdmFwd; jd.vjmRev = I.vjdmRev; jd.IVjmFwd = I.IVjdmFwd; jd.IVjmRev = I.IVjdmRev;
    jd.slpFwd = I.DslpFwd; jd.slpRev = I.DslpRev;
    { const Real2 r = junction_diode_v(M.dioMod, jd, W.vbd_jct, gmin); W.gbd = r.a; W.cbd = r.b; }

    // trap-assisted tunnelling / recombination in reverse bias
    real t1, d1, t2, d2, t3, d3, t4, d4, t5, d5, t6, d6;
    { const Real2 r = tat_term_v(M.vtss, M.vtm0 * M.njtsstemp, W.vbs_jct); t1 = r.a; d1 = r.b; }
    { const Real2 r = tat_term_v(M.vtsd, M.vtm0 * M.njtsdtemp, W.vbd_jct); t2 = r.a; d2 = r.b; }
    { const Real2 r = tat_term_v(M.vtssws, M.vtm0 * M.njtsswstemp, W.vbs_jct); t3 = r.a; d3 = r.b; }
    { const Real2 r = tat_term_v(M.vtsswd, M.vtm0 * M.njtsswdtemp, W.vbd_jct); t4 = r.a; d4 = r.b; }
    { const Real2 r = tat_term_v(M.vtsswgs, M.vtm0 * M.njtsswgstemp, W.vbs_jct); t5 = r.a; d5 = r.b; }
    { const Real2 r = tat_term_v(M.vtsswgd, M.vtm0 * M.njtsswgdtemp, W.vbd_jct); t6 = r.a; d6 = r.b; }
    W.gbs += I.SjctTempRevSatCur * d1 + I.SswTempRevSatCur * d3 + I.SswgTempRevSatCur * d5;
    W.cbs -= I.SjctTempRevSatCur * (t1 - 1.0) + I.SswTempRevSatCur * (t3 - 1.0)
           + I.SswgTempRevSatCur * (t5 - 1.0);
    W.gbd += I.DjctTempRevSatCur * d2 + I.DswTempRevSatCur * d4 + I.DswgTempRevSatCur * d6;
    W.cbd -= I.DjctTempRevSatCur * (t2 - 1.0) + I.DswTempRevSatCur * (t4 - 1.0)
           + I.DswgTempRevSatCur * (t6 - 1.0);
  }

  XB_SYNC_POINT(2);
  // ---- mode selection -----------------------------------------------------
  real Vds, Vgs, Vbs, Vdb;
  if (W.vds >= 0.0) { W.mode = 1;  Vds = W.vds;  Vgs = W.vgs; Vbs = W.vbs; Vdb = W.vds - W.vbs; }
  else              { W.mode = -1; Vds = -W.vds; Vgs = W.vgd; Vbs = W.vbd; Vdb = -W.vbs; }

  real epsrox, toxe, epssub;
  if (M.mtrlMod) { epsrox = 3.9; toxe = M.eot; epssub = kEps0B4 * M.epsrsub; }
  else { epsrox = M.epsrox; toxe = M.toxe; epssub = kEpsSi; }

  if (S.artParameterFlag) {   // DCOP homotopy (DeviceSupport::contVds / contVgst)
    real mn = S.vdsScaleMin; if (mn <= 0.0) mn = 0.3;
    Vds = Vds * (S.nltermScale * (1.0 - mn) + mn);
    Vgs = S.gainScale * Vgs + (1.0 - S.gainScale) * S.vgstConst;
  }

  XB_SYNC_POINT(2);
  // ---- effective body bias ------------------------------------------------
  real Vbseff, dVbseff_dVb;
  T0 = Vbs - I.vbsc - 0.001;
  T1 = sqrt(T0 * T0 - 0.004 * I.vbsc);
  if (T0 >= 0.0) {
    Vbseff = I.vbsc + 0.5 * (T0 + T1);
    dVbseff_dVb = 0.5 * (1.0 + T0 / T1);
  } else {
    T2 = -0.002 / (T1 - T0);
    Vbseff = I.vbsc * (1.0 + T2);
    dVbseff_dVb = T2 * I.vbsc / T1;
  }
  T9 = 0.95 * P.phi;
  T0 = T9 - Vbseff - 0.001;
  T1 = sqrt(T0 * T0 + 0.004 * T9);
  Vbseff = T9 - 0.5 * (T0 + T1);
  dVbseff_dVb *= 0.5 * (1.0 + T0 / T1);

  const real Phis = P.phi - Vbseff;
  const real dPhis_dVb = -1.0;
  const real sqrtPhis = sqrt(Phis);
  const real dsqrtPhis_dVb = -0.5 / sqrtPhis;
  const real Xdep = P.Xdep0 * sqrtPhis / P.sqrtPhi;
  const real dXdep_dVb = (P.Xdep0 / P.sqrtPhi) * dsqrtPhis_dVb;
  const real Leff = P.leff;
  const real Vtm = M.vtm;
  const real Vtm0 = M.vtm0;

  XB_SYNC_POINT(1);
  // ---- threshold voltage --------------------------------------------------
  T3 = sqrt(Xdep);
  const real V0 = P.vbi - P.phi;
  T0 = P.dvt2 * Vbseff;
  if (T0 >= -0.5) { T1 = 1.0 + T0; T2 = P.dvt2; }
  else { T4 = 1.0 / (3.0 + 8.0 * T0); T1 = (1.0 + 3.0 * T0) * T4; T2 = P.dvt2 * T4 * T4; }
  const real lt1 = M.factor1 * T3 * T1;
  const real dlt1_dVb = M.factor1 * (0.5 / T3 * T1 * dXdep_dVb + T3 * T2);

  T0 = P.dvt2w * Vbseff;
  if (T0 >= -0.5) { T1 = 1.0 + T0; T2 = P.dvt2w; }
  else { T4 = 1.0 / (3.0 + 8.0 * T0); T1 = (1.0 + 3.0 * T0) * T4; T2 = P.dvt2w * T4 * T4; }
  const real ltw = M.factor1 * T3 * T1;
  const real dltw_dVb = M.factor1 * (0.5 / T3 * T1 * dXdep_dVb + T3 * T2);

  real Theta0, dTheta0_dVb;
  T0 = P.dvt1 * Leff / lt1;
  if (T0 < kExpThr) {
    T1 = exp(T0);
    T2 = T1 - 1.0;
    T3 = T2 * T2;
    T4 = T3 + 2.0 * T1 * kMinExp;
    Theta0 = T1 / T4;
    dT1_dVb = -T0 * T1 * dlt1_dVb / lt1;
    dTheta0_dVb = dT1_dVb * (T4 - 2.0 * T1 * (T2 + kMinExp)) / T4 / T4;
  } else {
    Theta0 = 1.0 / (kMaxExp - 2.0);
    dTheta0_dVb = 0.0;
  }
  W.thetavth = P.dvt0 * Theta0;
  const real Delt_vth = W.thetavth * V0;
  const real dDelt_vth_dVb = P.dvt0 * dTheta0_dVb * V0;

  T0 = P.dvt1w * P.weff * Leff / ltw;
  if (T0 < kExpThr) {
    T1 = exp(T0);
    T2 = T1 - 1.0;
    T3 = T2 * T2;
    T4 = T3 + 2.0 * T1 * kMinExp;
    T5 = T1 / T4;
    dT1_dVb = -T0 * T1 * dltw_dVb / ltw;
    dT5_dVb = dT1_dVb * (T4 - 2.0 * T1 * (T2 + kMinExp)) / T4 / T4;
  } else {
    T5 = 1.0 / (kMaxExp - 2.0);
    dT5_dVb = 0.0;
  }
  T0 = P.dvt0w * T5;
  T2 = T0 * V0;
  dT2_dVb = P.dvt0w * dT5_dVb * V0;

  const real TempRatio = I.temp / M.tnom - 1.0;
  T0 = sqrt(1.0 + P.lpe0 / Leff);
  T1 = P.k1ox * (T0 - 1.0) * P.sqrtPhi + (P.kt1 + P.kt1l / Leff + P.kt2 * Vbseff) * TempRatio;
  const real Vth_NarrowW = toxe * P.phi / (P.weff + P.w0);

  T3 = I.eta0 + P.etab * Vbseff;
  if (T3 < 1.0e-4) {
    T9 = 1.0 / (3.0 - 2.0e4 * T3);
    T3 = (2.0e-4 - T3) * T9;
    T4 = T9 * T9;
  } else {
    T4 = 1.0;
  }
  const real dDIBL_Sft_dVd = T3 * P.theta0vb0;
  const real DIBL_Sft = dDIBL_Sft_dVd * Vds;
  const real Lpe_Vb = sqrt(1.0 + P.lpeb / Leff);

  real Vth = M.dtype * I.vth0 + (P.k1ox * sqrtPhis - P.k1 * P.sqrtPhi) * Lpe_Vb
             - I.k2ox * Vbseff - Delt_vth - T2 + (P.k3 + P.k3b * Vbseff) * Vth_NarrowW + T1 - DIBL_Sft;
  real dVth_dVb = Lpe_Vb * P.k1ox * dsqrtPhis_dVb - I.k2ox - dDelt_vth_dVb - dT2_dVb
                  + P.k3b * Vth_NarrowW - P.etab * Vds * P.theta0vb0 * T4 + P.kt2 * TempRatio;
  real dVth_dVd = -dDIBL_Sft_dVd;

  // subthreshold swing factor n
  real n, dn_dVb, dn_dVd;
  {
    const real tmp1 = epssub / Xdep;
    // (nstar, a noise-only quantity at B4p82.C:3908, is not evaluated)
    const real tmp2 = P.nfactor * tmp1;
    const real tmp3 = P.cdsc + P.cdscb * Vbseff + P.cdscd * Vds;
    const real tmp4 = (tmp2 + tmp3 * Theta0 + P.cit) / M.coxe;
    if (tmp4 >= -0.5) {
      n = 1.0 + tmp4;
      dn_dVb = (-tmp2 / Xdep * dXdep_dVb + tmp3 * dTheta0_dVb + P.cdscb * Theta0) / M.coxe;
      dn_dVd = P.cdscd * Theta0 / M.coxe;
    } else {
      T0 = 1.0 / (3.0 + 8.0 * tmp4);
      n = (1.0 + 3.0 * tmp4) * T0;
      T0 *= T0;
      dn_dVb = (-tmp2 / Xdep * dXdep_dVb + tmp3 * dTheta0_dVb + P.cdscb * Theta0) / M.coxe * T0;
      dn_dVd = P.cdscd * Theta0 / M.coxe * T0;
    }
  }

  // Vth corrections for pocket devices (DITS)
  if (P.dvtp0 > 0.0) {
    T0 = -P.dvtp1 * Vds;
    if (T0 < -kExpThr) { T2 = kMinExp; dT2_dVd = 0.0; }
    else { T2 = exp(T0); dT2_dVd = -P.dvtp1 * T2; }
    T3 = Leff + P.dvtp0 * (1.0 + T2);
    dT3_dVd = P.dvtp0 * dT2_dVd;
    if (M.tempMod < 2) { T4 = Vtm * log(Leff / T3); dT4_dVd = -Vtm * dT3_dVd / T3; }
    else { T4 = M.vtm0 * log(Leff / T3); dT4_dVd = -M.vtm0 * dT3_dVd / T3; }
    const real dDITS_Sft_dVd = dn_dVd * T4 + n * dT4_dVd;
    const real dDITS_Sft_dVb = T4 * dn_dVb;
    Vth -= n * T4;
    dVth_dVd -= dDITS_Sft_dVd;
    dVth_dVb -= dDITS_Sft_dVb;
  }
  if (kV47 && !((P.dvtp4 == 0.0) || (P.dvtp2factor == 0.0))) {      // DITS_SFT2: 4.7 and later
    T1 = 2.0 * P.dvtp4 * Vds;
    dexp(T1, T0, T10);
    const real DITS_Sft2 = P.dvtp2factor * (T0 - 1) / (T0 + 1);
    const real dDITS_Sft2_dVd = P.dvtp2factor * P.dvtp4 * 4.0 * T10 / ((T0 + 1) * (T0 + 1));
    Vth -= DITS_Sft2;
    dVth_dVd -= dDITS_Sft2_dVd;
  }
  W.Vth = Vth;
  W.von = Vth;

  XB_SYNC_POINT(2);
  // ---- poly gate depletion -------------------------------------------------
  T0 = I.vfb + P.phi;
  T1 = (M.mtrlMod == 0) ? kEpsSi : M.epsrgate * kEps0B4;
  { const Real2 r = poly_depletion_v(T0, P.ngate, T1, M.coxe, W.vgs); W.vgs_eff = r.a; W.dvgs_eff_dvg = r.b; }
  { const Real2 r = poly_depletion_v(T0, P.ngate, T1, M.coxe, W.vgd); W.vgd_eff = r.a; W.dvgd_eff_dvg = r.b; }
  real Vgs_eff, dVgs_eff_dVg;
  if (W.mode > 0) { Vgs_eff = W.vgs_eff; dVgs_eff_dVg = W.dvgs_eff_dvg; }
  else { Vgs_eff = W.vgd_eff; dVgs_eff_dVg = W.dvgd_eff_dvg; }
  const real Vgst = Vgs_eff - Vth;

  XB_SYNC_POINT(2);
  // ---- effective Vgst -------------------------------------------------------
  T0 = n * Vtm;
  T1 = P.mstar * Vgst;
  T2 = T1 / T0;
  if (T2 > kExpThr) {
    T10 = T1;
    dT10_dVg = P.mstar * dVgs_eff_dVg;
    dT10_dVd = -dVth_dVd * P.mstar;
    dT10_dVb = -dVth_dVb * P.mstar;
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
    dT10_dVg = P.mstar * ExpVgst / (1.0 + ExpVgst);
    dT10_dVb = T3 * dn_dVb - dT10_dVg * (dVth_dVb + Vgst * dn_dVb / n);
    dT10_dVd = T3 * dn_dVd - dT10_dVg * (dVth_dVd + Vgst * dn_dVd / n);
    dT10_dVg *= dVgs_eff_dVg;
  }
  T1 = P.voffcbn - (1.0 - P.mstar) * Vgst;
  T2 = T1 / T0;
  if (T2 < -kExpThr) {
    T3 = M.coxe * kMinExp / P.cdep0;
    T9 = P.mstar + T3 * n;
    dT9_dVg = 0.0; dT9_dVd = dn_dVd * T3; dT9_dVb = dn_dVb * T3;
  } else if (T2 > kExpThr) {
    T3 = M.coxe * kMaxExp / P.cdep0;
    T9 = P.mstar + T3 * n;
    dT9_dVg = 0.0; dT9_dVd = dn_dVd * T3; dT9_dVb = dn_dVb * T3;
  } else {
    const real ExpVgst = exp(T2);
    T3 = M.coxe / P.cdep0;
    T4 = T3 * ExpVgst;
    T5 = T1 * T4 / T0;
    T9 = P.mstar + n * T4;
    dT9_dVg = T3 * (P.mstar - 1.0) * ExpVgst / Vtm;
    dT9_dVb = T4 * dn_dVb - dT9_dVg * dVth_dVb - T5 * dn_dVb;
    dT9_dVd = T4 * dn_dVd - dT9_dVg * dVth_dVd - T5 * dn_dVd;
    dT9_dVg *= dVgs_eff_dVg;
  }
  const real Vgsteff = T10 / T9;
  W.Vgsteff = Vgsteff;
  T11 = T9 * T9;
  const real dVgsteff_dVg = (T9 * dT10_dVg - T10 * dT9_dVg) / T11;
  const real dVgsteff_dVd = (T9 * dT10_dVd - T10 * dT9_dVd) / T11;
  const real dVgsteff_dVb = (T9 * dT10_dVb - T10 * dT9_dVb) / T11;

  XB_SYNC_POINT(2);
  // ---- effective channel width & parasitic Rds ------------------------------
  T9 = sqrtPhis - P.sqrtPhi;
  real Weff = P.weff - 2.0 * (P.dwg * Vgsteff + P.dwb * T9);
  real dWeff_dVg = -2.0 * P.dwg;
  real dWeff_dVb = -2.0 * P.dwb * dsqrtPhis_dVb;
  if (Weff < 2.0e-8) {
    T0 = 1.0 / (6.0e-8 - 2.0 * Weff);
    Weff = 2.0e-8 * (4.0e-8 - Weff) * T0;
    T0 *= T0 * 4.0e-16;
    dWeff_dVg *= T0;
    dWeff_dVb *= T0;
  }
  real Rds, dRds_dVg, dRds_dVb;
  if (M.rdsMod == 1) {
    Rds = dRds_dVg = dRds_dVb = 0.0;
  } else {
    T0 = 1.0 + P.prwg * Vgsteff;
    dT0_dVg = -P.prwg / T0 / T0;
    T1 = P.prwb * T9;
    dT1_dVb = P.prwb * dsqrtPhis_dVb;
    T2 = 1.0 / T0 + T1;
    T3 = T2 + sqrt(T2 * T2 + 0.01);
    dT3_dVg = 1.0 + T2 / (T3 - T2);
    dT3_dVb = dT3_dVg * dT1_dVb;
    dT3_dVg *= dT0_dVg;
    T4 = P.rds0 * 0.5;
    Rds = P.rdswmin + T3 * T4;
    dRds_dVg = T4 * dT3_dVg;
    dRds_dVb = T4 * dT3_dVb;
    W.grdsw = (Rds > 0.0) ? (kV47 ? 1.0 / Rds * I.nf : 1.0 / Rds) : 0.0;      // no nf factor in B4p61.C
  }

  XB_SYNC_POINT(2);
  // ---- bulk charge effect (Abulk) -------------------------------------------
  real Abulk, dAbulk_dVb, dAbulk_dVg, Abulk0, dAbulk0_dVb;
  {
    T9 = 0.5 * P.k1ox * Lpe_Vb / sqrtPhis;
    T1 = T9 + I.k2ox - P.k3b * Vth_NarrowW;
    dT1_dVb = -T9 / sqrtPhis * dsqrtPhis_dVb;
    T9 = sqrt(P.xj * Xdep);
    const real tmp1 = Leff + 2.0 * T9;
    T5 = Leff / tmp1;
    const real tmp2 = P.a0 * T5;
    const real tmp3 = P.weff + P.b1;
    const real tmp4 = P.b0 / tmp3;
    T2 = tmp2 + tmp4;
    dT2_dVb = -T9 / tmp1 / Xdep * dXdep_dVb;
    T6 = T5 * T5;
    T7 = T5 * T6;
    Abulk0 = 1.0 + T1 * T2;
    dAbulk0_dVb = T1 * tmp2 * dT2_dVb + T2 * dT1_dVb;
    T8 = P.ags * P.a0 * T7;
    dAbulk_dVg = -T1 * T8;
    Abulk = Abulk0 + dAbulk_dVg * Vgsteff;
    dAbulk_dVb = dAbulk0_dVb - T8 * Vgsteff * (dT1_dVb + 3.0 * T1 * dT2_dVb);
    if (Abulk0 < 0.1) {
      T9 = 1.0 / (3.0 - 20.0 * Abulk0);
      Abulk0 = (0.2 - Abulk0) * T9;
      dAbulk0_dVb *= T9 * T9;
    }
    if (Abulk < 0.1) {
      T9 = 1.0 / (3.0 - 20.0 * Abulk);
      Abulk = (0.2 - Abulk) * T9;
      T10 = T9 * T9;
      dAbulk_dVb *= T10;
      dAbulk_dVg *= T10;
    }
    T2 = P.keta * Vbseff;
    if (T2 >= -0.9) {
      T0 = 1.0 / (1.0 + T2);
      dT0_dVb = -P.keta * T0 * T0;
    } else {
      T1 = 1.0 / (0.8 + T2);
      T0 = (17.0 + 20.0 * T2) * T1;
      dT0_dVb = -P.keta * T1 * T1;
    }
    dAbulk_dVg *= T0;
    dAbulk_dVb = dAbulk_dVb * T0 + Abulk * dT0_dVb;
    dAbulk0_dVb = dAbulk0_dVb * T0 + Abulk0 * dT0_dVb;
    Abulk *= T0;
    Abulk0 *= T0;
  }

  XB_SYNC_POINT(1);
  // ---- mobility ---------------------------------------------------------------
  real Denomi, dDenomi_dVg, dDenomi_dVd, dDenomi_dVb;
  if (M.mtrlMod && (!kV47 || M.mtrlCompatMod == 0))      // mtrlCompatMod: 4.7 and later
    T14 = 2.0 * M.dtype * (M.phig - M.easub - 0.5 * M.Eg0 + 0.45);
  else
    T14 = 0.0;
  if (M.mobMod == 0) {
    T0 = Vgsteff + Vth + Vth - T14;
    T2 = P.ua + P.uc * Vbseff;
    T3 = T0 / toxe;
    T12 = sqrt(Vth * Vth + 0.0001);
    T9 = 1.0 / (Vgsteff + 2 * T12);
    T10 = T9 * toxe;
    T8 = P.ud * T10 * T10 * Vth;
    T6 = T8 * Vth;
    T5 = T3 * (T2 + P.ub * T3) + T6;
    T7 = -2.0 * T6 * T9;
    T11 = T7 * Vth / T12;
    dDenomi_dVg = (T2 + 2.0 * P.ub * T3) / toxe;
    T13 = 2.0 * (dDenomi_dVg + T11 + T8);
    dDenomi_dVd = T13 * dVth_dVd;
    dDenomi_dVb = T13 * dVth_dVb + P.uc * T3;
    dDenomi_dVg += T7;
  } else if (M.mobMod == 1) {
    T0 = Vgsteff + Vth + Vth - T14;
    T2 = 1.0 + P.uc * Vbseff;
    T3 = T0 / toxe;
    T4 = T3 * (P.ua + P.ub * T3);
    T12 = sqrt(Vth * Vth + 0.0001);
    T9 = 1.0 / (Vgsteff + 2 * T12);
    T10 = T9 * toxe;
    T8 = P.ud * T10 * T10 * Vth;
    T6 = T8 * Vth;
    T5 = T4 * T2 + T6;
    T7 = -2.0 * T6 * T9;
    T11 = T7 * Vth / T12;
    dDenomi_dVg = (P.ua + 2.0 * P.ub * T3) * T2 / toxe;
    T13 = 2.0 * (dDenomi_dVg + T11 + T8);
    dDenomi_dVd = T13 * dVth_dVd;
    dDenomi_dVb = T13 * dVth_dVb + P.uc * T4;
    dDenomi_dVg += T7;
  } else if (M.mobMod == 2 || !kV47) {      // B4p61.C: plain else
    T0 = (Vgsteff + I.vtfbphi1) / toxe;
    T1 = exp(P.eu * log(T0));
    dT1_dVg = T1 * P.eu / T0 / toxe;
    T2 = P.ua + P.uc * Vbseff;
    T3 = T0 / toxe;
    T12 = sqrt(Vth * Vth + 0.0001);
    T9 = 1.0 / (Vgsteff + 2 * T12);
    T10 = T9 * toxe;
    T8 = P.ud * T10 * T10 * Vth;
    T6 = T8 * Vth;
    T5 = T1 * T2 + T6;
    T7 = -2.0 * T6 * T9;
    T11 = T7 * Vth / T12;
    dDenomi_dVg = T2 * dT1_dVg + T7;
    T13 = 2.0 * (T11 + T8);
    dDenomi_dVd = T13 * dVth_dVd;
    dDenomi_dVb = T13 * dVth_dVb + T1 * P.uc;
  } else if (M.mobMod == 4 && kV48) {      // mobMod 4-6: 4.8 and later (B4p82.C:4266-4319)
    T0 = Vgsteff + I.vtfbphi1 - T14;
    T2 = P.ua + P.uc * Vbseff;
    T3 = T0 / toxe;
    T12 = sqrt(I.vtfbphi1 * I.vtfbphi1 + 0.0001);
    T9 = 1.0 / (Vgsteff + 2 * T12);
    T10 = T9 * toxe;
    T8 = P.ud * T10 * T10 * I.vtfbphi1;
    T6 = T8 * I.vtfbphi1;
    T5 = T3 * (T2 + P.ub * T3) + T6;
    T7 = -2.0 * T6 * T9;
    dDenomi_dVg = (T2 + 2.0 * P.ub * T3) / toxe;
    dDenomi_dVd = 0.0;
    dDenomi_dVb = P.uc * T3;
    dDenomi_dVg += T7;
  } else if (M.mobMod == 5 && kV48) {
    T0 = Vgsteff + I.vtfbphi1 - T14;
    T2 = 1.0 + P.uc * Vbseff;
    T3 = T0 / toxe;
    T4 = T3 * (P.ua + P.ub * T3);
    T12 = sqrt(I.vtfbphi1 * I.vtfbphi1 + 0.0001);
    T9 = 1.0 / (Vgsteff + 2 * T12);
    T10 = T9 * toxe;
    T8 = P.ud * T10 * T10 * I.vtfbphi1;
    T6 = T8 * I.vtfbphi1;
    T5 = T4 * T2 + T6;
    T7 = -2.0 * T6 * T9;
    dDenomi_dVg = (P.ua + 2.0 * P.ub * T3) * T2 / toxe;
    dDenomi_dVd = 0.0;
    dDenomi_dVb = P.uc * T4;
    dDenomi_dVg += T7;
  } else if (M.mobMod == 6 && kV48) {
    T0 = (Vgsteff + I.vtfbphi1) / toxe;
    T1 = exp(P.eu * log(T0));
    dT1_dVg = T1 * P.eu / T0 / toxe;
    T2 = P.ua + P.uc * Vbseff;
    T12 = sqrt(I.vtfbphi1 * I.vtfbphi1 + 0.0001);
    T9 = 1.0 / (Vgsteff + 2 * T12);
    T10 = T9 * toxe;
    T8 = P.ud * T10 * T10 * I.vtfbphi1;
    T6 = T8 * I.vtfbphi1;
    T5 = T1 * T2 + T6;
    T7 = -2.0 * T6 * T9;
    dDenomi_dVg = T2 * dT1_dVg + T7;
    dDenomi_dVd = 0;
    dDenomi_dVb = T1 * P.uc;
  } else {
    T0 = (Vgsteff + I.vtfbphi1) * 1.0e-8 / toxe / 6.0;
    T1 = exp(P.eu * log(T0));
    dT1_dVg = T1 * P.eu * 1.0e-8 / T0 / toxe / 6.0;
    T2 = P.ua + P.uc * Vbseff;
    const real VgsteffVth = P.VgsteffVth;
    T10 = exp(P.ucs * log(0.5 + 0.5 * Vgsteff / VgsteffVth));
    T11 = P.ud / T10;
    const real dT11_dVg = -0.5 * P.ucs * T11 / (0.5 + 0.5 * Vgsteff / VgsteffVth) / VgsteffVth;
    dDenomi_dVg = T2 * dT1_dVg + dT11_dVg;
    dDenomi_dVd = 0.0;
    dDenomi_dVb = T1 * P.uc;
    T5 = T1 * T2 + T11;
  }
  if (T5 >= -0.8) {
    Denomi = 1.0 + T5;
  } else {
    T9 = 1.0 / (7.0 + 10.0 * T5);
    Denomi = (0.6 + T5) * T9;
    T9 *= T9;
    dDenomi_dVg *= T9;
    dDenomi_dVd *= T9;
    dDenomi_dVb *= T9;
  }
  const real ueff = I.u0temp / Denomi;
  W.ueff = ueff;
  T9 = -ueff / Denomi;
  const real dueff_dVg = T9 * dDenomi_dVg;
  const real dueff_dVd = T9 * dDenomi_dVd;
  const real dueff_dVb = T9 * dDenomi_dVb;

  XB_SYNC_POINT(1);
  // ---- saturation voltage ------------------------------------------------------
  const real WVCox = Weff * I.vsattemp * M.coxe;
  const real WVCoxRds = WVCox * Rds;
  real Esat = 2.0 * I.vsattemp / ueff;
  real EsatL = Esat * Leff;
  T0 = -EsatL / ueff;
  real dEsatL_dVg = T0 * dueff_dVg;
  real dEsatL_dVd = T0 * dueff_dVd;
  real dEsatL_dVb = T0 * dueff_dVb;

  real Lambda, dLambda_dVg;
  if (P.a1 == 0.0) {
    Lambda = P.a2;
    dLambda_dVg = 0.0;
  } else if (P.a1 > 0.0) {
    T0 = 1.0 - P.a2;
    T1 = T0 - P.a1 * Vgsteff - 0.0001;
    T2 = sqrt(T1 * T1 + 0.0004 * T0);
    Lambda = P.a2 + T0 - 0.5 * (T1 + T2);
    dLambda_dVg = 0.5 * P.a1 * (1.0 + T1 / T2);
  } else {
    T1 = P.a2 + P.a1 * Vgsteff - 0.0001;
    T2 = sqrt(T1 * T1 + 0.0004 * P.a2);
    Lambda = 0.5 * (T1 + T2);
    dLambda_dVg = 0.5 * P.a1 * (1.0 + T1 / T2);
  }

  const real Vgst2Vtm = Vgsteff + 2.0 * Vtm;
  real tmp1, tmp2, tmp3;
  if (Rds > 0) {
    tmp2 = dRds_dVg / Rds + dWeff_dVg / Weff;
    tmp3 = dRds_dVb / Rds + dWeff_dVb / Weff;
  } else {
    tmp2 = dWeff_dVg / Weff;
    tmp3 = dWeff_dVb / Weff;
  }
  real Vdsat, dVdsat_dVg, dVdsat_dVd, dVdsat_dVb;
  if ((Rds == 0.0) && (Lambda == 1.0)) {
    T0 = 1.0 / (Abulk * EsatL + Vgst2Vtm);
    tmp1 = 0.0;
    T1 = T0 * T0;
    T2 = Vgst2Vtm * T0;
    T3 = EsatL * Vgst2Vtm;
    Vdsat = T3 * T0;
    dT0_dVg = -(Abulk * dEsatL_dVg + EsatL * dAbulk_dVg + 1.0) * T1;
    dT0_dVd = -(Abulk * dEsatL_dVd) * T1;
    dT0_dVb = -(Abulk * dEsatL_dVb + dAbulk_dVb * EsatL) * T1;
    dVdsat_dVg = T3 * dT0_dVg + T2 * dEsatL_dVg + EsatL * T0;
    dVdsat_dVd = T3 * dT0_dVd + T2 * dEsatL_dVd;
    dVdsat_dVb = T3 * dT0_dVb + T2 * dEsatL_dVb;
  } else {
    tmp1 = dLambda_dVg / (Lambda * Lambda);
    T9 = Abulk * WVCoxRds;
    T8 = Abulk * T9;
    T7 = Vgst2Vtm * T9;
    T6 = Vgst2Vtm * WVCoxRds;
    T0 = 2.0 * Abulk * (T9 - 1.0 + 1.0 / Lambda);
    dT0_dVg = 2.0 * (T8 * tmp2 - Abulk * tmp1 + (2.0 * T9 + 1.0 / Lambda - 1.0) * dAbulk_dVg);
    dT0_dVb = 2.0 * (T8 * (2.0 / Abulk * dAbulk_dVb + tmp3) + (1.0 / Lambda - 1.0) * dAbulk_dVb);
    dT0_dVd = 0.0;
    T1 = Vgst2Vtm * (2.0 / Lambda - 1.0) + Abulk * EsatL + 3.0 * T7;
    dT1_dVg = (2.0 / Lambda - 1.0) - 2.0 * Vgst2Vtm * tmp1 + Abulk * dEsatL_dVg + EsatL * dAbulk_dVg
            + 3.0 * (T9 + T7 * tmp2 + T6 * dAbulk_dVg);
    dT1_dVb = Abulk * dEsatL_dVb + EsatL * dAbulk_dVb + 3.0 * (T6 * dAbulk_dVb + T7 * tmp3);
    dT1_dVd = Abulk * dEsatL_dVd;
    T2 = Vgst2Vtm * (EsatL + 2.0 * T6);
    dT2_dVg = EsatL + Vgst2Vtm * dEsatL_dVg + T6 * (4.0 + 2.0 * Vgst2Vtm * tmp2);
    dT2_dVb = Vgst2Vtm * (dEsatL_dVb + 2.0 * T6 * tmp3);
    dT2_dVd = Vgst2Vtm * dEsatL_dVd;
    T3 = sqrt(T1 * T1 - 2.0 * T0 * T2);
    Vdsat = (T1 - T3) / T0;
    dT3_dVg = (T1 * dT1_dVg - 2.0 * (T0 * dT2_dVg + T2 * dT0_dVg)) / T3;
    dT3_dVd = (T1 * dT1_dVd - 2.0 * (T0 * dT2_dVd + T2 * dT0_dVd)) / T3;
    dT3_dVb = (T1 * dT1_dVb - 2.0 * (T0 * dT2_dVb + T2 * dT0_dVb)) / T3;
    dVdsat_dVg = (dT1_dVg - (T1 * dT1_dVg - dT0_dVg * T2 - T0 * dT2_dVg) / T3 - Vdsat * dT0_dVg) / T0;
    dVdsat_dVb = (dT1_dVb - (T1 * dT1_dVb - dT0_dVb * T2 - T0 * dT2_dVb) / T3 - Vdsat * dT0_dVb) / T0;
    dVdsat_dVd = (dT1_dVd - (T1 * dT1_dVd - T0 * dT2_dVd) / T3) / T0;
  }
  W.Vdsat = Vdsat;

  XB_SYNC_POINT(2);
  // ---- effective Vds -------------------------------------------------------------
  real Vdseff, dVdseff_dVg, dVdseff_dVd, dVdseff_dVb;
  T1 = Vdsat - Vds - P.delta;
  dT1_dVg = dVdsat_dVg;
  dT1_dVd = dVdsat_dVd - 1.0;
  dT1_dVb = dVdsat_dVb;
  T2 = sqrt(T1 * T1 + 4.0 * P.delta * Vdsat);
  T0 = T1 / T2;
  T9 = 2.0 * P.delta;
  T3 = T9 / T2;
  dT2_dVg = T0 * dT1_dVg + T3 * dVdsat_dVg;
  dT2_dVd = T0 * dT1_dVd + T3 * dVdsat_dVd;
  dT2_dVb = T0 * dT1_dVb + T3 * dVdsat_dVb;
  if (T1 >= 0.0) {
    Vdseff = Vdsat - 0.5 * (T1 + T2);
    dVdseff_dVg = dVdsat_dVg - 0.5 * (dT1_dVg + dT2_dVg);
    dVdseff_dVd = dVdsat_dVd - 0.5 * (dT1_dVd + dT2_dVd);
    dVdseff_dVb = dVdsat_dVb - 0.5 * (dT1_dVb + dT2_dVb);
  } else {
    T4 = T9 / (T2 - T1);
    T5 = 1.0 - T4;
    T6 = Vdsat * T4 / (T2 - T1);
    Vdseff = Vdsat * T5;
    dVdseff_dVg = dVdsat_dVg * T5 + T6 * (dT2_dVg - dT1_dVg);
    dVdseff_dVd = dVdsat_dVd * T5 + T6 * (dT2_dVd - dT1_dVd);
    dVdseff_dVb = dVdsat_dVb * T5 + T6 * (dT2_dVb - dT1_dVb);
  }
  if (Vds == 0.0) {
    Vdseff = 0.0;
    dVdseff_dVg = 0.0;
    dVdseff_dVb = 0.0;
  }
  if (Vdseff > Vds) Vdseff = Vds;
  const real diffVds = Vds - Vdseff;
  W.Vdseff = Vdseff;

  XB_SYNC_POINT(2);
  // ---- velocity overshoot (lambda) ------------------------------------------------
  if (M.lambdaGiven && (M.lambda > 0.0)) {
    T1 = Leff * ueff;
    T2 = P.lambda / T1;
    T3 = -T2 / T1 * Leff;
    dT2_dVd = T3 * dueff_dVd;
    dT2_dVg = T3 * dueff_dVg;
    dT2_dVb = T3 * dueff_dVb;
    T5 = 1.0 / (Esat * P.litl);
    T4 = -T5 / EsatL;
    dT5_dVg = dEsatL_dVg * T4;
    dT5_dVd = dEsatL_dVd * T4;
    dT5_dVb = dEsatL_dVb * T4;
    T6 = 1.0 + diffVds * T5;
    dT6_dVg = dT5_dVg * diffVds - dVdseff_dVg * T5;
    dT6_dVd = dT5_dVd * diffVds + (1.0 - dVdseff_dVd) * T5;
    dT6_dVb = dT5_dVb * diffVds - dVdseff_dVb * T5;
    T7 = 2.0 / (T6 * T6 + 1.0);
    T8 = 1.0 - T7;
    T9 = T6 * T7 * T7;
    dT8_dVg = T9 * dT6_dVg;
    dT8_dVd = T9 * dT6_dVd;
    dT8_dVb = T9 * dT6_dVb;
    T10 = 1.0 + T2 * T8;
    dT10_dVg = dT2_dVg * T8 + T2 * dT8_dVg;
    dT10_dVd = dT2_dVd * T8 + T2 * dT8_dVd;
    dT10_dVb = dT2_dVb * T8 + T2 * dT8_dVb;
    if (T10 == 1.0) dT10_dVg = dT10_dVd = dT10_dVb = 0.0;
    dEsatL_dVg *= T10; dEsatL_dVg += EsatL * dT10_dVg;
    dEsatL_dVd *= T10; dEsatL_dVd += EsatL * dT10_dVd;
    dEsatL_dVb *= T10; dEsatL_dVb += EsatL * dT10_dVb;
    EsatL *= T10;
    if (kV47) Esat = EsatL / Leff;      // not in B4p61.C
  }
  W.EsatL = EsatL;

  XB_SYNC_POINT(2);
  // ---- Vasat -------------------------------------------------------------------
  real Vasat, dVasat_dVg, dVasat_dVb, dVasat_dVd;
  {
    const real tmp4 = 1.0 - 0.5 * Abulk * Vdsat / Vgst2Vtm;
    T9 = WVCoxRds * Vgsteff;
    T8 = T9 / Vgst2Vtm;
    T0 = EsatL + Vdsat + 2.0 * T9 * tmp4;
    T7 = 2.0 * WVCoxRds * tmp4;
    dT0_dVg = dEsatL_dVg + dVdsat_dVg + T7 * (1.0 + tmp2 * Vgsteff)
            - T8 * (Abulk * dVdsat_dVg - Abulk * Vdsat / Vgst2Vtm + Vdsat * dAbulk_dVg);
    dT0_dVb = dEsatL_dVb + dVdsat_dVb + T7 * tmp3 * Vgsteff
            - T8 * (dAbulk_dVb * Vdsat + Abulk * dVdsat_dVb);
    dT0_dVd = dEsatL_dVd + dVdsat_dVd - T8 * Abulk * dVdsat_dVd;
    T9 = WVCoxRds * Abulk;
    T1 = 2.0 / Lambda - 1.0 + T9;
    dT1_dVg = -2.0 * tmp1 + WVCoxRds * (Abulk * tmp2 + dAbulk_dVg);
    dT1_dVb = dAbulk_dVb * WVCoxRds + T9 * tmp3;
    Vasat = T0 / T1;
    dVasat_dVg = (dT0_dVg - Vasat * dT1_dVg) / T1;
    dVasat_dVb = (dT0_dVb - Vasat * dT1_dVb) / T1;
    dVasat_dVd = dT0_dVd / T1;
  }

  XB_SYNC_POINT(2);
  // ---- effective oxide capacitance and channel conductance ---------------------------
  real Idl, dIdl_dVg, dIdl_dVd, dIdl_dVb, dCoxeff_dVg;
  real beta, dbeta_dVg, dbeta_dVd, dbeta_dVb, CoxeffWovL;
  {
    tmp1 = I.vtfbphi2;
    tmp2 = 2.0e8 * I.toxp;
    dT0_dVg = 1.0 / tmp2;
    T0 = (Vgsteff + tmp1) * dT0_dVg;
    tmp3 = exp(M.bdos * 0.7 * log(T0));
    T1 = 1.0 + tmp3;
    T2 = M.bdos * 0.7 * tmp3 / T0;
    const real Tcen = M.ados * 1.9e-9 / T1;
    const real dTcen_dVg = -Tcen * T2 * dT0_dVg / T1;
    const real Coxeff = epssub * I.coxp / (epssub + I.coxp * Tcen);
    W.Coxeff = Coxeff;
    dCoxeff_dVg = -Coxeff * Coxeff * dTcen_dVg / epssub;
    CoxeffWovL = Coxeff * Weff / Leff;
    beta = ueff * CoxeffWovL;
    W.beta = beta;
    T3 = ueff / Leff;
    dbeta_dVg = CoxeffWovL * dueff_dVg + T3 * (Weff * dCoxeff_dVg + Coxeff * dWeff_dVg);
    dbeta_dVd = CoxeffWovL * dueff_dVd;
    dbeta_dVb = CoxeffWovL * dueff_dVb + T3 * Coxeff * dWeff_dVb;

    W.AbovVgst2Vtm = Abulk / Vgst2Vtm;
    T0 = 1.0 - 0.5 * Vdseff * W.AbovVgst2Vtm;
    dT0_dVg = -0.5 * (Abulk * dVdseff_dVg - Abulk * Vdseff / Vgst2Vtm + Vdseff * dAbulk_dVg) / Vgst2Vtm;
    dT0_dVd = -0.5 * Abulk * dVdseff_dVd / Vgst2Vtm;
    dT0_dVb = -0.5 * (Abulk * dVdseff_dVb + dAbulk_dVb * Vdseff) / Vgst2Vtm;
    const real fgche1 = Vgsteff * T0;
    const real dfgche1_dVg = Vgsteff * dT0_dVg + T0;
    const real dfgche1_dVd = Vgsteff * dT0_dVd;
    const real dfgche1_dVb = Vgsteff * dT0_dVb;
    T9 = Vdseff / EsatL;
    const real fgche2 = 1.0 + T9;
    const real dfgche2_dVg = (dVdseff_dVg - T9 * dEsatL_dVg) / EsatL;
    const real dfgche2_dVd = (dVdseff_dVd - T9 * dEsatL_dVd) / EsatL;
    const real dfgche2_dVb = (dVdseff_dVb - T9 * dEsatL_dVb) / EsatL;
    const real gche = beta * fgche1 / fgche2;
    const real dgche_dVg = (beta * dfgche1_dVg + fgche1 * dbeta_dVg - gche * dfgche2_dVg) / fgche2;
    const real dgche_dVd = (beta * dfgche1_dVd + fgche1 * dbeta_dVd - gche * dfgche2_dVd) / fgche2;
    const real dgche_dVb = (beta * dfgche1_dVb + fgche1 * dbeta_dVb - gche * dfgche2_dVb) / fgche2;
    T0 = 1.0 + gche * Rds;
    Idl = gche / T0;
    T1 = (1.0 - Idl * Rds) / T0;
    T2 = Idl * Idl;
    dIdl_dVg = T1 * dgche_dVg - T2 * dRds_dVg;
    dIdl_dVd = T1 * dgche_dVd;
    dIdl_dVb = T1 * dgche_dVb - T2 * dRds_dVb;
  }

  XB_SYNC_POINT(1);
  // ---- output-resistance components: FP, PvagTerm, VACLM, VADIBL, VADITS, VASCBE ------
  real FP, dFP_dVg;
  if (P.fprout <= 0.0) { FP = 1.0; dFP_dVg = 0.0; }
  else {
    T9 = P.fprout * sqrt(Leff) / Vgst2Vtm;
    FP = 1.0 / (1.0 + T9);
    dFP_dVg = FP * FP * T9 / Vgst2Vtm;
  }
  real PvagTerm, dPvagTerm_dVg, dPvagTerm_dVb, dPvagTerm_dVd;
  T8 = P.pvag / EsatL;
  T9 = T8 * Vgsteff;
  if (T9 > -0.9) {
    PvagTerm = 1.0 + T9;
    dPvagTerm_dVg = T8 * (1.0 - Vgsteff * dEsatL_dVg / EsatL);
    dPvagTerm_dVb = -T9 * dEsatL_dVb / EsatL;
    dPvagTerm_dVd = -T9 * dEsatL_dVd / EsatL;
  } else {
    T4 = 1.0 / (17.0 + 20.0 * T9);
    PvagTerm = (0.8 + T9) * T4;
    T4 *= T4;
    dPvagTerm_dVg = T8 * (1.0 - Vgsteff * dEsatL_dVg / EsatL) * T4;
    T9 *= T4 / EsatL;
    dPvagTerm_dVb = -T9 * dEsatL_dVb;
    dPvagTerm_dVd = -T9 * dEsatL_dVd;
  }
  real Cclm, dCclm_dVg, dCclm_dVd, dCclm_dVb, VACLM, dVACLM_dVg, dVACLM_dVd, dVACLM_dVb;
  if ((P.pclm > kMinExp) && (diffVds > 1.0e-10)) {
    T0 = 1.0 + Rds * Idl;
    dT0_dVg = dRds_dVg * Idl + Rds * dIdl_dVg;
    dT0_dVd = Rds * dIdl_dVd;
    dT0_dVb = dRds_dVb * Idl + Rds * dIdl_dVb;
    T2 = Vdsat / Esat;
    T1 = Leff + T2;
    dT1_dVg = (dVdsat_dVg - T2 * dEsatL_dVg / Leff) / Esat;
    dT1_dVd = (dVdsat_dVd - T2 * dEsatL_dVd / Leff) / Esat;
    dT1_dVb = (dVdsat_dVb - T2 * dEsatL_dVb / Leff) / Esat;
    Cclm = FP * PvagTerm * T0 * T1 / (P.pclm * P.litl);
    dCclm_dVg = Cclm * (dFP_dVg / FP + dPvagTerm_dVg / PvagTerm + dT0_dVg / T0 + dT1_dVg / T1);
    dCclm_dVb = Cclm * (dPvagTerm_dVb / PvagTerm + dT0_dVb / T0 + dT1_dVb / T1);
    dCclm_dVd = Cclm * (dPvagTerm_dVd / PvagTerm + dT0_dVd / T0 + dT1_dVd / T1);
    VACLM = Cclm * diffVds;
    dVACLM_dVg = dCclm_dVg * diffVds - dVdseff_dVg * Cclm;
    dVACLM_dVb = dCclm_dVb * diffVds - dVdseff_dVb * Cclm;
    dVACLM_dVd = dCclm_dVd * diffVds + (1.0 - dVdseff_dVd) * Cclm;
  } else {
    VACLM = Cclm = kMaxExp;
    dVACLM_dVd = dVACLM_dVg = dVACLM_dVb = 0.0;
    dCclm_dVd = dCclm_dVg = dCclm_dVb = 0.0;
  }
  real VADIBL, dVADIBL_dVg, dVADIBL_dVd, dVADIBL_dVb;
  if (P.thetaRout > kMinExp) {
    T8 = Abulk * Vdsat;
    T0 = Vgst2Vtm * T8;
    dT0_dVg = Vgst2Vtm * Abulk * dVdsat_dVg + T8 + Vgst2Vtm * Vdsat * dAbulk_dVg;
    dT0_dVb = Vgst2Vtm * (dAbulk_dVb * Vdsat + Abulk * dVdsat_dVb);
    dT0_dVd = Vgst2Vtm * Abulk * dVdsat_dVd;
    T1 = Vgst2Vtm + T8;
    dT1_dVg = 1.0 + Abulk * dVdsat_dVg + Vdsat * dAbulk_dVg;
    dT1_dVb = Abulk * dVdsat_dVb + dAbulk_dVb * Vdsat;
    dT1_dVd = Abulk * dVdsat_dVd;
    T9 = T1 * T1;
    T2 = P.thetaRout;
    VADIBL = (Vgst2Vtm - T0 / T1) / T2;
    dVADIBL_dVg = (1.0 - dT0_dVg / T1 + T0 * dT1_dVg / T9) / T2;
    dVADIBL_dVb = (-dT0_dVb / T1 + T0 * dT1_dVb / T9) / T2;
    dVADIBL_dVd = (-dT0_dVd / T1 + T0 * dT1_dVd / T9) / T2;
    T7 = P.pdiblb * Vbseff;
    if (T7 >= -0.9) {
      T3 = 1.0 / (1.0 + T7);
      VADIBL *= T3;
      dVADIBL_dVg *= T3;
      dVADIBL_dVb = (dVADIBL_dVb - VADIBL * P.pdiblb) * T3;
      dVADIBL_dVd *= T3;
    } else {
      T4 = 1.0 / (0.8 + T7);
      T3 = (17.0 + 20.0 * T7) * T4;
      dVADIBL_dVg *= T3;
      dVADIBL_dVb = dVADIBL_dVb * T3 - VADIBL * P.pdiblb * T4 * T4;
      dVADIBL_dVd *= T3;
      VADIBL *= T3;
    }
    dVADIBL_dVg = dVADIBL_dVg * PvagTerm + VADIBL * dPvagTerm_dVg;
    dVADIBL_dVb = dVADIBL_dVb * PvagTerm + VADIBL * dPvagTerm_dVb;
    dVADIBL_dVd = dVADIBL_dVd * PvagTerm + VADIBL * dPvagTerm_dVd;
    VADIBL *= PvagTerm;
  } else {
    VADIBL = kMaxExp;
    dVADIBL_dVd = dVADIBL_dVg = dVADIBL_dVb = 0.0;
  }
  const real Va = Vasat + VACLM;
  const real dVa_dVg = dVasat_dVg + dVACLM_dVg;
  const real dVa_dVb = dVasat_dVb + dVACLM_dVb;
  const real dVa_dVd = dVasat_dVd + dVACLM_dVd;

  real VADITS, dVADITS_dVg, dVADITS_dVd;
  T0 = P.pditsd * Vds;
  if (T0 > kExpThr) { T1 = kMaxExp; dT1_dVd = 0; }
  else { T1 = exp(T0); dT1_dVd = T1 * P.pditsd; }
  if (P.pdits > kMinExp) {
    T2 = 1.0 + M.pditsl * Leff;
    VADITS = (1.0 + T2 * T1) / P.pdits;
    dVADITS_dVg = VADITS * dFP_dVg;
    dVADITS_dVd = FP * T2 * dT1_dVd / P.pdits;
    VADITS *= FP;
  } else {
    VADITS = kMaxExp;
    dVADITS_dVg = dVADITS_dVd = 0;
  }
  real VASCBE, dVASCBE_dVg, dVASCBE_dVd, dVASCBE_dVb;
  if ((P.pscbe2 > 0.0) && (!kV47 || P.pscbe1 >= 0.0)) {
    if (diffVds > P.pscbe1 * P.litl / kExpThr) {
      T0 = P.pscbe1 * P.litl / diffVds;
      VASCBE = Leff * exp(T0) / P.pscbe2;
      T1 = T0 * VASCBE / diffVds;
      dVASCBE_dVg = T1 * dVdseff_dVg;
      dVASCBE_dVd = -T1 * (1.0 - dVdseff_dVd);
      dVASCBE_dVb = T1 * dVdseff_dVb;
    } else {
      VASCBE = kMaxExp * Leff / P.pscbe2;
      dVASCBE_dVg = dVASCBE_dVd = dVASCBE_dVb = 0.0;
    }
  } else {
    VASCBE = kMaxExp;
    dVASCBE_dVg = dVASCBE_dVd = dVASCBE_dVb = 0.0;
  }

  XB_SYNC_POINT(2);
  // ---- Idsa: DIBL, DITS, CLM ----------------------------------------------------------
  real Idsa, dIdsa_dVg, dIdsa_dVd, dIdsa_dVb;
  T9 = diffVds / VADIBL;
  T0 = 1.0 + T9;
  Idsa = Idl * T0;
  dIdsa_dVg = T0 * dIdl_dVg - Idl * (dVdseff_dVg + T9 * dVADIBL_dVg) / VADIBL;
  dIdsa_dVd = T0 * dIdl_dVd + Idl * (1.0 - dVdseff_dVd - T9 * dVADIBL_dVd) / VADIBL;
  dIdsa_dVb = T0 * dIdl_dVb - Idl * (dVdseff_dVb + T9 * dVADIBL_dVb) / VADIBL;
  T9 = diffVds / VADITS;
  T0 = 1.0 + T9;
  dIdsa_dVg = T0 * dIdsa_dVg - Idsa * (dVdseff_dVg + T9 * dVADITS_dVg) / VADITS;
  dIdsa_dVd = T0 * dIdsa_dVd + Idsa * (1.0 - dVdseff_dVd - T9 * dVADITS_dVd) / VADITS;
  dIdsa_dVb = T0 * dIdsa_dVb - Idsa * dVdseff_dVb / VADITS;
  Idsa *= T0;
  T0 = log(Va / Vasat);
  dT0_dVg = dVa_dVg / Va - dVasat_dVg / Vasat;
  dT0_dVb = dVa_dVb / Va - dVasat_dVb / Vasat;
  dT0_dVd = dVa_dVd / Va - dVasat_dVd / Vasat;
  T1 = T0 / Cclm;
  T9 = 1.0 + T1;
  dT9_dVg = (dT0_dVg - T1 * dCclm_dVg) / Cclm;
  dT9_dVb = (dT0_dVb - T1 * dCclm_dVb) / Cclm;
  dT9_dVd = (dT0_dVd - T1 * dCclm_dVd) / Cclm;
  dIdsa_dVg = dIdsa_dVg * T9 + Idsa * dT9_dVg;
  dIdsa_dVb = dIdsa_dVb * T9 + Idsa * dT9_dVb;
  dIdsa_dVd = dIdsa_dVd * T9 + Idsa * dT9_dVd;
  Idsa *= T9;

  XB_SYNC_POINT(1);
  // ---- substrate current ----------------------------------------------------------------
  real Isub, Gbd, Gbb, Gbg;
  {
    const real tmp = P.alpha0 + P.alpha1 * Leff;
    if ((tmp <= 0.0) || (P.beta0 <= 0.0)) {
      Isub = Gbd = Gbb = Gbg = 0.0;
    } else {
      T2 = tmp / Leff;
      if (diffVds > P.beta0 / kExpThr) {
        T0 = -P.beta0 / diffVds;
        T1 = T2 * diffVds * exp(T0);
        T3 = T1 / diffVds * (T0 - 1.0);
        dT1_dVg = T3 * dVdseff_dVg;
        dT1_dVd = T3 * (dVdseff_dVd - 1.0);
        dT1_dVb = T3 * dVdseff_dVb;
      } else {
        T3 = T2 * kMinExp;
        T1 = T3 * diffVds;
        dT1_dVg = -T3 * dVdseff_dVg;
        dT1_dVd = T3 * (1.0 - dVdseff_dVd);
        dT1_dVb = -T3 * dVdseff_dVb;
      }
      T4 = Idsa * Vdseff;
      Isub = T1 * T4;
      Gbg = T1 * (dIdsa_dVg * Vdseff + Idsa * dVdseff_dVg) + T4 * dT1_dVg;
      Gbd = T1 * (dIdsa_dVd * Vdseff + Idsa * dVdseff_dVd) + T4 * dT1_dVd;
      Gbb = T1 * (dIdsa_dVb * Vdseff + Idsa * dVdseff_dVb) + T4 * dT1_dVb;
      Gbd += Gbg * dVgsteff_dVd;
      Gbb += Gbg * dVgsteff_dVb;
      Gbg *= dVgsteff_dVg;
      Gbb *= dVbseff_dVb;
    }
  }
  W.csub = Isub; W.gbbs = Gbb; W.gbgs = Gbg; W.gbds = Gbd;

  XB_SYNC_POINT(2);
  // ---- drain current with SCBE; chain rule back to terminal voltages ---------------------
  real Ids, Gm, Gds, Gmb;
  T9 = diffVds / VASCBE;
  T0 = 1.0 + T9;
  Ids = Idsa * T0;
  Gm = T0 * dIdsa_dVg - Idsa * (dVdseff_dVg + T9 * dVASCBE_dVg) / VASCBE;
  Gds = T0 * dIdsa_dVd + Idsa * (1.0 - dVdseff_dVd - T9 * dVASCBE_dVd) / VASCBE;
  Gmb = T0 * dIdsa_dVb - Idsa * (dVdseff_dVb + T9 * dVASCBE_dVb) / VASCBE;
  tmp1 = Gds + Gm * dVgsteff_dVd;
  tmp2 = Gmb + Gm * dVgsteff_dVb;
  tmp3 = Gm;
  Gm = (Ids * dVdseff_dVg + Vdseff * tmp3) * dVgsteff_dVg;
  Gds = Ids * (dVdseff_dVd + dVdseff_dVg * dVgsteff_dVd) + Vdseff * tmp1;
  Gmb = (Ids * (dVdseff_dVb + dVdseff_dVg * dVgsteff_dVb) + Vdseff * tmp2) * dVbseff_dVb;
  real cdrain = Ids * Vdseff;

  // source-end velocity limit
  if (M.vtlGiven && (M.vtl > 0.0)) {
    T12 = 1.0 / Leff / CoxeffWovL;
    T11 = T12 / Vgsteff;
    T10 = -T11 / Vgsteff;
    const real vs = cdrain * T11;
    const real dvs_dVg = Gm * T11 + cdrain * T10 * dVgsteff_dVg;
    const real dvs_dVd = Gds * T11 + cdrain * T10 * dVgsteff_dVd;
    const real dvs_dVb = Gmb * T11 + cdrain * T10 * dVgsteff_dVb;
    T0 = 2 * kMM;
    T1 = vs / (P.vtl * P.tfactor);
    real Fsevl, dFsevl_dVg, dFsevl_dVd, dFsevl_dVb;
    if (T1 > 0.0) {
      T2 = 1.0 + exp(T0 * log(T1));
      T3 = (T2 - 1.0) * T0 / vs;
      Fsevl = 1.0 / exp(log(T2) / T0);
      dT2_dVg = T3 * dvs_dVg;
      dT2_dVd = T3 * dvs_dVd;
      dT2_dVb = T3 * dvs_dVb;
      T4 = -1.0 / T0 * Fsevl / T2;
      dFsevl_dVg = T4 * dT2_dVg;
      dFsevl_dVd = T4 * dT2_dVd;
      dFsevl_dVb = T4 * dT2_dVb;
    } else {
      Fsevl = 1.0;
      dFsevl_dVg = dFsevl_dVd = dFsevl_dVb = 0.0;
    }
    Gm *= Fsevl;  Gm += cdrain * dFsevl_dVg;
    Gmb *= Fsevl; Gmb += cdrain * dFsevl_dVb;
    Gds *= Fsevl; Gds += cdrain * dFsevl_dVd;
    cdrain *= Fsevl;
  }
  W.cdrain = cdrain;
  W.gds = Gds; W.gm = Gm; W.gmbs = Gmb;
  W.IdovVds = Ids;
  {   // the floor became a model parameter in 4.8 (N_DEV_MOSFET_B4p82.C:4972; 1.0e-9 in B4p70.C / B4p61.C)
    const real floor_ = (M.versionDouble >= 4.8) ? M.idovvdsc : 1.0e-9;
    if (W.IdovVds <= floor_) W.IdovVds = floor_;
  }

  XB_SYNC_POINT(2);
  // ---- bias-dependent intrinsic-input (gate) resistance ---------------------------------------
  if ((I.rgateMod > 1) || (I.trnqsMod != 0) || (I.acnqsMod != 0)) {
    T9 = P.xrcrg2 * M.vtm;
    T0 = T9 * beta;
    dT0_dVd = (dbeta_dVd + dbeta_dVg * dVgsteff_dVd) * T9;
    dT0_dVb = (dbeta_dVb + dbeta_dVg * dVgsteff_dVb) * T9;
    dT0_dVg = dbeta_dVg * T9;
    W.gcrg = P.xrcrg1 * (T0 + Ids);
    W.gcrgd = P.xrcrg1 * (dT0_dVd + tmp1);
    W.gcrgb = P.xrcrg1 * (dT0_dVb + tmp2) * dVbseff_dVb;
    W.gcrgg = P.xrcrg1 * (dT0_dVg + tmp3) * dVgsteff_dVg;
    if (I.nf != 1.0) { W.gcrg *= I.nf; W.gcrgg *= I.nf; W.gcrgd *= I.nf; W.gcrgb *= I.nf; }
    if (I.rgateMod == 2) {
      T10 = I.grgeltd * I.grgeltd;
      T11 = I.grgeltd + W.gcrg;
      W.gcrg = I.grgeltd * W.gcrg / T11;
      T12 = T10 / T11 / T11;
      W.gcrgg *= T12; W.gcrgd *= T12; W.gcrgb *= T12;
    }
    W.gcrgs = -(W.gcrgg + W.gcrgd + W.gcrgb);
  }

  XB_SYNC_POINT(2);
  // ---- bias-dependent source / drain resistance (rdsMod) --------------------------------------
  if (M.rdsMod) {
    real dgstot_dvd, dgstot_dvg, dgstot_dvs, dgstot_dvb;
    real dgdtot_dvd, dgdtot_dvg, dgdtot_dvs, dgdtot_dvb;
    T0 = W.vgs - P.vfbsd;
    T1 = sqrt(T0 * T0 + 1.0e-4);
    W.vgs_eff = 0.5 * (T0 + T1);
    W.dvgs_eff_dvg = W.vgs_eff / T1;
    T0 = 1.0 + P.prwg * W.vgs_eff;
    real dT0_dvg = -P.prwg / T0 / T0 * W.dvgs_eff_dvg;
    T1 = -P.prwb * W.vbs;
    real dT1_dvb = -P.prwb;
    T2 = 1.0 / T0 + T1;
    T3 = T2 + sqrt(T2 * T2 + 0.01);
    real dT3_dvg = T3 / (T3 - T2);
    real dT3_dvb = dT3_dvg * dT1_dvb;
    dT3_dvg *= dT0_dvg;
    T4 = P.rs0 * 0.5;
    const real Rs = P.rswmin + T3 * T4;
    const real dRs_dvg = T4 * dT3_dvg;
    const real dRs_dvb = T4 * dT3_dvb;
    T0 = 1.0 + I.sourceConductance * Rs;
    W.gstot = I.sourceConductance / T0;
    T0 = -W.gstot * W.gstot;
    dgstot_dvd = 0.0;
    dgstot_dvg = T0 * dRs_dvg;
    dgstot_dvb = T0 * dRs_dvb;
    dgstot_dvs = -(dgstot_dvg + dgstot_dvb + dgstot_dvd);

    T0 = W.vgd - P.vfbsd;
    T1 = sqrt(T0 * T0 + 1.0e-4);
    W.vgd_eff = 0.5 * (T0 + T1);
    W.dvgd_eff_dvg = W.vgd_eff / T1;
    T0 = 1.0 + P.prwg * W.vgd_eff;
    dT0_dvg = -P.prwg / T0 / T0 * W.dvgd_eff_dvg;
    T1 = -P.prwb * W.vbd;
    dT1_dvb = -P.prwb;
    T2 = 1.0 / T0 + T1;
    T3 = T2 + sqrt(T2 * T2 + 0.01);
    dT3_dvg = T3 / (T3 - T2);
    dT3_dvb = dT3_dvg * dT1_dvb;
    dT3_dvg *= dT0_dvg;
    T4 = P.rd0 * 0.5;
    const real Rd = P.rdwmin + T3 * T4;
    const real dRd_dvg = T4 * dT3_dvg;
    const real dRd_dvb = T4 * dT3_dvb;
    T0 = 1.0 + I.drainConductance * Rd;
    W.gdtot = I.drainConductance / T0;
    T0 = -W.gdtot * W.gdtot;
    dgdtot_dvs = 0.0;
    dgdtot_dvg = T0 * dRd_dvg;
    dgdtot_dvb = T0 * dRd_dvb;
    dgdtot_dvd = -(dgdtot_dvg + dgdtot_dvb + dgdtot_dvs);

    W.gstotd = W.vses * dgstot_dvd;
    W.gstotg = W.vses * dgstot_dvg;
    W.gstots = W.vses * dgstot_dvs;
    W.gstotb = W.vses * dgstot_dvb;
    T2 = W.vdes - W.vds;
    W.gdtotd = T2 * dgdtot_dvd;
    W.gdtotg = T2 * dgdtot_dvg;
    W.gdtots = T2 * dgdtot_dvs;
    W.gdtotb = T2 * dgdtot_dvb;
  } else {
    W.gstot = W.gstotd = W.gstotg = 0.0;
    W.gstots = W.gstotb = 0.0;
    W.gdtot = W.gdtotd = W.gdtotg = 0.0;
    W.gdtots = W.gdtotb = 0.0;
  }

  XB_SYNC_POINT(2);
  // ---- GIDL / GISL ----------------------------------------------------------------------------
  {
    T0 = (M.mtrlMod == 0) ? 3.0 * toxe : M.epsrsub * toxe / epsrox;
    const real voff = (M.mtrlMod == 0) ? 0.0 : P.vfbsd;
    // the T4 clamp of the gidlMod = 1 branch exists from 4.8 on (B4p82.C:5233, :5290; not in B4p70.C)
    const real gidlclamp = (M.versionDouble >= 4.8) ? M.gidlclamp : 1.7976931348623157e308;
    if (M.gidlMod == 0 || !kV47) {      // gidlMod: 4.7 and later
      T1 = (M.mtrlMod == 0) ? (W.vds - W.vgs_eff - P.egidl) / T0
                            : (W.vds - W.vgs_eff - P.egidl + P.vfbsd) / T0;
      { const Real4 r = gidl_mod0_v(T0, T1, W.dvgs_eff_dvg, P.agidl, P.bgidl, P.cgidl, P.weffCJ, W.vbd); W.Igidl = r.a; W.ggidld = r.b; W.ggidlg = r.c; W.ggidlb = r.d; }
      T1 = (M.mtrlMod == 0) ? (-W.vds - W.vgd_eff - P.egisl) / T0
                            : (-W.vds - W.vgd_eff - P.egisl + P.vfbsd) / T0;
      { const Real4 r = gidl_mod0_v(T0, T1, W.dvgd_eff_dvg, P.agisl, P.bgisl, P.cgisl, P.weffCJ, W.vbs); W.Igisl = r.a; W.ggisls = r.b; W.ggislg = r.c; W.ggislb = r.d; }
    } else {
      T1 = (M.mtrlMod == 0) ? (-W.vds - P.rgisl * W.vgd_eff - P.egisl) / T0
                            : (-W.vds - P.rgisl * W.vgd_eff - P.egisl + P.vfbsd) / T0;
      { const Real4 r = gidl_mod1_v(T0, T1, P.rgisl, W.dvgd_eff_dvg, P.agisl, P.bgisl, P.cgisl, P.kgisl, P.fgisl,
                P.weffCJ, W.vbs, gidlclamp); W.Igisl = r.a; W.ggisls = r.b; W.ggislg = r.c; W.ggislb = r.d; }
      T1 = (M.mtrlMod == 0) ? (W.vds - P.rgidl * W.vgs_eff - P.egidl) / T0
                            : (W.vds - P.rgidl * W.vgs_eff - P.egidl + P.vfbsd) / T0;
      { const Real4 r = gidl_mod1_v(T0, T1, P.rgidl, W.dvgs_eff_dvg, P.agidl, P.bgidl, P.cgidl, P.kgidl, P.fgidl,
                P.weffCJ, W.vbd, gidlclamp); W.Igidl = r.a; W.ggidld = r.b; W.ggidlg = r.c; W.ggidlb = r.d; }
    }
    (void)voff;
  }

  XB_SYNC_POINT(1);
  // ---- gate direct-tunnelling currents ------------------------------------------------------------
  real Vfbeff = 0.0, dVfbeff_dVg = 0.0, dVfbeff_dVb = 0.0;
  real Voxacc = 0.0, dVoxacc_dVg = 0.0, dVoxacc_dVb = 0.0;
  real Voxdepinv = 0.0, dVoxdepinv_dVg = 0.0, dVoxdepinv_dVd = 0.0, dVoxdepinv_dVb = 0.0;
  real Vfb = 0.0;
  if ((M.igcMod != 0) || (M.igbMod != 0)) {
    Vfb = I.vfbzb;
    const real V3 = Vfb - Vgs_eff + Vbseff - kDelta3;
    if (Vfb <= 0.0) T0 = sqrt(V3 * V3 - 4.0 * kDelta3 * Vfb);
    else T0 = sqrt(V3 * V3 + 4.0 * kDelta3 * Vfb);
    T1 = 0.5 * (1.0 + V3 / T0);
    Vfbeff = Vfb - 0.5 * (V3 + T0);
    dVfbeff_dVg = T1 * dVgs_eff_dVg;
    dVfbeff_dVb = -T1;
    Voxacc = Vfb - Vfbeff;
    dVoxacc_dVg = -dVfbeff_dVg;
    dVoxacc_dVb = -dVfbeff_dVb;
    if (Voxacc < 0.0) Voxacc = dVoxacc_dVg = dVoxacc_dVb = 0.0;

    T0 = 0.5 * P.k1ox;
    T3 = Vgs_eff - Vfbeff - Vbseff - Vgsteff;
    if (P.k1ox == 0.0) {
      Voxdepinv = dVoxdepinv_dVg = dVoxdepinv_dVd = dVoxdepinv_dVb = 0.0;
    } else if (T3 < 0.0) {
      Voxdepinv = -T3;
      dVoxdepinv_dVg = -dVgs_eff_dVg + dVfbeff_dVg + dVgsteff_dVg;
      dVoxdepinv_dVd = dVgsteff_dVd;
      dVoxdepinv_dVb = dVfbeff_dVb + 1.0 + dVgsteff_dVb;
    } else {
      T1 = sqrt(T0 * T0 + T3);
      T2 = T0 / T1;
      Voxdepinv = P.k1ox * (T1 - T0);
      dVoxdepinv_dVg = T2 * (dVgs_eff_dVg - dVfbeff_dVg - dVgsteff_dVg);
      dVoxdepinv_dVd = -T2 * dVgsteff_dVd;
      dVoxdepinv_dVb = -T2 * (dVfbeff_dVb + 1.0 + dVgsteff_dVb);
    }
    Voxdepinv += Vgsteff;
    dVoxdepinv_dVg += dVgsteff_dVg;
    dVoxdepinv_dVd += dVgsteff_dVd;
    dVoxdepinv_dVb += dVgsteff_dVb;
  }
  const real vtm_ig = (M.tempMod < 2) ? Vtm : Vtm0;
  if (M.igcMod) {
    real VxNVt = 0.0, Vaux = 0.0, dVaux_dVg = 0.0, dVaux_dVd = 0.0, dVaux_dVb = 0.0;
    T0 = vtm_ig * P.nigc;
    if (M.igcMod == 1) {
      VxNVt = (Vgs_eff - M.dtype * I.vth0) / T0;
      if (VxNVt > kExpThr) {
        Vaux = Vgs_eff - M.dtype * I.vth0;
        dVaux_dVg = dVgs_eff_dVg; dVaux_dVd = 0.0; dVaux_dVb = 0.0;
      }
    } else if (M.igcMod == 2) {
      VxNVt = (Vgs_eff - Vth) / T0;
      if (VxNVt > kExpThr) {
        Vaux = Vgs_eff - Vth;
        dVaux_dVg = dVgs_eff_dVg; dVaux_dVd = -dVth_dVd; dVaux_dVb = -dVth_dVb;
      }
    }
    if (VxNVt < -kExpThr) {
      Vaux = T0 * log(1.0 + kMinExp);
      dVaux_dVg = dVaux_dVd = dVaux_dVb = 0.0;
    } else if ((VxNVt >= -kExpThr) && (VxNVt <= kExpThr)) {
      const real ExpVxNVt = exp(VxNVt);
      Vaux = T0 * log(1.0 + ExpVxNVt);
      dVaux_dVg = ExpVxNVt / (1.0 + ExpVxNVt);
      if (M.igcMod == 1) { dVaux_dVd = 0.0; dVaux_dVb = 0.0; }
      else if (M.igcMod == 2) {
        // 4.8: -dVaux_dVg (B4p82.C:5423); 4.7 / 4.6.1: -dVgs_eff_dVg (B4p70.C, B4p61.C same place)
        const real k_ = (M.versionDouble >= 4.8) ? dVaux_dVg : dVgs_eff_dVg;
        dVaux_dVd = -k_ * dVth_dVd; dVaux_dVb = -k_ * dVth_dVb;
      }
      dVaux_dVg *= dVgs_eff_dVg;
    }
    T2 = Vgs_eff * Vaux;
    dT2_dVg = dVgs_eff_dVg * Vaux + Vgs_eff * dVaux_dVg;
    dT2_dVd = Vgs_eff * dVaux_dVd;
    dT2_dVb = Vgs_eff * dVaux_dVb;
    T11 = P.Aechvb;
    T12 = P.Bechvb;
    T3 = P.aigc * P.cigc - P.bigc;
    T4 = P.bigc * P.cigc;
    T5 = T12 * (P.aigc + T3 * Voxdepinv - T4 * Voxdepinv * Voxdepinv);
    if (T5 > kExpThr) { T6 = kMaxExp; dT6_dVg = dT6_dVd = dT6_dVb = 0.0; }
    else if (T5 < -kExpThr) { T6 = kMinExp; dT6_dVg = dT6_dVd = dT6_dVb = 0.0; }
    else {
      T6 = exp(T5);
      dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * Voxdepinv);
      dT6_dVd = dT6_dVg * dVoxdepinv_dVd;
      dT6_dVb = dT6_dVg * dVoxdepinv_dVb;
      dT6_dVg *= dVoxdepinv_dVg;
    }
    const real Igc = T11 * T2 * T6;
    const real dIgc_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
    const real dIgc_dVd = T11 * (T2 * dT6_dVd + T6 * dT2_dVd);
    const real dIgc_dVb = T11 * (T2 * dT6_dVb + T6 * dT2_dVb);
    real Pigcd, dPigcd_dVg, dPigcd_dVd, dPigcd_dVb;
    if (M.pigcdGiven) {
      Pigcd = P.pigcd;
      dPigcd_dVg = dPigcd_dVd = dPigcd_dVb = 0.0;
    } else {
      T11 = kV47 ? -P.Bechvb : P.Bechvb * toxe;      // B4p61.C (same place)
      T12 = Vgsteff + 1.0e-20;
      T13 = T11 / T12 / T12;
      T14 = -T13 / T12;
      Pigcd = T13 * (1.0 - 0.5 * Vdseff / T12);
      dPigcd_dVg = T14 * (2.0 + 0.5 * (dVdseff_dVg - 3.0 * Vdseff / T12));
      dPigcd_dVd = 0.5 * T14 * dVdseff_dVd;
      dPigcd_dVb = 0.5 * T14 * dVdseff_dVb;
    }
    T7 = -Pigcd * Vdseff;
    dT7_dVg = -Vdseff * dPigcd_dVg - Pigcd * dVdseff_dVg;
    dT7_dVd = -Vdseff * dPigcd_dVd - Pigcd * dVdseff_dVd + dT7_dVg * dVgsteff_dVd;
    dT7_dVb = -Vdseff * dPigcd_dVb - Pigcd * dVdseff_dVb + dT7_dVg * dVgsteff_dVb;
    dT7_dVg *= dVgsteff_dVg;
    if (M.versionDouble < 4.8) dT7_dVb *= dVbseff_dVb;      // dropped in 4.8 (B4p82.C:5482)
    T8 = T7 * T7 + 2.0e-4;
    dT8_dVg = 2.0 * T7;
    dT8_dVd = dT8_dVg * dT7_dVd;
    dT8_dVb = dT8_dVg * dT7_dVb;
    dT8_dVg *= dT7_dVg;
    if (T7 > kExpThr) { T9 = kMaxExp; dT9_dVg = dT9_dVd = dT9_dVb = 0.0; }
    else if (T7 < -kExpThr) { T9 = kMinExp; dT9_dVg = dT9_dVd = dT9_dVb = 0.0; }
    else {
      T9 = exp(T7);
      dT9_dVg = T9 * dT7_dVg;
      dT9_dVd = T9 * dT7_dVd;
      dT9_dVb = T9 * dT7_dVb;
    }
    T1 = T9 - 1.0 + 1.0e-4;
    T10 = (T1 - T7) / T8;
    dT10_dVg = (dT9_dVg - dT7_dVg - T10 * dT8_dVg) / T8;
    dT10_dVd = (dT9_dVd - dT7_dVd - T10 * dT8_dVd) / T8;
    dT10_dVb = (dT9_dVb - dT7_dVb - T10 * dT8_dVb) / T8;
    W.Igcs = Igc * T10;
    const real dIgcs_dVg = dIgc_dVg * T10 + Igc * dT10_dVg;
    const real dIgcs_dVd = dIgc_dVd * T10 + Igc * dT10_dVd;
    const real dIgcs_dVb = dIgc_dVb * T10 + Igc * dT10_dVb;
    T1 = T9 - 1.0 - 1.0e-4;
    T10 = (T7 * T9 - T1) / T8;
    dT10_dVg = (dT7_dVg * T9 + (T7 - 1.0) * dT9_dVg - T10 * dT8_dVg) / T8;
    dT10_dVd = (dT7_dVd * T9 + (T7 - 1.0) * dT9_dVd - T10 * dT8_dVd) / T8;
    dT10_dVb = (dT7_dVb * T9 + (T7 - 1.0) * dT9_dVb - T10 * dT8_dVb) / T8;
    W.Igcd = Igc * T10;
    const real dIgcd_dVg = dIgc_dVg * T10 + Igc * dT10_dVg;
    const real dIgcd_dVd = dIgc_dVd * T10 + Igc * dT10_dVd;
    const real dIgcd_dVb = dIgc_dVb * T10 + Igc * dT10_dVb;
    W.gIgcsg = dIgcs_dVg; W.gIgcsd = dIgcs_dVd; W.gIgcsb = dIgcs_dVb * dVbseff_dVb;
    W.gIgcdg = dIgcd_dVg; W.gIgcdd = dIgcd_dVd; W.gIgcdb = dIgcd_dVb * dVbseff_dVb;

    // gate-to-S/D overlap tunnelling (note: overwrites vgs_eff / vgd_eff like the reference)
    T0 = W.vgs - (P.vfbsd + P.vfbsdoff);
    W.vgs_eff = sqrt(T0 * T0 + 1.0e-4);
    W.dvgs_eff_dvg = T0 / W.vgs_eff;
    T2 = W.vgs * W.vgs_eff;
    dT2_dVg = W.vgs * W.dvgs_eff_dvg + W.vgs_eff;
    T11 = P.AechvbEdgeS;
    T12 = P.BechvbEdge;
    T3 = P.aigs * P.cigs - P.bigs;
    T4 = P.bigs * P.cigs;
    T5 = T12 * (P.aigs + T3 * W.vgs_eff - T4 * W.vgs_eff * W.vgs_eff);
    if (T5 > kExpThr) { T6 = kMaxExp; dT6_dVg = 0.0; }
    else if (T5 < -kExpThr) { T6 = kMinExp; dT6_dVg = 0.0; }
    else { T6 = exp(T5); dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * W.vgs_eff) * W.dvgs_eff_dvg; }
    W.Igs = T11 * T2 * T6;
    const real dIgs_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
    const real dIgs_dVs = -dIgs_dVg;

    T0 = W.vgd - (P.vfbsd + P.vfbsdoff);
    W.vgd_eff = sqrt(T0 * T0 + 1.0e-4);
    W.dvgd_eff_dvg = T0 / W.vgd_eff;
    T2 = W.vgd * W.vgd_eff;
    dT2_dVg = W.vgd * W.dvgd_eff_dvg + W.vgd_eff;
    T11 = P.AechvbEdgeD;
    T3 = P.aigd * P.cigd - P.bigd;
    T4 = P.bigd * P.cigd;
    T5 = T12 * (P.aigd + T3 * W.vgd_eff - T4 * W.vgd_eff * W.vgd_eff);
    if (T5 > kExpThr) { T6 = kMaxExp; dT6_dVg = 0.0; }
    else if (T5 < -kExpThr) { T6 = kMinExp; dT6_dVg = 0.0; }
    else { T6 = exp(T5); dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * W.vgd_eff) * W.dvgd_eff_dvg; }
    W.Igd = T11 * T2 * T6;
    const real dIgd_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
    const real dIgd_dVd = -dIgd_dVg;
    W.gIgsg = dIgs_dVg; W.gIgss = dIgs_dVs;
    W.gIgdg = dIgd_dVg; W.gIgdd = dIgd_dVd;
  } else {
    W.Igcs = W.gIgcsg = W.gIgcsd = W.gIgcsb = 0.0;
    W.Igcd = W.gIgcdg = W.gIgcdd = W.gIgcdb = 0.0;
    W.Igs = W.gIgsg = W.gIgss = 0.0;
    W.Igd = W.gIgdg = W.gIgdd = 0.0;
  }

  if (M.igbMod) {
    real VxNVt, Vaux, dVaux_dVg, dVaux_dVd, dVaux_dVb;
    T0 = vtm_ig * P.nigbacc;
    T1 = -Vgs_eff + Vbseff + Vfb;
    VxNVt = T1 / T0;
    if (VxNVt > kExpThr) { Vaux = T1; dVaux_dVg = -dVgs_eff_dVg; dVaux_dVb = 1.0; }
    else if (VxNVt < -kExpThr) { Vaux = T0 * log(1.0 + kMinExp); dVaux_dVg = dVaux_dVb = 0.0; }
    else {
      const real ExpVxNVt = exp(VxNVt);
      Vaux = T0 * log(1.0 + ExpVxNVt);
      dVaux_dVb = ExpVxNVt / (1.0 + ExpVxNVt);
      dVaux_dVg = -dVaux_dVb * dVgs_eff_dVg;
    }
    T2 = (Vgs_eff - Vbseff) * Vaux;
    dT2_dVg = dVgs_eff_dVg * Vaux + (Vgs_eff - Vbseff) * dVaux_dVg;
    dT2_dVb = -Vaux + (Vgs_eff - Vbseff) * dVaux_dVb;
    T11 = 4.97232e-7 * P.weff * P.leff * P.ToxRatio;
    T12 = -7.45669e11 * toxe;
    T3 = P.aigbacc * P.cigbacc - P.bigbacc;
    T4 = P.bigbacc * P.cigbacc;
    T5 = T12 * (P.aigbacc + T3 * Voxacc - T4 * Voxacc * Voxacc);
    if (T5 > kExpThr) { T6 = kMaxExp; dT6_dVg = dT6_dVb = 0.0; }
    else if (T5 < -kExpThr) { T6 = kMinExp; dT6_dVg = dT6_dVb = 0.0; }
    else {
      T6 = exp(T5);
      dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * Voxacc);
      dT6_dVb = dT6_dVg * dVoxacc_dVb;
      dT6_dVg *= dVoxacc_dVg;
    }
    const real dIgbacc_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
    const real dIgbacc_dVb = T11 * (T2 * dT6_dVb + T6 * dT2_dVb);
    const real Igbacc = T11 * T2 * T6;

    T0 = vtm_ig * P.nigbinv;
    T1 = Voxdepinv - P.eigbinv;
    VxNVt = T1 / T0;
    if (VxNVt > kExpThr) {
      Vaux = T1;
      dVaux_dVg = dVoxdepinv_dVg; dVaux_dVd = dVoxdepinv_dVd; dVaux_dVb = dVoxdepinv_dVb;
    } else if (VxNVt < -kExpThr) {
      Vaux = T0 * log(1.0 + kMinExp);
      dVaux_dVg = dVaux_dVd = dVaux_dVb = 0.0;
    } else {
      const real ExpVxNVt = exp(VxNVt);
      Vaux = T0 * log(1.0 + ExpVxNVt);
      dVaux_dVg = ExpVxNVt / (1.0 + ExpVxNVt);
      dVaux_dVd = dVaux_dVg * dVoxdepinv_dVd;
      dVaux_dVb = dVaux_dVg * dVoxdepinv_dVb;
      dVaux_dVg *= dVoxdepinv_dVg;
    }
    T2 = (Vgs_eff - Vbseff) * Vaux;
    dT2_dVg = dVgs_eff_dVg * Vaux + (Vgs_eff - Vbseff) * dVaux_dVg;
    dT2_dVd = (Vgs_eff - Vbseff) * dVaux_dVd;
    dT2_dVb = -Vaux + (Vgs_eff - Vbseff) * dVaux_dVb;
    T11 *= 0.75610;
    T12 *= 1.31724;
    T3 = P.aigbinv * P.cigbinv - P.bigbinv;
    T4 = P.bigbinv * P.cigbinv;
    T5 = T12 * (P.aigbinv + T3 * Voxdepinv - T4 * Voxdepinv * Voxdepinv);
    if (T5 > kExpThr) { T6 = kMaxExp; dT6_dVg = dT6_dVd = dT6_dVb = 0.0; }
    else if (T5 < -kExpThr) { T6 = kMinExp; dT6_dVg = dT6_dVd = dT6_dVb = 0.0; }
    else {
      T6 = exp(T5);
      dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * Voxdepinv);
      dT6_dVd = dT6_dVg * dVoxdepinv_dVd;
      dT6_dVb = dT6_dVg * dVoxdepinv_dVb;
      dT6_dVg *= dVoxdepinv_dVg;
    }
    const real Igbinv = T11 * T2 * T6;
    const real dIgbinv_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
    const real dIgbinv_dVd = T11 * (T2 * dT6_dVd + T6 * dT2_dVd);
    const real dIgbinv_dVb = T11 * (T2 * dT6_dVb + T6 * dT2_dVb);
    W.Igb = Igbinv + Igbacc;
    W.gIgbg = dIgbinv_dVg + dIgbacc_dVg;
    W.gIgbd = dIgbinv_dVd;
    W.gIgbb = (dIgbinv_dVb + dIgbacc_dVb) * dVbseff_dVb;
  } else {
    W.Igb = W.gIgbg = W.gIgbd = W.gIgbs = W.gIgbb = 0.0;
  }

  XB_SYNC_POINT(1);
  // ---- multi-finger scaling ------------------------------------------------------------------------
  if (I.nf != 1.0) {
    const real nf = I.nf;
    W.cdrain *= nf; W.gds *= nf; W.gm *= nf; W.gmbs *= nf; W.IdovVds *= nf;
    W.gbbs *= nf; W.gbgs *= nf; W.gbds *= nf; W.csub *= nf;
    W.Igidl *= nf; W.ggidld *= nf; W.ggidlg *= nf; W.ggidlb *= nf;
    W.Igisl *= nf; W.ggisls *= nf; W.ggislg *= nf; W.ggislb *= nf;
    W.Igcs *= nf; W.gIgcsg *= nf; W.gIgcsd *= nf; W.gIgcsb *= nf;
    W.Igcd *= nf; W.gIgcdg *= nf; W.gIgcdd *= nf; W.gIgcdb *= nf;
    W.Igs *= nf; W.gIgsg *= nf; W.gIgss *= nf;
    W.Igd *= nf; W.gIgdg *= nf; W.gIgdd *= nf;
    W.Igb *= nf; W.gIgbg *= nf; W.gIgbd *= nf; W.gIgbb *= nf;
  }
  W.ggidls = -(W.ggidld + W.ggidlg + W.ggidlb);
  W.ggisld = -(W.ggisls + W.ggislg + W.ggislb);
  W.gIgbs = -(W.gIgbg + W.gIgbd + W.gIgbb);
  W.gIgcss = -(W.gIgcsg + W.gIgcsd + W.gIgcsb);
  W.gIgcds = -(W.gIgcdg + W.gIgcdd + W.gIgcdb);
  W.cd = W.cdrain;

  // The reference recomputes Abulk / Vdsat / Vdseff for the thermal-noise qinv when
  // tnoiMod == 0 (B4p82.C:5787-5822).  They are instance members there, so the values
  // later published to the store vector (Vdsat) are these, not the DC ones.
  if (M.tnoiMod == 0) {
    Abulk = Abulk0 * P.abulkCVfactor;
    Vdsat = Vgsteff / Abulk;
    T0 = Vdsat - Vds - kDelta4;
    T1 = sqrt(T0 * T0 + 4.0 * kDelta4 * Vdsat);
    if (T0 >= 0.0) {
      Vdseff = Vdsat - 0.5 * (T0 + T1);
    } else {
      T3 = (kDelta4 + kDelta4) / (T1 - T0);
      T4 = 1.0 - T3;
      Vdseff = Vdsat * T4;
    }
    if (Vds == 0.0) Vdseff = 0.0;
    W.Vdsat = Vdsat;
    W.Vdseff = Vdseff;
  }

  W.Vds_s = Vds; W.Vgs_s = Vgs; W.Vbs_s = Vbs;   // reference members Vds, Vgs, Vbs

  XB_SYNC_POINT(2);
  // ---- hand over to the C-V stage ---------------------------------------------------------------------
  C.Vds = Vds; C.Vgs = Vgs; C.Vbs = Vbs; C.Vdb = Vdb;
  C.Vbseff = Vbseff; C.dVbseff_dVb = dVbseff_dVb;
  C.Phis = Phis; C.dPhis_dVb = dPhis_dVb; C.sqrtPhis = sqrtPhis; C.dsqrtPhis_dVb = dsqrtPhis_dVb;
  C.Vth = Vth; C.dVth_dVb = dVth_dVb; C.dVth_dVd = dVth_dVd;
  C.Vgs_eff = Vgs_eff; C.dVgs_eff_dVg = dVgs_eff_dVg; C.Vgst = Vgst;
  C.n = n; C.dn_dVb = dn_dVb; C.dn_dVd = dn_dVd; C.Vtm = Vtm; C.Vtm0 = Vtm0;
  C.Vgsteff = Vgsteff; C.dVgsteff_dVg = dVgsteff_dVg; C.dVgsteff_dVd = dVgsteff_dVd; C.dVgsteff_dVb = dVgsteff_dVb;
  C.Vdseff = Vdseff; C.dVdseff_dVg = dVdseff_dVg; C.dVdseff_dVd = dVdseff_dVd; C.dVdseff_dVb = dVdseff_dVb;
  C.Abulk = Abulk; C.dAbulk_dVb = dAbulk_dVb; C.dAbulk_dVg = dAbulk_dVg;
  C.Abulk0 = Abulk0; C.dAbulk0_dVb = dAbulk0_dVb;
  C.Weff = Weff; C.Leff = Leff; C.epsrox = epsrox; C.toxe = toxe; C.epssub = epssub;
  C.Vfb = Vfb; C.dCoxeff_dVg = dCoxeff_dVg;
  C.Vdsat = Vdsat;
  (void)dPhis_dVb; (void)T13; (void)T14;
}

}  // namespace b4
}  // namespace xb
