// xyce_b200 -- BSIM4 load stage: turns the evaluated intermediates (B4Mid) into
// per-instance contributions to the DAE residual vectors F, Q, the limiter
// correction vectors dFdxdVp, dQdxdVp and the Jacobian matrices dF/dx, dQ/dx,
// addressed by *general-stamp* row / slot ids (B4Node / B4Slot).  The caller's
// emitter decides where each general row/slot lands (collapsed 4x4 stamp for the
// default topology, or a per-instance map).
//
// Behavioural specification (reference):
//   Instance::setupCapacitors_oldDAE / _newDAE   N_DEV_MOSFET_B4.C:7412-7898
//   Instance::setupFVectorVars                   N_DEV_MOSFET_B4.C:8101-8429
//   Instance::auxChargeCalculations              N_DEV_MOSFET_B4.C:7290-7410
//   Master::updateState (state/store part)       N_DEV_MOSFET_B4.C:10552-10617
//   Master::loadDAEVectors                       N_DEV_MOSFET_B4.C:10678-10991
//   Master::loadDAEMatrices                      N_DEV_MOSFET_B4.C:11001-11278, :11527-11606
// Not carried over (rejected at the C-ABI): trnqsMod = 1 (the reference itself
// flags it unimplemented, N_DEV_MOSFET_B4.C:11528), IC= branch rows
// (icVBS/icVDS/icVGS) and lead-current/junction-voltage outputs.
#pragma once
#include "bsim4_eval.h"

namespace xb {
namespace b4 {

// Terminal currents through the optional series resistances, final terminal
// charges and the capacitance set of the "new DAE" formulation.
XB_HD void stage_caps(const B4Model &M, const B4Size &P, const B4Inst &I, B4Mid &W) {
  const real cgdo = W.cgdo, cgso = W.cgso, cgbo = P.cgbo;
  // ---- terminal charges (setupCapacitors_oldDAE, trnqsMod == 0 branch) ----
  if (W.mode > 0) {
    W.qdrn -= W.qgdo;
    if (I.rgateMod == 3) {
      const real qgmb = cgbo * W.vgmb;
      W.qgmid = W.qgdo + W.qgso + qgmb;
      W.qbulk -= qgmb;
      W.qsrc = -(W.qgate + W.qgmid + W.qbulk + W.qdrn);
    } else {
      const real qgb = cgbo * W.vgb;
      W.qgate += W.qgdo + W.qgso + qgb;
      W.qbulk -= qgb;
      W.qsrc = -(W.qgate + W.qbulk + W.qdrn);
    }
  } else {
    W.qsrc = W.qdrn - W.qgso;
    if (I.rgateMod == 3) {
      const real qgmb = cgbo * W.vgmb;
      W.qgmid = W.qgdo + W.qgso + qgmb;
      W.qbulk -= qgmb;
      W.qdrn = -(W.qgate + W.qgmid + W.qbulk + W.qsrc);
    } else {
      const real qgb = cgbo * W.vgb;
      W.qgate += W.qgdo + W.qgso + qgb;
      W.qbulk -= qgb;
      W.qdrn = -(W.qgate + W.qbulk + W.qsrc);
    }
  }

  XB_SYNC_POINT(2);
  // ---- capacitances (setupCapacitors_newDAE, trnqsMod == 0 branch) ----
  // Forward-frame intrinsic blocks; reverse mode swaps the drain/source roles.
  const bool fwd = (W.mode > 0);
  const real g_d = fwd ? W.cgdb : W.cgsb;           // gate row, column "drain"
  const real g_s = fwd ? W.cgsb : W.cgdb;
  const real b_d = fwd ? W.cbdb : W.cbsb;
  const real b_s = fwd ? W.cbsb : W.cbdb;
  const real Xg = -(W.cggb + W.cbgb + W.cdgb);      // "source-like" row of the intrinsic block
  const real Xd = -(W.cgdb + W.cbdb + W.cddb);
  if (I.rgateMod == 3) {
    W.CAPcgmgmb = (cgdo + cgso + cgbo);
    W.CAPcgmdb = -cgdo;
    W.CAPcgmsb = -cgso;
    W.CAPcgmbb = -cgbo;
    W.CAPcdgmb = W.CAPcgmdb;
    W.CAPcsgmb = W.CAPcgmsb;
    W.CAPcbgmb = W.CAPcgmbb;
    W.CAPcggb = W.cggb;
    W.CAPcgdb = g_d;
    W.CAPcgsb = g_s;
    W.CAPcgbb = -(W.CAPcggb + W.CAPcgdb + W.CAPcgsb);
    W.CAPcdgb = fwd ? W.cdgb : Xg;
    W.CAPcsgb = fwd ? Xg : W.cdgb;
    W.CAPcbgb = W.cbgb;
  } else {
    W.CAPcgmgmb = W.CAPcgmdb = W.CAPcgmsb = W.CAPcgmbb = 0.0;
    W.CAPcggb = (W.cggb + cgdo + cgso + cgbo);
    W.CAPcgdb = (g_d - cgdo);
    W.CAPcgsb = (g_s - cgso);
    W.CAPcgbb = -(W.CAPcggb + W.CAPcgdb + W.CAPcgsb);
    if (fwd) {
      W.CAPcdgb = (W.cdgb - cgdo);
      W.CAPcsgb = -(W.cggb + W.cbgb + W.cdgb + cgso);
    } else {
      W.CAPcdgb = -(W.cggb + W.cbgb + W.cdgb + cgdo);
      W.CAPcsgb = (W.cdgb - cgso);
    }
    W.CAPcbgb = (W.cbgb - cgbo);
    W.CAPcdgmb = W.CAPcsgmb = W.CAPcbgmb = 0.0;
  }
  if (fwd) {
    W.CAPcddb = (W.cddb + W.capbd + cgdo);
    W.CAPcdsb = W.cdsb;
    W.CAPcsdb = Xd;
    W.CAPcssb = (W.capbs + cgso - (W.cgsb + W.cbsb + W.cdsb));
  } else {
    W.CAPcddb = (W.capbd + cgdo - (W.cgsb + W.cbsb + W.cdsb));
    W.CAPcdsb = Xd;
    W.CAPcsdb = W.cdsb;
    W.CAPcssb = (W.cddb + W.capbs + cgso);
  }
  if (!I.rbodyMod) {
    W.CAPcdbb = -(W.CAPcdgb + W.CAPcddb + W.CAPcdsb + W.CAPcdgmb);
    W.CAPcsbb = -(W.CAPcsgb + W.CAPcsdb + W.CAPcssb + W.CAPcsgmb);
    W.CAPcbdb = (b_d - W.capbd);
    W.CAPcbsb = (b_s - W.capbs);
    W.CAPcdbdb = 0.0;
    W.CAPcsbsb = 0.0;
  } else {
    if (fwd) {
      W.CAPcdbb = -(W.cddb + W.cdgb + W.cdsb);
      W.CAPcsbb = -(W.CAPcsgb + W.CAPcsdb + W.CAPcssb + W.CAPcsgmb) + W.capbs;
    } else {
      W.CAPcdbb = -(W.CAPcdgb + W.CAPcddb + W.CAPcdsb + W.CAPcdgmb) + W.capbd;
      W.CAPcsbb = -(W.cddb + W.cdgb + W.cdsb);
    }
    W.CAPcbdb = b_d;
    W.CAPcbsb = b_s;
    W.CAPcdbdb = -W.capbd;
    W.CAPcsbsb = -W.capbs;
  }
  W.CAPcbbb = -(W.CAPcbdb + W.CAPcbgb + W.CAPcbsb + W.CAPcbgmb);

  XB_SYNC_POINT(2);
  // ---- series-resistance currents (B4p82.C:7011-7032) ----
  if (M.rdsMod == 1) {
    W.Idrain = W.gdtot * W.Vddp;
    W.Isource = W.gstot * W.Vssp;
  } else {
    W.Idrain = I.drainConductance * W.Vddp;
    W.Isource = I.sourceConductance * W.Vssp;
  }
  if (M.rbodyMod != 0) {
    W.Idbb = I.grbdb * W.Vdbb;
    W.Idbbp = I.grbpd * W.Vdbbp;
    W.Isbb = I.grbsb * W.Vsbb;
    W.Isbbp = I.grbps * W.Vsbbp;
    W.Ibpb = I.grbpb * W.Vbpb;
  } else {
    W.Idbb = W.Idbbp = W.Isbb = W.Isbbp = W.Ibpb = 0.0;
  }

  // ---- state-vector charges (Master::updateState) ----
  W.qg = W.qgate;
  W.qd = W.qdrn - W.qbd;
  W.qb = (!I.rbodyMod) ? W.qbulk + W.qbd + W.qbs : W.qbulk;
}

// Equivalent currents, limiter (Jdxp) terms and the mode-dependent conductance
// bookkeeping of setupFVectorVars.
XB_HD void stage_fvars(const B4Model &M, const B4Inst &I, B4Mid &W) {
  const real ty = real(M.dtype);
  const real dvds = W.vds - W.vds_orig, dvgs = W.vgs - W.vgs_orig, dvbs = W.vbs - W.vbs_orig;
  const real dvgd = W.vgd - W.vgd_orig, dvbd = W.vbd - W.vbd_orig;
  real T0 = 0.0;
  W.ceqgcrg_Jdxp = 0.0;
  if (W.mode >= 0) {
    W.Gm = W.gm;
    W.Gmbs = W.gmbs;
    W.FwdSum = W.Gm + W.Gmbs;
    W.RevSum = 0.0;
    W.ceqdrn = ty * W.cdrain;
    W.ceqdrn_Jdxp = ty * (-W.gds * dvds - W.Gm * dvgs - W.Gmbs * dvbs);
    W.ceqbd = ty * (W.csub + W.Igidl);
    W.ceqbd_Jdxp = ty * (-(W.gbds + W.ggidld) * dvds - (W.gbgs + W.ggidlg) * dvgs - (W.gbbs + W.ggidlb) * dvbs);
    W.ceqbs = ty * W.Igisl;
    W.ceqbs_Jdxp = ty * (+W.ggisls * dvds - W.ggislg * dvgd - W.ggislb * dvbd);
    W.gbbdp = -(W.gbds);
    W.gbbsp = W.gbds + W.gbgs + W.gbbs;
    W.gbdpg = W.gbgs;
    W.gbdpdp = W.gbds;
    W.gbdpb = W.gbbs;
    W.gbdpsp = -(W.gbdpg + W.gbdpdp + W.gbdpb);
    W.gbspg = 0.0; W.gbspdp = 0.0; W.gbspb = 0.0; W.gbspsp = 0.0;
    if (M.igcMod) {
      W.gIstotg = W.gIgsg + W.gIgcsg;
      W.gIstotd = W.gIgcsd;
      W.gIstots = W.gIgss + W.gIgcss;
      W.gIstotb = W.gIgcsb;
      W.Istoteq = ty * (W.Igs + W.Igcs);
      W.Istoteq_Jdxp = ty * (-W.gIstotg * dvgs - W.gIgcsd * dvds - W.gIgcsb * dvbs);
      W.gIdtotg = W.gIgdg + W.gIgcdg;
      W.gIdtotd = W.gIgdd + W.gIgcdd;
      W.gIdtots = W.gIgcds;
      W.gIdtotb = W.gIgcdb;
      W.Idtoteq = ty * (W.Igd + W.Igcd);
      W.Idtoteq_Jdxp = ty * (-W.gIgdg * dvgd - W.gIgcdg * dvgs - W.gIgcdd * dvds - W.gIgcdb * dvbs);
    } else {
      W.gIstotg = W.gIstotd = W.gIstots = W.gIstotb = W.Istoteq = 0.0;
      W.gIdtotg = W.gIdtotd = W.gIdtots = W.gIdtotb = W.Idtoteq = 0.0;
      W.Istoteq_Jdxp = 0.0;
      W.Idtoteq_Jdxp = 0.0;
    }
    if (M.igbMod) {
      W.gIbtotg = W.gIgbg; W.gIbtotd = W.gIgbd; W.gIbtots = W.gIgbs; W.gIbtotb = W.gIgbb;
      W.Ibtoteq = ty * W.Igb;
      W.Ibtoteq_Jdxp = ty * (-W.gIgbg * dvgs - W.gIgbd * dvds - W.gIgbb * dvbs);
    } else {
      W.gIbtotg = W.gIbtotd = W.gIbtots = W.gIbtotb = W.Ibtoteq = 0.0;
      W.Ibtoteq_Jdxp = 0.0;
    }
  } else {
    W.Gm = -W.gm;
    W.Gmbs = -W.gmbs;
    W.FwdSum = 0.0;
    W.RevSum = -(W.Gm + W.Gmbs);
    W.ceqdrn = -ty * W.cdrain;
    W.ceqdrn_Jdxp = -ty * (+W.gds * dvds + W.Gm * dvgd + W.Gmbs * dvbd);
    W.ceqbs = ty * (W.csub + W.Igisl);
    W.ceqbs_Jdxp = ty * (+(W.gbds + W.ggisls) * dvds - (W.gbgs + W.ggislg) * dvgd - (W.gbbs + W.ggislb) * dvbd);
    W.ceqbd = ty * W.Igidl;
    W.ceqbd_Jdxp = ty * (-W.ggidld * dvds - W.ggidlg * dvgs - W.ggidlb * dvbs);
    W.gbbsp = -(W.gbds);
    W.gbbdp = W.gbds + W.gbgs + W.gbbs;
    W.gbdpg = 0.0; W.gbdpsp = 0.0; W.gbdpb = 0.0; W.gbdpdp = 0.0;
    W.gbspg = W.gbgs;
    W.gbspsp = W.gbds;
    W.gbspb = W.gbbs;
    W.gbspdp = -(W.gbspg + W.gbspsp + W.gbspb);
    if (M.igcMod) {
      W.gIstotg = W.gIgsg + W.gIgcdg;
      W.gIstotd = W.gIgcds;
      W.gIstots = W.gIgss + W.gIgcdd;
      W.gIstotb = W.gIgcdb;
      W.Istoteq = ty * (W.Igs + W.Igcd);
      W.Istoteq_Jdxp = ty * (-W.gIgsg * dvgs - W.gIgcdg * dvgd + W.gIgcdd * dvds - W.gIgcdb * dvbd);
      W.gIdtotg = W.gIgdg + W.gIgcsg;
      W.gIdtotd = W.gIgdd + W.gIgcss;
      W.gIdtots = W.gIgcsd;
      W.gIdtotb = W.gIgcsb;
      W.Idtoteq = ty * (W.Igd + W.Igcs);
      W.Idtoteq_Jdxp = ty * (-(W.gIgdg + W.gIgcsg) * dvgd + W.gIgcsd * dvds - W.gIgcsb * dvbd);
    } else {
      W.gIstotg = W.gIstotd = W.gIstots = W.gIstotb = W.Istoteq = 0.0;
      W.gIdtotg = W.gIdtotd = W.gIdtots = W.gIdtotb = W.Idtoteq = 0.0;
      W.Istoteq_Jdxp = 0.0;
      W.Idtoteq_Jdxp = 0.0;
    }
    if (M.igbMod) {
      W.gIbtotg = W.gIgbg; W.gIbtotd = W.gIgbs; W.gIbtots = W.gIgbd; W.gIbtotb = W.gIgbb;
      W.Ibtoteq = ty * W.Igb;
      W.Ibtoteq_Jdxp = ty * (-W.gIgbg * dvgd + W.gIgbd * dvds - W.gIgbb * dvbd);
    } else {
      W.gIbtotg = W.gIbtotd = W.gIbtots = W.gIbtotb = W.Ibtoteq = 0.0;
      W.Ibtoteq_Jdxp = 0.0;
    }
  }
  if ((M.igcMod != 0) || (M.igbMod != 0)) {
    W.gIgtotg = W.gIstotg + W.gIdtotg + W.gIbtotg;
    W.gIgtotd = W.gIstotd + W.gIdtotd + W.gIbtotd;
    W.gIgtots = W.gIstots + W.gIdtots + W.gIbtots;
    W.gIgtotb = W.gIstotb + W.gIdtotb + W.gIbtotb;
    W.Igtoteq = W.Istoteq + W.Idtoteq + W.Ibtoteq;
    W.Igtoteq_Jdxp = W.Istoteq_Jdxp + W.Idtoteq_Jdxp + W.Ibtoteq_Jdxp;
  } else {
    W.gIgtotg = W.gIgtotd = W.gIgtots = W.gIgtotb = W.Igtoteq = 0.0;
    W.Igtoteq_Jdxp = 0.0;
  }
  if (I.rgateMod == 2) T0 = W.vges - W.vgs;
  else if (I.rgateMod == 3) T0 = W.vgms - W.vgs;
  if (I.rgateMod > 1) {
    if (W.mode >= 0) {
      W.gcrgd = W.gcrgd * T0;
      W.gcrgg = W.gcrgg * T0;
      W.gcrgs = W.gcrgs * T0;
      W.gcrgb = W.gcrgb * T0;
      W.ceqgcrg = 0.0;
      W.ceqgcrg_Jdxp = -(W.gcrgd * dvds + W.gcrgg * dvgs + W.gcrgb * dvbs);
    } else {
      const real tmp_gcrgd = W.gcrgd;
      W.gcrgd = W.gcrgs * T0;
      W.gcrgg = W.gcrgg * T0;
      W.gcrgs = tmp_gcrgd * T0;
      W.gcrgb = W.gcrgb * T0;
      W.ceqgcrg = 0.0;
      W.ceqgcrg_Jdxp = -(W.gcrgg * dvgd - W.gcrgs * dvds + W.gcrgb * dvbd);
    }
    W.gcrgg -= W.gcrg;
  } else {
    W.ceqgcrg = W.gcrg = W.gcrgd = W.gcrgg = W.gcrgs = W.gcrgb = 0.0;
  }

  if (M.rdsMod == 1) {
    W.ceqgstot = 0.0;
    W.ceqgstot_Jdxp = ty * (W.gstotd * dvds + W.gstotg * dvgs + W.gstotb * dvbs);
    W.gstots = W.gstots - W.gstot;
    W.ceqgdtot = 0.0;
    W.ceqgdtot_Jdxp = -ty * (W.gdtotd * dvds + W.gdtotg * dvgs + W.gdtotb * dvbs);
    W.gdtotd = W.gdtotd - W.gdtot;
  } else {
    W.gstot = W.gstotd = W.gstotg = W.gstots = W.gstotb = W.ceqgstot = 0.0;
    W.gdtot = W.gdtotd = W.gdtotg = W.gdtots = W.gdtotb = W.ceqgdtot = 0.0;
    W.ceqgstot_Jdxp = 0.0;
    W.ceqgdtot_Jdxp = 0.0;
  }

  if (M.dtype > 0) {
    W.ceqjs = W.cbs;
    W.ceqjs_Jdxp = (-W.gbs * (W.vbs_jct - W.vbs_jct_orig));
    W.ceqjd = W.cbd;
    W.ceqjd_Jdxp = (-W.gbd * (W.vbd_jct - W.vbd_jct_orig));
  } else {
    W.ceqjs = -W.cbs;
    W.ceqjs_Jdxp = (W.gbs * (W.vbs_jct - W.vbs_jct_orig));
    W.ceqjd = -W.cbd;
    W.ceqjd_Jdxp = (W.gbd * (W.vbd_jct - W.vbd_jct_orig));
    W.ceqgcrg = -W.ceqgcrg;
    W.ceqgcrg_Jdxp = -W.ceqgcrg_Jdxp;
  }

  // limiter terms of the charge equations (auxChargeCalculations, trnqsMod == 0)
  W.Qeqqg_Jdxp = W.Qeqqd_Jdxp = W.Qeqqb_Jdxp = 0.0;
  W.Qeqqgmid_Jdxp = W.Qeqqjs_Jdxp = W.Qeqqjd_Jdxp = 0.0;
  if (W.ChargeComputationNeeded && !W.origFlag) {
    const real dvgb = W.vgb - W.vgb_orig, dvgmb = W.vgmb - W.vgmb_orig;
    W.Qeqqg_Jdxp = -W.CAPcggb * dvgb + W.CAPcgdb * dvbd + W.CAPcgsb * dvbs;
    W.Qeqqd_Jdxp = -W.CAPcdgb * dvgb - W.CAPcdgmb * dvgmb + (W.CAPcddb + W.CAPcdbdb) * dvbd
                   - W.CAPcdbdb * (W.vbd_jct - W.vbd_jct_orig) + W.CAPcdsb * dvbs;
    W.Qeqqb_Jdxp = -W.CAPcbgb * dvgb - W.CAPcbgmb * dvgmb + W.CAPcbdb * dvbd + W.CAPcbsb * dvbs;
    if (I.rgateMod == 3)
      W.Qeqqgmid_Jdxp = +W.CAPcgmdb * dvbd + W.CAPcgmsb * dvbs - W.CAPcgmgmb * dvgmb;
    if (I.rbodyMod) {
      W.Qeqqjs_Jdxp = W.CAPcsbsb * (W.vbs_jct - W.vbs_jct_orig);
      W.Qeqqjd_Jdxp = W.CAPcdbdb * (W.vbd_jct - W.vbd_jct_orig);
    }
  }
}

// ---------------------------------------------------------------------------
// Emission.  `E` provides:
//   template<int ROW>  void f(real), q(real), fl(real), ql(real)   (+=)
//   template<int SLOT> void jf(real), jq(real)                          (+=)
// ---------------------------------------------------------------------------
template <class E>
XB_HD void emit_vectors(const SolverFlags &S, const B4Model &M, const B4Inst &I, const B4Mid &W, E &e) {
  const real np = I.numberParallel;
  e.template f<kDP>(-(W.ceqjd - W.ceqbd - W.ceqdrn + W.Idtoteq) * np);
  e.template f<kGP>(-(-(-W.ceqgcrg + W.Igtoteq) * np));
  if (I.rgateMod == 1) {
    e.template f<kGE>((W.Igate) * np);
    e.template f<kGP>(-((W.Igate) * np));
  } else if (I.rgateMod == 2) {
    e.template f<kGE>((W.Igate + W.ceqgcrg) * np);
    e.template f<kGP>(-(+(W.Igate) * np));
  } else if (I.rgateMod == 3) {
    e.template f<kGE>((W.Igate) * np);
    e.template f<kGM>((W.IgateMid - W.Igate + W.ceqgcrg) * np);
    e.template f<kGP>(-(-(W.IgateMid) * np));
  }
  if (!I.rbodyMod) {
    e.template f<kBP>(-(W.ceqbd + W.ceqbs - W.ceqjd - W.ceqjs + W.Ibtoteq) * np);
    e.template f<kSP>(-(W.ceqdrn - W.ceqbs + W.ceqjs + W.Istoteq) * np);
  } else {
    e.template f<kDB>(-(-(W.ceqjd + W.Idbb + W.Idbbp) * np));
    e.template f<kBP>(-(W.ceqbd + W.ceqbs + W.Ibtoteq + W.Idbbp + W.Isbbp - W.Ibpb) * np);
    e.template f<kB>(-(W.Isbb + W.Idbb + W.Ibpb) * np);
    e.template f<kSB>(-(-(W.ceqjs + W.Isbb + W.Isbbp) * np));
    e.template f<kSP>(-(W.ceqdrn - W.ceqbs + W.ceqjs + W.Istoteq) * np);
  }
  if (M.rdsMod) {
    e.template f<kD>(-(-W.ceqgdtot) * np);
    e.template f<kS>(-(W.ceqgstot) * np);
    e.template f<kDP>(-(W.ceqgdtot) * np);
    e.template f<kSP>(-(-W.ceqgstot) * np);
  }
  if (I.drainMOSFET_B4Exists) {
    e.template f<kD>(-(-W.Idrain) * np);
    e.template f<kDP>(-(W.Idrain) * np);
  }
  if (I.sourceMOSFET_B4Exists) {
    e.template f<kS>(-(-W.Isource) * np);
    e.template f<kSP>(-(+W.Isource) * np);
  }

  const bool lim = S.voltageLimiterFlag && !W.origFlag;
  if (lim) {
    e.template fl<kDP>((W.ceqjd_Jdxp - W.ceqbd_Jdxp - W.ceqdrn_Jdxp + W.Idtoteq_Jdxp) * np);
    e.template fl<kGP>(-((-W.ceqgcrg_Jdxp + W.Igtoteq_Jdxp) * np));
    if (I.rgateMod == 2) e.template fl<kGE>((-W.ceqgcrg_Jdxp) * np);
    else if (I.rgateMod == 3) e.template fl<kGM>((-W.ceqgcrg_Jdxp) * np);
    if (!I.rbodyMod) {
      e.template fl<kBP>((W.ceqbd_Jdxp + W.ceqbs_Jdxp - W.ceqjd_Jdxp - W.ceqjs_Jdxp + W.Ibtoteq_Jdxp) * np);
      e.template fl<kSP>((W.ceqdrn_Jdxp - W.ceqbs_Jdxp + W.ceqjs_Jdxp + W.Istoteq_Jdxp) * np);
    } else {
      e.template fl<kDB>(-((W.ceqjd_Jdxp) * np));
      e.template fl<kBP>((W.ceqbd_Jdxp + W.ceqbs_Jdxp + W.Ibtoteq_Jdxp) * np);
      e.template fl<kSB>(-((W.ceqjs_Jdxp) * np));
      e.template fl<kSP>((W.ceqdrn_Jdxp - W.ceqbs_Jdxp + W.ceqjs_Jdxp + W.Istoteq_Jdxp) * np);
    }
    if (M.rdsMod) {
      e.template fl<kD>(-((W.ceqgdtot_Jdxp) * np));
      e.template fl<kS>((W.ceqgstot_Jdxp) * np);
      e.template fl<kDP>((W.ceqgdtot_Jdxp) * np);
      e.template fl<kSP>(-((W.ceqgstot_Jdxp) * np));
    }
  }

  // charge rows
  const real sg = (M.dtype > 0) ? 1.0 : -1.0;
  const real Qg = sg * W.qg, Qd = sg * W.qd, Qb = sg * W.qb;
  const real Qjs = I.rbodyMod ? sg * W.qbs : 0.0;
  const real Qjd = I.rbodyMod ? sg * W.qbd : 0.0;
  const real Qgmid = (I.rgateMod == 3) ? sg * W.qgmid : 0.0;
  e.template q<kDP>(-(-Qd) * np);
  e.template q<kGP>(-(-(Qg) * np));
  if (I.rgateMod == 3) e.template q<kGM>(-(-(+Qgmid) * np));
  if (!I.rbodyMod) {
    e.template q<kBP>(-(-Qb) * np);
    e.template q<kSP>(-(+Qg + Qb + Qd + Qgmid) * np);
  } else {
    e.template q<kDB>(-(-(Qjd) * np));
    e.template q<kBP>(-(-Qb) * np);
    e.template q<kSB>(-(-(Qjs) * np));
    e.template q<kSP>(-(Qd + Qg + Qb + Qjd + Qjs + Qgmid) * np);
  }
  if (lim) {
    e.template ql<kDP>((-W.Qeqqd_Jdxp) * np);
    e.template ql<kGP>(-((W.Qeqqg_Jdxp) * np));
    if (I.rgateMod == 3) e.template ql<kGM>(-((+W.Qeqqgmid_Jdxp) * np));
    if (!I.rbodyMod) {
      e.template ql<kBP>((-W.Qeqqb_Jdxp) * np);
      e.template ql<kSP>((+W.Qeqqg_Jdxp + W.Qeqqb_Jdxp + W.Qeqqd_Jdxp + W.Qeqqgmid_Jdxp) * np);
    } else {
      e.template ql<kDB>(-((W.Qeqqjd_Jdxp) * np));
      e.template ql<kBP>((-W.Qeqqb_Jdxp) * np);
      e.template ql<kSB>(-((+W.Qeqqjs_Jdxp) * np));
      e.template ql<kSP>((+W.Qeqqd_Jdxp + W.Qeqqg_Jdxp + W.Qeqqb_Jdxp + W.Qeqqjd_Jdxp + W.Qeqqjs_Jdxp + W.Qeqqgmid_Jdxp) * np);
    }
  }
}

// Lead currents and charges of the four terminals (id, ig, is, ib): Master::loadDAEVectors' loadLeadCurrent block
// (N_DEV_MOSFET_B4.C:10933-10987), for any topology.
XB_HD void emit_lead(const B4Model &M, const B4Inst &I, const B4Mid &W, real (&leadF)[4], real (&leadQ)[4]) {
  const real np = I.numberParallel;
  const real sg = (M.dtype > 0) ? 1.0 : -1.0;
  const real Qg = sg * W.qg, Qd = sg * W.qd, Qb = sg * W.qb;
  const real Qjs = I.rbodyMod ? sg * W.qbs : 0.0;
  const real Qjd = I.rbodyMod ? sg * W.qbd : 0.0;
  const real Qgmid = (I.rgateMod == 3) ? sg * W.qgmid : 0.0;
  leadQ[0] = (Qd) * np;
  leadQ[1] = (Qg) * np;
  leadQ[3] = (Qb) * np;
  if (!I.rbodyMod) leadQ[2] = -(+Qg + Qb + Qd + Qgmid) * np;
  else leadQ[2] = -(Qd + Qg + Qb + Qjd + Qjs + Qgmid) * np;
  leadF[0] = -(W.ceqjd - W.ceqbd - W.ceqdrn + W.Idtoteq) * np;
  leadF[2] = W.Isource * np;
  leadF[1] = (-W.ceqgcrg + W.Igtoteq) * np;
  leadF[3] = 0.0;
  if (I.rgateMod == 1) leadF[1] += (W.Igate) * np;
  else if (I.rgateMod == 2) leadF[1] += (W.Igate) * np;
  else if (I.rgateMod == 3) leadF[1] += (W.IgateMid) * np;
  if (!I.rbodyMod) {
    leadF[3] += -(W.ceqbd + W.ceqbs - W.ceqjd - W.ceqjs + W.Ibtoteq) * np;
    leadF[2] += -(W.ceqdrn - W.ceqbs + W.ceqjs + W.Istoteq) * np;
  } else {
    leadF[3] = -(W.Isbb + W.Idbb + W.Ibpb) * np;
    leadF[2] += -(W.ceqdrn - W.ceqbs + W.ceqjs + W.Istoteq) * np;
  }
  if (M.rdsMod) {
    leadF[0] += (W.ceqgdtot) * np;
    leadF[2] += -(W.ceqgstot) * np;
  }
}

template <class E>
XB_HD void emit_matrices(const B4Model &M, const B4Inst &I, const B4Mid &W, E &e) {
  const real np = I.numberParallel;
  const real gjbd = (!I.rbodyMod) ? W.gbd : 0.0;
  const real gjbs = (!I.rbodyMod) ? W.gbs : 0.0;
  const real gdpr = (!M.rdsMod) ? I.drainConductance : 0.0;
  const real gspr = (!M.rdsMod) ? I.sourceConductance : 0.0;
  const real geltd = I.grgeltd;
  // trnqsMod == 0: ggt* = 0, T1 = qdef*gtau multiplies ddxpart = 0, dxpart/sxpart irrelevant.
  if (I.rgateMod == 1) {
    e.template jf<sGEge>((geltd) * np);
    e.template jf<sGEgp>(-((geltd) * np));
    e.template jf<sGPge>(-((geltd) * np));
    e.template jf<sGPgp>((+geltd + W.gIgtotg) * np);
    e.template jf<sGPdp>((W.gIgtotd) * np);
    e.template jf<sGPsp>((W.gIgtots) * np);
    e.template jf<sGPbp>((W.gIgtotb) * np);
  } else if (I.rgateMod == 2) {
    e.template jf<sGEge>((W.gcrg) * np);
    e.template jf<sGEgp>((W.gcrgg) * np);
    e.template jf<sGEdp>((W.gcrgd) * np);
    e.template jf<sGEsp>((W.gcrgs) * np);
    e.template jf<sGEbp>((W.gcrgb) * np);
    e.template jf<sGPge>(-((W.gcrg) * np));
    e.template jf<sGPgp>((-W.gcrgg + W.gIgtotg) * np);
    e.template jf<sGPdp>((-W.gcrgd + W.gIgtotd) * np);
    e.template jf<sGPsp>((-W.gcrgs + W.gIgtots) * np);
    e.template jf<sGPbp>((-W.gcrgb + W.gIgtotb) * np);
  } else if (I.rgateMod == 3) {
    e.template jf<sGEge>((geltd) * np);
    e.template jf<sGEgm>(-((geltd) * np));
    e.template jf<sGMge>(-((geltd) * np));
    e.template jf<sGMgm>((geltd + W.gcrg) * np);
    e.template jf<sGMdp>((W.gcrgd) * np);
    e.template jf<sGMgp>((W.gcrgg) * np);
    e.template jf<sGMsp>((W.gcrgs) * np);
    e.template jf<sGMbp>((W.gcrgb) * np);
    e.template jf<sGPgm>(-((W.gcrg) * np));
    e.template jf<sGPgp>((-W.gcrgg + W.gIgtotg) * np);
    e.template jf<sGPdp>((-W.gcrgd + W.gIgtotd) * np);
    e.template jf<sGPsp>((-W.gcrgs + W.gIgtots) * np);
    e.template jf<sGPbp>((-W.gcrgb + W.gIgtotb) * np);
  } else {
    e.template jf<sGPgp>((W.gIgtotg) * np);
    e.template jf<sGPdp>((W.gIgtotd) * np);
    e.template jf<sGPsp>((W.gIgtots) * np);
    e.template jf<sGPbp>((W.gIgtotb) * np);
  }
  if (M.rdsMod) {
    e.template jf<sDgp>((W.gdtotg) * np);
    e.template jf<sDsp>((W.gdtots) * np);
    e.template jf<sDbp>((W.gdtotb) * np);
    e.template jf<sSdp>((W.gstotd) * np);
    e.template jf<sSgp>((W.gstotg) * np);
    e.template jf<sSbp>((W.gstotb) * np);
  }
  e.template jf<sDPdp>((gdpr + W.gds + W.gbd - W.gdtotd + W.RevSum + W.gbdpdp - W.gIdtotd) * np);
  e.template jf<sDPd>(-((gdpr + W.gdtot) * np));
  e.template jf<sDPgp>((W.Gm - W.gdtotg + W.gbdpg - W.gIdtotg) * np);
  e.template jf<sDPsp>(-((W.gds + W.gdtots + W.gIdtots + W.FwdSum - W.gbdpsp) * np));
  e.template jf<sDPbp>(-((gjbd + W.gdtotb - W.Gmbs - W.gbdpb + W.gIdtotb) * np));
  e.template jf<sDdp>(-((gdpr - W.gdtotd) * np));
  e.template jf<sDd>((gdpr + W.gdtot) * np);
  e.template jf<sSPdp>(-((W.gds + W.gstotd + W.RevSum - W.gbspdp + W.gIstotd) * np));
  e.template jf<sSPgp>((-W.Gm - W.gstotg + W.gbspg - W.gIstotg) * np);
  e.template jf<sSPsp>((gspr + W.gds + W.gbs - W.gstots + W.FwdSum + W.gbspsp - W.gIstots) * np);
  e.template jf<sSPs>(-((gspr + W.gstot) * np));
  e.template jf<sSPbp>(-((gjbs + W.gstotb + W.Gmbs - W.gbspb + W.gIstotb) * np));
  e.template jf<sSsp>(-((gspr - W.gstots) * np));
  e.template jf<sSs>((gspr + W.gstot) * np);
  e.template jf<sBPdp>((-gjbd + W.gbbdp - W.gIbtotd) * np);
  e.template jf<sBPgp>((-W.gbgs - W.gIbtotg) * np);
  e.template jf<sBPsp>((-gjbs + W.gbbsp - W.gIbtots) * np);
  e.template jf<sBPbp>((gjbd + gjbs - W.gbbs - W.gIbtotb) * np);
  // GIDL / GISL
  e.template jf<sDPdp>((W.ggidld) * np);
  e.template jf<sDPgp>((W.ggidlg) * np);
  e.template jf<sDPsp>(-(((W.ggidlg + W.ggidld + W.ggidlb)) * np));
  e.template jf<sDPbp>((W.ggidlb) * np);
  e.template jf<sBPdp>(-((W.ggidld) * np));
  e.template jf<sBPgp>(-((W.ggidlg) * np));
  e.template jf<sBPsp>(((W.ggidlg + W.ggidld + W.ggidlb)) * np);
  e.template jf<sBPbp>(-((W.ggidlb) * np));
  e.template jf<sSPdp>(-(((W.ggisls + W.ggislg + W.ggislb)) * np));
  e.template jf<sSPgp>((W.ggislg) * np);
  e.template jf<sSPsp>((W.ggisls) * np);
  e.template jf<sSPbp>((W.ggislb) * np);
  e.template jf<sBPdp>(((W.ggislg + W.ggisls + W.ggislb)) * np);
  e.template jf<sBPgp>(-((W.ggislg) * np));
  e.template jf<sBPsp>(-((W.ggisls) * np));
  e.template jf<sBPbp>(-((W.ggislb) * np));
  if (I.rbodyMod) {
    e.template jf<sDPdb>((-W.gbd) * np);
    e.template jf<sSPsb>(-((W.gbs) * np));
    e.template jf<sDBdp>((-W.gbd) * np);
    e.template jf<sDBdb>((W.gbd + I.grbpd + I.grbdb) * np);
    e.template jf<sDBbp>(-((I.grbpd) * np));
    e.template jf<sDBb>(-((I.grbdb) * np));
    e.template jf<sBPdb>(-((I.grbpd) * np));
    e.template jf<sBPb>(-((I.grbpb) * np));
    e.template jf<sBPsb>(-((I.grbps) * np));
    e.template jf<sBPbp>((I.grbpd + I.grbps + I.grbpb) * np);
    e.template jf<sSBsp>((-W.gbs) * np);
    e.template jf<sSBbp>(-((I.grbps) * np));
    e.template jf<sSBb>(-((I.grbsb) * np));
    e.template jf<sSBsb>((W.gbs + I.grbps + I.grbsb) * np);
    e.template jf<sBdb>(-((I.grbdb) * np));
    e.template jf<sBbp>(-((I.grbpb) * np));
    e.template jf<sBsb>(-((I.grbsb) * np));
    e.template jf<sBb>((I.grbsb + I.grbdb + I.grbpb) * np);
  }

  // dQ/dx (loaded unconditionally, like the reference)
  {
    if (I.rgateMod == 3) {
      e.template jq<sGMgm>((+W.CAPcgmgmb) * np);
      e.template jq<sGMdp>((W.CAPcgmdb) * np);
      e.template jq<sGMsp>((W.CAPcgmsb) * np);
      e.template jq<sGMbp>((W.CAPcgmbb) * np);
      e.template jq<sDPgm>((W.CAPcdgmb) * np);
      e.template jq<sSPgm>((W.CAPcsgmb) * np);
      e.template jq<sBPgm>((W.CAPcbgmb) * np);
    }
    e.template jq<sGPgp>((W.CAPcggb) * np);
    e.template jq<sGPdp>((W.CAPcgdb) * np);
    e.template jq<sGPsp>((W.CAPcgsb) * np);
    e.template jq<sGPbp>((W.CAPcgbb) * np);
    e.template jq<sDPdp>((W.CAPcddb) * np);
    e.template jq<sDPgp>((+W.CAPcdgb) * np);
    e.template jq<sDPsp>(-((-W.CAPcdsb) * np));
    e.template jq<sDPbp>(-((-W.CAPcdbb) * np));
    e.template jq<sSPdp>(-((-W.CAPcsdb) * np));
    e.template jq<sSPgp>((W.CAPcsgb) * np);
    e.template jq<sSPsp>((W.CAPcssb) * np);
    e.template jq<sSPbp>(-((-W.CAPcsbb) * np));
    e.template jq<sBPdp>((W.CAPcbdb) * np);
    e.template jq<sBPgp>((W.CAPcbgb) * np);
    e.template jq<sBPsp>((W.CAPcbsb) * np);
    e.template jq<sBPbp>((W.CAPcbbb) * np);
    if (I.rbodyMod) {
      e.template jq<sDPdb>((W.CAPcdbdb) * np);
      e.template jq<sSPsb>(-((-W.CAPcsbsb) * np));
      e.template jq<sDBdp>((W.CAPcdbdb) * np);
      e.template jq<sDBdb>((-W.CAPcdbdb) * np);
      e.template jq<sSBsp>((W.CAPcsbsb) * np);
      e.template jq<sSBsb>((-W.CAPcsbsb) * np);
    }
  }
}

}  // namespace b4
}  // namespace xb
