// xyce_b200 -- BSIM4 (level 14/54, version 4.8.2) bias-dependent evaluation.
//
// One call evaluates one MOSFET instance at one Newton iterate: terminal
// voltage limiting, S/D junction diodes, threshold voltage, mobility, Vdsat,
// drain current and its derivatives, substrate current, GIDL/GISL, gate
// tunnelling, bias-dependent S/D resistance, intrinsic C-V (capMod 0/1/2),
// junction and overlap capacitances.  The result is the `B4Mid` record that the
// load stage (bsim4_load.h) turns into f/q rows and dF/dx, dQ/dx stamps.
//
// Behavioural specification: the reference's
//   Instance::updateIntermediateVars4p82_  (src/DeviceModelPKG/OpenModels/N_DEV_MOSFET_B4p82.C:2744-7035)
// with helper Instance::polyDepletion       (src/DeviceModelPKG/OpenModels/N_DEV_MOSFET_B4.C:8819-8857).
// This is a from-scratch restatement organised for a one-thread-per-instance
// SoA GPU kernel: model card (`M`), size-dependent bin (`P`) and instance
// constants (`I`) arrive as read-only records, every intermediate is a local,
// and nothing is cached on an instance object.  Compiled by nvcc for sm_100a
// (product) and by g++ only inside tests/host_mirror.
#pragma once
#include "xb_common.h"
#include "bsim4_types.h"

namespace xb {
namespace b4 {

// Poly-gate depletion (N_DEV_MOSFET_B4.C:8819-8857).
XB_HD void poly_depletion(real phi, real ngate, real epsgate, real coxe,
                          real vg, real &vg_eff, real &dvg_eff_dvg) {
  if ((ngate > 1.0e18) && (ngate < 1.0e25) && (vg > phi) && (epsgate != 0)) {
    const real t1 = 1.0e6 * kQ * epsgate * ngate / (coxe * coxe);
    const real t8 = vg - phi;
    const real t4 = sqrt(1.0 + 2.0 * t8 / t1);
    const real t2 = 2.0 * t8 / (t4 + 1.0);
    const real t3 = 0.5 * t2 * t2 / t1;
    const real t7 = 1.12 - t3 - 0.05;
    const real t6 = sqrt(t7 * t7 + 0.224);
    const real t5 = 1.12 - 0.5 * (t7 + t6);
    vg_eff = vg - t5;
    dvg_eff_dvg = 1.0 - (0.5 - 0.5 / t4) * (1.0 + t7 / t6);
  } else {
    vg_eff = vg;
    dvg_eff_dvg = 1.0;
  }
}

// One bulk junction diode (source or drain side) -- B4p82.C:3348-3480 (source)
// and :3481-3600 (drain) are the same code with S/D parameter sets.
struct JctPar {
  real Nvtm, Isat, xjbv, bv, XExpBV, vjmFwd, vjmRev, IVjmFwd, IVjmRev, slpFwd, slpRev;
};
XB_HD void junction_diode(int dioMod, const JctPar &j, real vj, real gmin,
                          real &g, real &c) {
  if (j.Isat <= 0.0) {
    g = gmin;
    c = g * vj;
    return;
  }
  switch (dioMod) {
    case 0: {
      const real ev = exp(vj / j.Nvtm);
      const real t1 = j.xjbv * exp(-(j.bv + vj) / j.Nvtm);
      g = j.Isat * (ev + t1) / j.Nvtm + gmin;
      c = j.Isat * (ev + j.XExpBV - t1 - 1.0) + gmin * vj;
    } break;
    case 1: {
      const real t2 = vj / j.Nvtm;
      if (t2 < -kExpThr) {
        g = gmin;
        c = j.Isat * (kMinExp - 1.0) + gmin * vj;
      } else if (vj <= j.vjmFwd) {
        const real ev = exp(t2);
        g = j.Isat * ev / j.Nvtm + gmin;
        c = j.Isat * (ev - 1.0) + gmin * vj;
      } else {
        const real t0 = j.IVjmFwd / j.Nvtm;
        g = t0 + gmin;
        c = j.IVjmFwd - j.Isat + t0 * (vj - j.vjmFwd) + gmin * vj;
      }
    } break;
    case 2: {
      if (vj < j.vjmRev) {
        const real t0 = vj / j.Nvtm;
        real ev, dev;
        if (t0 < -kExpThr) { ev = kMinExp; dev = 0.0; }
        else { ev = exp(t0); dev = ev / j.Nvtm; }
        const real t1 = ev - 1.0;
        const real t2 = j.IVjmRev + j.slpRev * (vj - j.vjmRev);
        g = dev * t2 + t1 * j.slpRev + gmin;
        c = t1 * t2 + gmin * vj;
      } else if (vj <= j.vjmFwd) {
        const real t0 = vj / j.Nvtm;
        real ev, dev;
        if (t0 < -kExpThr) { ev = kMinExp; dev = 0.0; }
        else { ev = exp(t0); dev = ev / j.Nvtm; }
        const real t1 = (j.bv + vj) / j.Nvtm;
        real t2, t3;
        if (t1 > kExpThr) { t2 = kMinExp; t3 = 0.0; }
        else { t2 = exp(-t1); t3 = -t2 / j.Nvtm; }
        g = j.Isat * (dev - j.xjbv * t3) + gmin;
        c = j.Isat * (ev + j.XExpBV - 1.0 - j.xjbv * t2) + gmin * vj;
      } else {
        g = j.slpFwd + gmin;
        c = j.IVjmFwd + j.slpFwd * (vj - j.vjmFwd) + gmin * vj;
      }
    } break;
    default: break;
  }
}

// Trap-assisted tunnelling factor for one junction component (B4p82.C:3608-3690):
// returns T = DEXP(arg) and its derivative w.r.t. the junction voltage.
XB_HD void tat_term(real vts, real nvtmr, real vj, real &t, real &dt_dvb) {
  real t0, t9, t10;
  if ((vts - vj) < (vts * 1e-3)) {
    t9 = 1.0e3;
    t0 = -vj / nvtmr * t9;
    dexp(t0, t, t10);
    dt_dvb = t10 / nvtmr * t9;
  } else {
    t9 = 1.0 / (vts - vj);
    t0 = -vj / nvtmr * vts * t9;
    const real dt0_dvb = vts / nvtmr * (t9 + vj * t9 * t9);
    dexp(t0, t, t10);
    dt_dvb = t10 * dt0_dvb;
  }
}

// value-in / value-out entry points of the helpers above (real functions in the fast device build, see XB_HELPER)
XB_HELPER Real2 junction_diode_v(int dioMod, JctPar j, real vj, real gmin) {
  Real2 r; junction_diode(dioMod, j, vj, gmin, r.a, r.b); return r;
}
XB_HELPER Real2 tat_term_v(real vts, real nvtmr, real vj) { Real2 r; tat_term(vts, nvtmr, vj, r.a, r.b); return r; }
XB_HELPER Real2 poly_depletion_v(real phi, real ngate, real epsgate, real coxe, real vg) {
  Real2 r; poly_depletion(phi, ngate, epsgate, coxe, vg, r.a, r.b); return r;
}

// ---------------------------------------------------------------------------
// Stage 1: terminal voltages, initial-condition overrides, Newton limiting.
//   V[]      node voltages in general-stamp order (see B4Node)
//   sto_old  the 13 limiting voltages of the previous iterate / previous step
//            (caller picks currSto vs nextSto exactly as B4p82.C:3096-3148)
// ---------------------------------------------------------------------------
XB_HD void stage_voltages(const SolverFlags &S, const B4Model &M, const B4Inst &I,
                          const real *V, const real *sto_old, bool have_old,
                          real von_prev, B4Mid &W) {
  const real Vd = V[kD], Vs = V[kS], Vb = V[kB], Vsp = V[kSP], Vdp = V[kDP];
  const real Vgp = V[kGP], Vbp = V[kBP], Vge = V[kGE];
  const real Vgm = V[kGM];   // li_GateMid aliases li_GateExt unless rgateMod == 3 (B4.C:6228-6235)
  const real Vdb = V[kDB], Vsb = V[kSB];
  const real Qtotal = I.trnqsMod ? V[kQ] : 0.0;
  const real ty = real(M.dtype);

  W.Vddp = Vd - Vdp;   W.Vssp = Vs - Vsp;
  W.Vdbb = Vdb - Vb;   W.Vdbbp = Vdb - Vbp;
  W.Vsbb = Vsb - Vb;   W.Vsbbp = Vsb - Vbp;
  W.Vbpb = Vbp - Vb;
  W.Vgegp = Vge - Vgp; W.Vgegm = Vge - Vgm; W.Vgmgp = Vgm - Vgp;

  real vds = ty * (Vdp - Vsp), vgs = ty * (Vgp - Vsp), vbs = ty * (Vbp - Vsp);
  real vges = ty * (Vge - Vsp), vgms = ty * (Vgm - Vsp);
  real vdbs = ty * (Vdb - Vsp), vsbs = ty * (Vsb - Vsp);
  real vses = ty * (Vs - Vsp), vdes = ty * (Vd - Vsp);
  real qdef = ty * Qtotal;
  real vbd = vbs - vds, vgd = vgs - vds;
  real vged = vges - vds, vgmd = vgms - vds, vdbd = vdbs - vds;

  int origFlag = 1;
  W.vbd_orig = vbd; W.vbs_orig = vbs; W.vgs_orig = vgs; W.vds_orig = vds;
  W.vgd_orig = vgd; W.vges_orig = vges; W.vgms_orig = vgms; W.vdes_orig = vdes;
  W.vses_orig = vses; W.vdbs_orig = vdbs; W.vsbs_orig = vsbs; W.vdbd_orig = vdbd;
  W.vged_orig = vged; W.vgmd_orig = vgmd;
  W.vbs_jct_orig = (!I.rbodyMod) ? vbs : vsbs;
  W.vbd_jct_orig = (!I.rbodyMod) ? vbd : vdbd;
  W.vgmb_orig = vgms - vbs;
  W.vgb_orig = vgs - vbs;

  if (S.initJctFlag && !I.OFF && S.voltageLimiterFlag) {
    // (inputOPFlag path of the reference needs the host flag vector; the C-ABI
    //  rejects inputOPFlag, see capi.)
    vds = 0.1; vdes = 0.11; vses = -0.01;
    vgs = vges = vgms = ty * I.vth0 + 0.1;
    origFlag = 0;
    vbs = vdbs = vsbs = 0.0;
    vbd = vbs - vds; vdbd = vdbs - vds; vgd = vgs - vds;
    vged = vges - vds; vgmd = vgms - vds;
  } else if ((S.initFixFlag || S.initJctFlag) && I.OFF) {
    vds = vgs = vbs = vges = vgms = 0.0;
    vds = vsbs = vdes = vses = qdef = 0.0;
  }

  real o_vbd, o_vbs, o_vgs, o_vds, o_vges, o_vgms, o_vdes, o_vses, o_vdbs, o_vsbs, o_vdbd, o_vged, o_vgmd;
  if (have_old) {
    o_vbd = sto_old[0]; o_vbs = sto_old[1]; o_vgs = sto_old[2]; o_vds = sto_old[3];
    o_vges = sto_old[4]; o_vgms = sto_old[5]; o_vdes = sto_old[6]; o_vses = sto_old[7];
    o_vdbs = sto_old[8]; o_vsbs = sto_old[9]; o_vdbd = sto_old[10]; o_vged = sto_old[11];
    o_vgmd = sto_old[12];
  } else {
    o_vbd = vbd; o_vbs = vbs; o_vgs = vgs; o_vds = vds; o_vges = vges; o_vgms = vgms;
    o_vdes = vdes; o_vses = vses; o_vdbs = vdbs; o_vsbs = vsbs; o_vdbd = vdbd;
    o_vged = vged; o_vgmd = vgmd;
  }
  const real o_vgd = o_vgs - o_vds;

  int limited = 0;
  if (S.voltageLimiterFlag && !(S.initFixFlag && I.OFF)) {
    int Check = 0, Check1 = 0, Check2 = 0;
    const real vonl = von_prev;
    if (S.newtonIter >= 0 && !S.initJctFlag) {
      if (o_vds >= 0.0) {
        vgs = fetlim(vgs, o_vgs, vonl);
        vds = vgs - vgd;
        vds = limvds(vds, o_vds);
        vgd = vgs - vds;
        if (I.rgateMod == 3) {
          vges = fetlim(vges, o_vges, vonl);
          vgms = fetlim(vgms, o_vgms, vonl);
          vged = vges - vds;
          vgmd = vgms - vds;
        } else if ((I.rgateMod == 1) || (I.rgateMod == 2)) {
          vges = fetlim(vges, o_vges, vonl);
          vged = vges - vds;
        }
        if (M.rdsMod) {
          vdes = limvds(vdes, o_vdes);
          vses = -limvds(-vses, -o_vses);
        }
      } else {
        vgd = fetlim(vgd, o_vgd, vonl);
        vds = vgs - vgd;
        vds = -limvds(-vds, -o_vds);
        vgs = vgd + vds;
        if (I.rgateMod == 3) {
          vged = fetlim(vged, o_vged, vonl);
          vges = vged + vds;
          vgmd = fetlim(vgmd, o_vgmd, vonl);
          vgms = vgmd + vds;
        }
        if ((I.rgateMod == 1) || (I.rgateMod == 2)) {
          vged = fetlim(vged, o_vged, vonl);
          vges = vged + vds;
        }
        if (M.rdsMod) {
          vdes = -limvds(-vdes, -o_vdes);
          vses = limvds(vses, o_vses);
        }
      }
      if (vds >= 0.0) {
        vbs = pnjlim(vbs, o_vbs, kVt0, M.vcrit, Check);
        vbd = vbs - vds;
        if (I.rbodyMod) {
          vdbs = pnjlim(vdbs, o_vdbs, kVt0, M.vcrit, Check1);
          vdbd = vdbs - vds;
          vsbs = pnjlim(vsbs, o_vsbs, kVt0, M.vcrit, Check2);
          if ((Check1 != 0) || (Check2 != 0)) Check = 1;
        }
      } else {
        vbd = pnjlim(vbd, o_vbd, kVt0, M.vcrit, Check);
        vbs = vbd + vds;
        if (I.rbodyMod) {
          vdbd = pnjlim(vdbd, o_vdbd, kVt0, M.vcrit, Check1);
          vdbs = vdbd + vds;
          const real o_vsbd = o_vsbs - o_vds;
          real vsbd = vsbs - vds;
          vsbd = pnjlim(vsbd, o_vsbd, kVt0, M.vcrit, Check2);
          vsbs = vsbd + vds;
          if ((Check1 != 0) || (Check2 != 0)) Check = 1;
        }
      }
    }
    if (Check == 1) limited = 1;
  }

  vbd = vbs - vds; vgd = vgs - vds;
  vged = vges - vds; vgmd = vgms - vds; vdbd = vdbs - vds;

  if (S.voltageLimiterFlag) {
    if (fabs(W.vbs_orig - vbs) > kMachEps || fabs(W.vgs_orig - vgs) > kMachEps ||
        fabs(W.vds_orig - vds) > kMachEps || fabs(W.vges_orig - vges) > kMachEps ||
        fabs(W.vgms_orig - vgms) > kMachEps || fabs(W.vdes_orig - vdes) > kMachEps ||
        fabs(W.vses_orig - vses) > kMachEps || fabs(W.vdbs_orig - vdbs) > kMachEps ||
        fabs(W.vsbs_orig - vsbs) > kMachEps)
      origFlag = 0;
  }

  W.vds = vds; W.vgs = vgs; W.vbs = vbs; W.vbd = vbd; W.vgd = vgd;
  W.vges = vges; W.vgms = vgms; W.vdes = vdes; W.vses = vses;
  W.vdbs = vdbs; W.vsbs = vsbs; W.vdbd = vdbd; W.vged = vged; W.vgmd = vgmd;
  W.vgb = vgs - vbs; W.vgmb = vgms - vbs;
  W.vbs_jct = (!I.rbodyMod) ? vbs : vsbs;
  W.vbd_jct = (!I.rbodyMod) ? vbd : vdbd;
  W.qdef = qdef;
  W.origFlag = origFlag;
  W.limitedFlag = limited;
}

}  // namespace b4
}  // namespace xb

#include "bsim4_eval_dc.h"
#include "bsim4_eval_cv.h"
