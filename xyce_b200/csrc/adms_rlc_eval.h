// xyce_b200 -- ADMS-shaped device evaluation, demonstrated on the reference's plugin example
// user_plugin/rlc.va (series R-L-C with two internal nodes and one branch-current unknown).
//
// The reference runs Verilog-A models through admsXml templates (utils/ADMS/xyceImplementationFile_nosac.xml,
// xyceBasicTemplates_nosac.xml); every generated device has the same shape (example:
// src/DeviceModelPKG/ADMS/N_DEV_ADMSHBT_X.C:3400-3401, :4217-4221, :4344-4347, :4363; loads :3011, :3082):
//   * probeVars[] = branch voltages / branch currents read from the solution, with d_probeVars = 1
//   * staticContributions[node], dynamicContributions[node] and their probe derivatives
//   * a potential source  V(a,b) <+ rhs  adds the branch current to KCL of a (+) and b (-), and the branch
//     equation row receives  rhs - (V(a) - V(b))
//   * loads are pure copies: F += static, Q += dynamic, dFdx/dQdx += probe derivatives chained to node columns.
// AdmsContrib reproduces that shape generically; rlc::evaluate is what the templates emit for rlc.va:46-69.
#pragma once
#include "xb_common.h"

namespace xb {
namespace adms {

template <int NODES, int PROBES>
struct AdmsContrib {
  real stat[NODES], dyn[NODES], d_stat[NODES][PROBES], d_dyn[NODES][PROBES];
  XB_HD void clear() {
    for (int i = 0; i < NODES; ++i) {
      stat[i] = dyn[i] = 0.0;
      for (int p = 0; p < PROBES; ++p) d_stat[i][p] = d_dyn[i][p] = 0.0;
    }
  }
};

namespace rlc {
// unknowns in admsNodeID order: p, n, internal1, internal2, branch current of (internal2, n)
enum { kP = 0, kN = 1, kI1 = 2, kI2 = 3, kBr = 4, kNodes = 5 };
// probes: V(p,internal1), V(internal1,internal2), I(internal2,n)
enum { pV_p_i1 = 0, pV_i1_i2 = 1, pI_i2_n = 2, kProbes = 3 };
// Jacobian slots (row, col)
enum { sPP = 0, sPI1, sI1P, sI1I1, sI1I2, sI2I1, sI2I2, sI2Br, sNBr, sBrI2, sBrN, sBrBr, kSlots };
XB_HD constexpr int slot_row(int s) { constexpr int t[kSlots] = {0, 0, 2, 2, 2, 3, 3, 3, 1, 4, 4, 4}; return t[s]; }
XB_HD constexpr int slot_col(int s) { constexpr int t[kSlots] = {0, 2, 0, 2, 3, 2, 3, 4, 4, 3, 1, 4}; return t[s]; }
constexpr int kNumFields = 3;   // R, L, C  (rlc.va:51-53)

struct Out { real F[kNodes], Q[kNodes], JF[kSlots], JQ[kSlots]; };

XB_HD void evaluate(real R, real L, real C, const real *V, Out &o) {
  AdmsContrib<kNodes, kProbes> c;
  c.clear();
  real probe[kProbes];
  probe[pV_p_i1] = V[kP] - V[kI1];
  probe[pV_i1_i2] = V[kI1] - V[kI2];
  probe[pI_i2_n] = V[kBr];
  // I(p,internal1) <+ V(p,internal1)/R;
  c.stat[kP] += probe[pV_p_i1] / R;       c.d_stat[kP][pV_p_i1] += 1.0 / R;
  c.stat[kI1] -= probe[pV_p_i1] / R;      c.d_stat[kI1][pV_p_i1] -= 1.0 / R;
  // CapacitorCharge = V(internal1,internal2)*C;  I(internal1,internal2) <+ ddt(CapacitorCharge);
  const real q = probe[pV_i1_i2] * C;
  c.dyn[kI1] += q;                        c.d_dyn[kI1][pV_i1_i2] += C;
  c.dyn[kI2] -= q;                        c.d_dyn[kI2][pV_i1_i2] -= C;
  // InductorCurrent = I(internal2,n);  V(internal2,n) <+ L*ddt(InductorCurrent);
  c.dyn[kBr] += probe[pI_i2_n] * L;       c.d_dyn[kBr][pI_i2_n] += L;
  // finish-up of the potential source: branch current into KCL, branch equation minus the node voltages
  c.stat[kI2] += probe[pI_i2_n];          c.d_stat[kI2][pI_i2_n] += 1.0;
  c.stat[kN] -= probe[pI_i2_n];           c.d_stat[kN][pI_i2_n] -= 1.0;
  c.stat[kBr] -= V[kI2] - V[kN];
  // ---- loads: copies + probe -> node-column chain rule ----
  for (int i = 0; i < kNodes; ++i) { o.F[i] = c.stat[i]; o.Q[i] = c.dyn[i]; }
  for (int s = 0; s < kSlots; ++s) o.JF[s] = o.JQ[s] = 0.0;
  o.JF[sPP] += c.d_stat[kP][pV_p_i1];    o.JF[sPI1] -= c.d_stat[kP][pV_p_i1];
  o.JF[sI1P] += c.d_stat[kI1][pV_p_i1];  o.JF[sI1I1] -= c.d_stat[kI1][pV_p_i1];
  o.JQ[sI1I1] += c.d_dyn[kI1][pV_i1_i2]; o.JQ[sI1I2] -= c.d_dyn[kI1][pV_i1_i2];
  o.JQ[sI2I1] += c.d_dyn[kI2][pV_i1_i2]; o.JQ[sI2I2] -= c.d_dyn[kI2][pV_i1_i2];
  o.JF[sI2Br] += c.d_stat[kI2][pI_i2_n];
  o.JF[sNBr] += c.d_stat[kN][pI_i2_n];
  o.JQ[sBrBr] += c.d_dyn[kBr][pI_i2_n];
  o.JF[sBrI2] -= 1.0;
  o.JF[sBrN] += 1.0;
}
}  // namespace rlc
}  // namespace adms
}  // namespace xb
