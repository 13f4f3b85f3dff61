python - <<'PY' 2>&1 | grep -v Netlist | tail -30
import sys; sys.path.insert(0,'tests')
import numpy as np, xyce_b200
from b4_common import load_golden, FLAG_NAMES, solver_state, rel_err
GOLD = load_golden()
for case in ["mob3_v461__dc_nocharge","mob3_v461__tran_iter1","mob3__dc_nocharge"]:
  for arith in (2,0):
    g = GOLD[case]
    eng = xyce_b200.Engine(0)
    eng.set_option("b4_arith", arith)
    eng.set_pattern(g["rowptr"], g["colind"])
    eng.set_sizes(int(g["n_sta"]), int(g["n_sto"]))
    eng.b4_set_models(g["rec_model_d"], g["rec_model_i"], g["rec_size_d"])
    eng.b4_add_group(g["rec_inst_d"], g["rec_inst_i"], g["rec_model_idx"], g["rec_size_idx"], g["rec_lids"], g["rec_sto0"], 1, g["rec_sta0"], 1)
    eng.finalize()
    eng.set_state(0, g["nsto"]); eng.set_state(1, g["csto"]); eng.b4_set_von(0, g["von"])
    flags = dict(zip(FLAG_NAMES, [int(v) for v in g["flags"]]))
    got = eng.load_host(g["x"], solver_state(**flags))
    print(case, arith, {k: rel_err(got[k], g["ref_"+k], 1e-3*np.max(np.abs(g["ref_"+k])) if np.any(g["ref_"+k]) else 1e-300) for k in ("f","q","dFdx","dQdx")})
    print(" sto", rel_err(eng.get_state(0), g["next_sto"], 1e-30), "sta", rel_err(eng.get_state(2), g["next_sta"], 1e-30))
    print(" sta got", eng.get_state(2)[:6], "want", g["next_sta"][:6])
    print(" mobMod", g["rec_model_i"][:, 8], "names?")
    eng.close()
PY
