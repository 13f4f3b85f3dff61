import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import test_gpu_bsim4_parity as T
import xyce_b200
from b4_common import FLAG_NAMES
for case in sys.argv[1:]:
    g = T.GOLD[case]
    for arith in (0, 1, 2):
        eng = xyce_b200.Engine(0)
        eng.set_option("b4_arith", arith)
        eng.set_pattern(g["rowptr"], g["colind"]); eng.set_sizes(int(g["n_sta"]), int(g["n_sto"]))
        eng.b4_set_models(g["rec_model_d"], g["rec_model_i"], g["rec_size_d"])
        eng.b4_add_group(g["rec_inst_d"], g["rec_inst_i"], g["rec_model_idx"], g["rec_size_idx"], g["rec_lids"], g["rec_sto0"], 1, g["rec_sta0"], 1)
        eng.finalize()
        eng.set_state(0, g["nsto"]); eng.set_state(1, g["csto"]); eng.b4_set_von(0, g["von"])
        flags = dict(zip(FLAG_NAMES, [int(v) for v in g["flags"]]))
        got = eng.load_host(g["x"], T.solver_state(**flags))
        for k in ("f", "q", "dFdxdVp", "dQdxdVp", "dFdx", "dQdx"):
            want = g["ref_" + k]
            scale = 1e-3 * np.max(np.abs(want)) if np.any(want) else 1e-300
            e = np.abs(got[k] - want) / np.maximum(np.abs(want), scale)
            e = np.where(np.isnan(e), np.inf, e)
            i = int(np.argmax(e))
            print(case, "arith", arith, k, "max err %.3e at %d got %.17g want %.17g scale %.3g" % (e[i], i, got[k][i], want[i], scale))
        for nm, a, b in (("sto", eng.get_state(0), g["next_sto"]), ("sta", eng.get_state(2), g["next_sta"])):
            e = np.abs(a - b) / np.maximum(np.abs(b), 1e-30); e = np.where(np.isnan(e), np.inf, e); i = int(np.argmax(e))
            print(case, "arith", arith, nm, "max err %.3e at %d got %.17g want %.17g" % (e[i], i, a[i], b[i]))
        eng.close()
