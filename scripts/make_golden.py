"""Generate tests/golden/b4_cases.npz: seeded inputs and the outputs of the reference's own BSIM4 objects
(oracle/_ref, i.e. N_DEV_MOSFET_B4*.C compiled in place) for every model-card variant x solver-flag case.
Run where /root/reference exists; the fixture is committed."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import oracle_ref  # noqa: E402
from b4_common import VARIANTS, isolated_devices, records_from_ref  # noqa: E402

FLAG_CASES = {
    "tran_iter1": dict(transient=1, newtonIter=1),
    "tran_iter0_init": dict(transient=1, newtonIter=0, initTran=1),
    "dcop_initjct": dict(dcop=1, tranop=1, transient=1, initJct=1, newtonIter=0),
    "dc_nocharge": dict(dcop=1, newtonIter=1),
}
out = {}
names = []
for vi, variant in enumerate(sorted(VARIANTS)):
    for ci, (case, flags) in enumerate(sorted(FLAG_CASES.items())):
        ref = isolated_devices(oracle_ref.RefCircuit, 3, variant, seed=10 + vi)
        rng = np.random.default_rng(1000 * vi + ci)
        x = rng.uniform(-0.3, 1.3, ref.n)
        nsto, csto = rng.normal(0.3, 0.3, ref.n_sto), rng.normal(0.3, 0.3, ref.n_sto)
        von = rng.uniform(0.2, 0.6, ref.n_inst)
        ref.set_flags(**flags)
        ref.set_state(curr_sto=csto, next_sto=nsto, curr_sta=np.zeros(ref.n_sta))
        ref.set_von(von)
        want = ref.load(x)
        st = ref.get_state()
        key = "%s__%s" % (variant, case)
        names.append(key)
        rec = records_from_ref(ref)
        for k, v in rec.items():
            out[key + "/rec_" + k] = v
        for k, v in dict(x=x, nsto=nsto, csto=csto, von=von, rowptr=ref.rowptr, colind=ref.colind,
                         flags=np.array([flags.get(f, 1 if f == "voltageLimiter" else 0) for f in oracle_ref.FLAG_NAMES]),
                         next_sto=st["next_sto"], next_sta=st["next_sta"], curr_sta=st["curr_sta"],
                         von_out=ref.get_von(), n_sta=np.array(ref.n_sta), n_sto=np.array(ref.n_sto)).items():
            out[key + "/" + k] = v
        for k, v in want.items():
            out[key + "/ref_" + k] = v
out["cases"] = np.array(names)
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "b4_cases.npz"), **out)
print(len(names), "cases")
