"""Timing of the small-device evaluation kernels (diode, MOSFET level 1, BJT, ADMS MVS) at 1M instances each: records
exported from the reference objects (oracle/_ref), tiled; CUDA events around xgpu_update_state, L2 flushed.
usage: simple_kernels_timing.py [n_instances] [out.json]"""
import sys, os, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle_ref, xyce_b200
from xyce_b200.capi import SolverState
from dev_common import SIMPLE, simple_circuit, diode_circuit, DIODE_CARDS, DIODE_SLOT_ROW, DIODE_SLOT_COL

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
out_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/simple_kernels.json"
res = []
cases = [("diode", 1, None), ("mos1", 2, "basic"), ("bjt", 3, "basic"), ("mvs", 5, "nmos")]
# models written by the ADMS translator (generic kernel adms_gen_kernel<Traits>)
from adms_common import ADMS_CARDS, adms_circuit
GEN = {m["name"]: m for m in xyce_b200.capi.Engine.adms_gen_models()}
cases += [("adms:" + n, GEN[n]["type"], sorted(ADMS_CARDS[n])[0]) for n in sorted(GEN) if n in ADMS_CARDS]
for kind, tid, card in cases:
    if kind.startswith("adms:"):
        g = GEN[kind[5:]]
        ref = adms_circuit(oracle_ref.RefCircuit, g["name"], card, g["ext"], n_dev=4, seed=1)
        ex = [ref.adms_export(i, g["name"]) for i in range(ref.n_inst)]
        nodes, nstore, nstate, srow, scol = g["nodes"], 0, 0, g["slot_row"], g["slot_col"]
    elif kind == "diode":
        ref = diode_circuit(oracle_ref.RefCircuit, sorted(DIODE_CARDS)[0], n_dev=4, seed=1)
        ex = [ref.diode_export(i) for i in range(ref.n_inst)]
        nodes, nstore, nstate, srow, scol = 3, 3, 0, DIODE_SLOT_ROW, DIODE_SLOT_COL
    else:
        _, key, nodes, nstore, nstate, srow, scol = SIMPLE[kind]
        ref = simple_circuit(oracle_ref.RefCircuit, kind, card, n_dev=4, seed=1)
        ex = [ref.dev_export(i, key) for i in range(ref.n_inst)]
    e0 = ex[0]
    nn = len(e0["lids"])
    rec = np.tile(e0["rec"], (N, 1))
    lids = (np.arange(N)[:, None] * nn + np.arange(nn)[None, :]).astype(np.int32)
    n_unk = N * nn
    # diagonal-block pattern: every device couples only its own nodes
    cols = (np.repeat(np.arange(N) * nn, nn * nn).reshape(N, nn, nn) + np.arange(nn)[None, None, :]).reshape(-1)
    rowptr = np.arange(0, n_unk * nn + 1, nn, dtype=np.int32)
    eng = xyce_b200.Engine(0)
    eng.set_pattern(rowptr, cols.astype(np.int32))
    eng.set_sizes(max(N * nstate, 1), max(N * nstore, 1))
    eng.add_simple_group(tid, rec, [e0["flags"]] * N, lids, np.arange(N) * max(nstore, 1) if nstore else np.zeros(N), 1,
                         np.arange(N) * max(nstate, 1) if nstate else np.zeros(N), 1)
    eng.finalize()
    stream = torch.cuda.current_stream(); eng.set_stream(stream.cuda_stream)
    ss = SolverState(transientFlag=1, newtonIter=1)
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.3, 0.8, n_unk)
    if kind == "adms:mvs_2_0_0_hemt":
        x[5::nn] = rng.uniform(-0.6, -0.02, N)      # V(sf) window of the HEMT variant (tests/adms_common.py)
    eng.load_host(x, ss)
    b = [eng.device_buffer(i) for i in range(11)]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    ts = []
    for it in range(12):
        flush.fill_(0.0)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream); eng.update_state(b[0], b[9], b[10], b[7], b[8], ss); a1.record(stream)
        torch.cuda.synchronize()
        ts.append(a0.elapsed_time(a1))
    ms = float(np.median(ts[2:]))
    nfields = rec.shape[1]
    slots = len(srow)
    # algorithmic bytes per instance: record + node LIDs and voltages + store/state old+new + 4 vector planes x nodes + 2 matrix planes x slots
    bytes_per = 8 * nfields + 12 * nn + 16 * nstore + 8 * nstate + 32 * nn + 16 * slots
    r = dict(device=kind, instances=N, fields=nfields, nodes=nn, slots=slots, eval_ms=ms, evals_per_s=N / (ms * 1e-3),
             algorithmic_bytes_per_eval=bytes_per, achieved_GBps=bytes_per * N / (ms * 1e-3) / 1e9)
    res.append(r); print(json.dumps(r), flush=True)
    eng.close()
os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)
