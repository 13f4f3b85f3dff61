"""Small runs of the one-launch assembly (rail chunks + ticket finisher), the LU kernels and the transient driver with
Gear / DCOP: target of compute-sanitizer memcheck and racecheck."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from xyce_b200 import workloads as wl
from xyce_b200.capi import SolverState
ss = SolverState(transientFlag=1, newtonIter=1)
for n_inv in (3, 700, 2100):           # 2100 inverters: several chunks per rail destination
    w = wl.inverter_array(n_inv, store_noise=0.3)
    eng = wl.build_engine(w)
    outs = []
    for rep in range(2):               # same carried state (store, von) before each pass
        eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
        outs.append(eng.load_host(w["x"], ss))
    a, b = outs
    for k in a:
        assert np.array_equal(a[k], b[k]), k          # bitwise reproducible, ticket counters reset
    eng.close()
w = wl.ring_oscillator_array(3, 11)
eng = wl.build_engine(w)
for method in (7, 8):
    r = eng.tran_run(w["x"], 2e-10, 1e-12, [0, 1], method=method)
    assert r["rc"] == 0, r
eng.close()
print("sanitize run ok")
# round 2: the pipelined host path (windowed assembly launches, second stream), the Newton step behind host buffers in
# plain and bordered form, the batched LU kernels, and the generic kernel of two translated ADMS models
import ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
w = wl.inverter_array(2100, store_noise=0.3)
for border in (0, 1):
    eng = wl.build_engine(w)
    if border:
        eng.border_set(1)
    for pipe in (0, 1):
        eng.set_option("pipeline_host", pipe)
        eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
        r, J = eng.load_host_jr(w["x"], ss, 1e12, 0.5)
    dx = eng.newton_step_host(w["x"], ss, 1e12, 0.5)
    dx = eng.newton_step_host(w["x"] + 0.01, ss, 1e12, 0.5)
    assert np.all(np.isfinite(dx))
    eng.close()
w = wl.ring_oscillator_array(40, 31)          # 40 equal BTF blocks: one batched group
eng = wl.build_engine(w)
r = eng.tran_run(w["x"], 5e-11, 1e-12, [0, 1])
assert r["rc"] == 0
eng.close()
try:
    import oracle_ref, xyce_b200
    from adms_common import adms_circuit, bias_vector
    gen = {m["name"]: m for m in xyce_b200.capi.Engine.adms_gen_models()}
    for model, card in (("ekv_va", "nmos"), ("hicumL2va", "res")):
        info = gen[model]
        ref = adms_circuit(oracle_ref.RefCircuit, model, card, info["ext"], n_dev=40, seed=1)
        ex = [ref.adms_export(i, model) for i in range(ref.n_inst)]
        eng = xyce_b200.Engine(0)
        eng.set_pattern(ref.rowptr, ref.colind); eng.set_sizes(ref.n_sta, ref.n_sto)
        eng.add_simple_group(info["type"], np.array([e["rec"] for e in ex]), [0] * len(ex), np.array([e["lids"] for e in ex]),
                             [e["sto0"] for e in ex], 1, [e["sta0"] for e in ex], 1)
        eng.finalize()
        x = bias_vector(model, ref.n, [e["lids"] for e in ex], np.random.default_rng(2))
        got = eng.load_host(x, SolverState(transientFlag=1, newtonIter=1))
        assert np.all(np.isfinite(got["dFdx"]))
        eng.close()
except ImportError:
    pass
print("sanitize round-2 paths ok")
