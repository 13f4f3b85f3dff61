"""Timing of end-to-end .TRAN runs (device eval + assembly + LU + Newton/OneStep) on the GPU, next to the same
control flow on the reference device code + Kundert Sparse on one host core (bounded size)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from xyce_b200 import workloads as wl

def gpu_run(n_rings, stages, tstop):
    w = wl.ring_oscillator_array(n_rings, stages)
    t0 = time.perf_counter()
    eng = wl.build_engine(w)
    if os.environ.get("XYCE_B200_LU_GRAPHS"):
        eng.set_option("lu_graphs", int(os.environ["XYCE_B200_LU_GRAPHS"]))
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    r = eng.tran_run(w["x"], tstop, 1e-12, [0, 1, w["vdd"]])
    dt = time.perf_counter() - t0
    info = eng.lu_info()
    launches = eng.launch_count()
    eng.close()
    s = r["stats"]
    return dict(impl="gpu", mosfets=w["n_inst"], unknowns=w["n_unknowns"], tstop=tstop, rc=r["rc"], setup_s=t_setup, wall_s=dt,
                attempts=s["attempts"], accepted=s["accepted"], newton_iters=s["newton_iters"],
                ms_per_newton_iter=1e3 * dt / max(s["newton_iters"], 1), ms_per_step=1e3 * dt / max(s["attempts"], 1),
                lu=info, launches=launches, lu_analyses=s["lu_analyses"], lu_refactors=s["lu_refactors"],
                setup_s_inside=s["setup_s"], run_s_inside=s["run_s"], max_readback_wait_s=s["max_readback_wait_s"])

def cpu_run(n_rings, stages, tstop):
    import oracle_ref
    from b4_common import ref_circuit_from_workload
    w = wl.ring_oscillator_array(n_rings, stages)
    ref = ref_circuit_from_workload(oracle_ref.RefCircuit, w)
    ref.set_flags(transient=1)
    t0 = time.perf_counter()
    r = ref.tran_run(w["x"], tstop, 1e-12, [0, 1, w["vdd"]], w["linear"], w["sources"])
    dt = time.perf_counter() - t0
    s = r["stats"]
    return dict(impl="reference devices + ksparse, 1 core", mosfets=w["n_inst"], unknowns=w["n_unknowns"], tstop=tstop, rc=r["rc"],
                wall_s=dt, attempts=s["attempts"], accepted=s["accepted"], newton_iters=s["newton_iters"],
                ms_per_newton_iter=1e3 * dt / max(s["newton_iters"], 1), ms_per_step=1e3 * dt / max(s["attempts"], 1))

out = []
for (nr, st, ts, cpu) in [(1, 101, 2e-9, True), (50, 101, 2e-10, True), (495, 101, 2e-10, False), (4950, 101, 1e-10, False)]:
    if len(sys.argv) > 1 and nr > int(sys.argv[1]):
        continue
    r = gpu_run(nr, st, ts); out.append(r); print(json.dumps(r), flush=True)
    if cpu and not os.environ.get("XYCE_B200_NO_CPU"):
        r = cpu_run(nr, st, ts); out.append(r); print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tran_bench.json"), "w"), indent=1)
