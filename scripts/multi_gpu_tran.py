"""Multi-GPU .TRAN inside the library (BASELINE config 3 / 4 shape): graph partition of the ring-oscillator array, one
process per GPU, NCCL communicator attached to the engine (xgpu_comm_init), border = the unknowns shared between
partitions (xgpu_border_set).  xgpu_tran_run then runs the whole Newton / OneStep loop distributed: local evaluation and
assembly, border rows summed with ncclAllReduce, BTF blocks of the interior factored per GPU, the border (Schur) system
all-reduced and solved redundantly, norms / convergence flags combined in one all-gather per evaluation.

Checks (rank 0 prints one JSON line):
  * --check-single 1: the same circuit on ONE GPU -- identical step sequence and Newton counts, waveforms at 1e-9;
  * --check-oracle K: K rings per rank against the single-ring oracle (reference BSIM4 objects + Kundert Sparse) replayed
    on the distributed run's accepted steps: Newton counts identical, waveforms within RELTOL / ABSTOL.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/multi_gpu_tran.py --rings 4950
"""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from xyce_b200 import partition as pt, workloads as wl
from xyce_b200.capi import Engine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rings", type=int, default=400)
    ap.add_argument("--stages", type=int, default=101)
    ap.add_argument("--tstop", type=float, default=2e-10)
    ap.add_argument("--check-single", type=int, default=1)
    ap.add_argument("--check-oracle", type=int, default=2)
    ap.add_argument("--graph-partition", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=2, help="timed repetitions after the first run (LU already analysed)")
    ap.add_argument("--p2p", type=int, default=1, help="1 = small collectives through NVLink peer mailboxes, 0 = NCCL calls")
    ap.add_argument("--direct", type=int, default=0, help="1 = every rank builds only its own rings (large arrays)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                 # plumbing only (id exchange, barriers, gathering the report); the data path is the library's own NCCL communicator
    if a.direct:
        # large arrays (BASELINE config 4): every rank builds ITS rings only -- the partition the graph partitioner finds for
        # this netlist (whole rings per rank, supply node + source branch shared) -- without materialising the 10M-instance
        # global arrays on every rank.  The supply source's linear stamps belong to rank 0 (counted once by the reduction).
        shifts = np.random.default_rng(12345).integers(0, a.stages, a.rings)
        b = pt.split_ranges(a.rings, world)
        w = wl.ring_oscillator_array(b[rank + 1] - b[rank], a.stages, shifts=shifts[b[rank]:b[rank + 1]])
        if rank != 0:
            for k in ("g_row", "g_col", "g_val"):
                w["linear"][k] = w["linear"][k][:0]
        nloc = (b[rank + 1] - b[rank]) * a.stages
        w["n_shared"] = 2
        w["owner"] = None
        wg = dict(n_inst=2 * a.rings * a.stages, n_unknowns=a.rings * a.stages + 2, shift=shifts, vdd=a.rings * a.stages, branch=a.rings * a.stages + 1)
        ring0 = b[rank]
    else:
        wg = wl.ring_oscillator_array(a.rings, a.stages)
        w = pt.partition_workload(wg, world, rank) if a.graph_partition else pt.partition_ring_array(wg, world, rank)
    eng = wl.build_engine(w, device=local)
    ids = [Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng.comm_init(ids[0], rank, world)
    if a.p2p:
        hs = [None] * world
        dist.all_gather_object(hs, eng.p2p_handle())
        eng.p2p_attach(hs)
    eng.border_set(w["n_shared"])
    S = a.stages
    if a.direct:
        my_rings = np.arange(b[rank], b[rank + 1])
        sample = my_rings[np.linspace(0, len(my_rings) - 1, min(a.check_oracle, len(my_rings))).astype(int)] if a.check_oracle else []
        probe_glob = None
        probes = np.concatenate([(r - ring0) * S + np.arange(S) for r in sample] + [[w["vdd"], w["branch"]]]).astype(np.int32)
    else:
        mine = np.where(w["owner"] == rank)[0]
        my_rings = np.unique(mine[mine < a.rings * S] // S)
        sample = my_rings[np.linspace(0, len(my_rings) - 1, min(a.check_oracle, len(my_rings))).astype(int)] if a.check_oracle else []
        loc = np.full(wg["n_unknowns"], -1); loc[w["glob_of_local"]] = np.arange(w["n_unknowns"])
        probe_glob = np.concatenate([r * S + np.arange(S) for r in sample] + [[wg["vdd"], wg["branch"]]]).astype(np.int64)
        probes = loc[probe_glob].astype(np.int32)
        assert np.all(probes >= 0)

    def run():
        eng.set_state(0, w["store"]); eng.set_state(1, w["store"]); eng.b4_set_von(0, w["von"])
        dist.barrier()
        t0 = time.perf_counter()
        r = eng.tran_run(w["x"], a.tstop, 1e-12, probes)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return r, float(t)
    r, t_first = run()
    walls = []
    for _ in range(a.repeat):
        r, t = run(); walls.append(t)
    assert r["rc"] == 0, r.get("error")
    s = r["stats"]
    acc = r["steps"][r["steps"][:, 4] > 0]
    report = {"n_gpus": world, "small_collectives": "peer mailboxes" if a.p2p else "nccl", "p2p_error": eng.p2p_error() if a.p2p else 0, "mosfets_total": int(wg["n_inst"]), "unknowns_total": int(wg["n_unknowns"]), "border_unknowns": int(w["n_shared"]),
              "tstop": a.tstop, "accepted_steps": s["accepted"], "rejected_steps": s["attempts"] - s["accepted"], "newton_iters": s["newton_iters"],
              "wall_s_first_incl_analysis": t_first, "wall_s": min(walls) if walls else t_first,
              "ms_per_newton_iter": 1e3 * (min(walls) if walls else t_first) / max(s["newton_iters"], 1), "lu_analyses": s["lu_analyses"]}
    # ---- the same circuit on one GPU ----
    if a.check_single and not a.direct:
        worst = 0.0
        if rank == 0:
            e1 = wl.build_engine(wg, device=local)
            allp = probe_glob.astype(np.int32)
            r1 = e1.tran_run(wg["x"], a.tstop, 1e-12, allp)
            e1.close()
            # Newton iterations and integration order per attempt identical, accept / reject pattern identical (the positive
            # return codes 1 "norm too small" and 2 "normal convergence" may swap when ||RHS||_2 sits at machine epsilon:
            # the distributed sum of squares adds the ranks' parts in another order)
            same_steps = (r1["steps"].shape == r["steps"].shape and np.array_equal(r1["steps"][:, 2:4], r["steps"][:, 2:4])
                          and np.array_equal(np.sign(r1["steps"][:, 4]), np.sign(r["steps"][:, 4]))
                          and np.array_equal(r1["steps"][:, 4] == -100, r["steps"][:, 4] == -100)
                          and np.allclose(r1["steps"][:, :2], r["steps"][:, :2], rtol=1e-9, atol=0))
            if not same_steps:
                report["debug_steps"] = {"single": r1["steps"][:12].tolist(), "multi": r["steps"][:12].tolist(), "shapes": [list(r1["steps"].shape), list(r["steps"].shape)]}
            worst = float(np.max(np.abs(r1["wave"] - r["wave"]))) if r1["wave"].shape == r["wave"].shape else float("inf")
            report["single_gpu"] = {"identical_step_sequence_and_newton_counts": bool(same_steps), "max_abs_waveform_diff": worst,
                                    "newton_iters": r1["stats"]["newton_iters"]}
    # ---- sampled rings of every rank against the single-ring oracle on the run's own steps ----
    if a.check_oracle and len(sample):
        import oracle_ref
        from b4_common import ref_circuit_from_workload
        h, order, iters = acc[:, 1], acc[:, 3].astype(np.int32), acc[:, 2]
        worst, same, counts_equal, extra = 0.0, True, True, 0
        for j, rr in enumerate(sample):
            w1 = wl.ring_oscillator_array(1, S, shifts=[int(wg["shift"][rr])])
            ref = ref_circuit_from_workload(oracle_ref.RefCircuit, w1); ref.set_flags(transient=1)
            want = ref.tran_run(w1["x"], a.tstop, 1e-12, np.arange(S), w1["linear"], w1["sources"], replay=(h, order))
            counts_equal = bool(np.array_equal(want["steps"][:, 2], iters))
            extra = int(np.sum(iters) - np.sum(want["steps"][:, 2]))        # Newton iterations the array took beyond the lone ring
            gw = r["wave"][:, j * S:(j + 1) * S]
            tol = 1e-3 * np.maximum(np.abs(want["wave"]), np.abs(gw)) + 1e-6
            same = same and bool(np.all(np.abs(gw - want["wave"]) <= tol))
            worst = max(worst, float(np.max(np.abs(gw - want["wave"]))))
        res = [None] * world
        dist.all_gather_object(res, {"rank": rank, "rings": [int(x) for x in sample], "waveforms_within_reltol_abstol": same,
                                     "newton_counts_equal": counts_equal, "newton_iters_array_minus_lone_ring": extra, "max_abs_dv": worst})
        report["oracle_single_ring_replay"] = res
    if rank == 0:
        print(json.dumps(report), flush=True)
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
