mkdir -p gpurun_out
python -m pytest tests/test_gpu_lu.py -x -q -k "pivot_monitor or batched or ring_arrays" 2>&1 | grep -v Netlist | tail -4
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from xyce_b200 import workloads as wl
w = wl.ring_oscillator_array(4950, 101)
eng = wl.build_engine(w)
r = eng.tran_run(w["x"], 1e-11, 1e-12, [0])
print(r["stats"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 700 --csv --log-file gpurun_out/r02_launches_tran_c3_v2.csv python /tmp/one.py > gpurun_out/r02_tran_ncu_v2.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_launches_tran_c3_v2.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ik, iv = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= iv: continue
    name = r[ik].split("(")[0][-60:]
    agg[name][0] += 1; agg[name][1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
nit = agg[[k for k in agg if "lu_refactor_kernel" in k][0]][0]
print("Newton iterations in window:", nit, " total kernel time per iteration: %.1f us" % (tot / 1e3 / nit))
for k, v in sorted(agg.items(), key=lambda t: -t[1][1]):
    print("%-62s n=%4d per-iter=%8.1f us  %5.1f%%  avg=%8.1f us" % (k, v[0], v[1] / 1e3 / nit, 100 * v[1] / tot, v[1] / 1e3 / v[0]))
PY
python - <<'PY' 2>&1 | grep -v Netlist
import sys, os, time
sys.path.insert(0, os.getcwd())
from xyce_b200 import workloads as wl
w = wl.ring_oscillator_array(4950, 101)
eng = wl.build_engine(w)
r = eng.tran_run(w["x"], 4e-11, 1e-12, [0])
for tstop in (4e-11, 2e-10):
    t0 = time.perf_counter(); r = eng.tran_run(w["x"], tstop, 1e-12, [0]); dt = time.perf_counter() - t0
    s = r["stats"]
    print("tstop %.0e: %d steps, %d Newton iterations, %.3f ms per Newton iteration, %.3f ms per step" % (tstop, s["attempts"], s["newton_iters"], 1e3 * dt / s["newton_iters"], 1e3 * dt / s["attempts"]))
PY
