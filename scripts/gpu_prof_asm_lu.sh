mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from xyce_b200 import workloads as wl
w = wl.ring_oscillator_array(4950, 101)
eng = wl.build_engine(w)
r = eng.tran_run(w["x"], 4e-12, 1e-12, [0])
print(r["stats"])
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.per_cycle_active,launch__registers_per_thread,sm__warps_active.avg.per_cycle_active --clock-control none -k regex:"gather|lu_|linear_combo|residual_norms_k|spmv" -s 30 -c 40 --csv --log-file gpurun_out/asm_lu_metrics.csv python /tmp/one.py > gpurun_out/asm_lu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/asm_lu_metrics.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= iv: continue
    per.setdefault((r[iid], r[ik].split("(")[0][-45:]), {})[r[im]] = float(r[iv].replace(",", ""))
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for (i, name), m in per.items():
    for k, v in m.items(): agg[name][k].append(v)
print("%-46s %5s %9s %9s %9s %8s %7s" % ("kernel", "n", "time us", "rd MB", "wr MB", "GB/s", "dram%"))
for name, m in agg.items():
    t = sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"]) / 1e3
    rd = sum(m["dram__bytes_read.sum"]) / len(m["dram__bytes_read.sum"]); wr = sum(m["dram__bytes_write.sum"]) / len(m["dram__bytes_write.sum"])
    # ncu reports bytes in scaled units per row; normalise by magnitude heuristics is unsafe -> print raw too
    print("%-46s %5d %9.1f %9.3f %9.3f %8.1f %7.1f" % (name, len(m["gpu__time_duration.sum"]), t, rd, wr, 0.0,
          sum(m["dram__throughput.avg.pct_of_peak_sustained_elapsed"]) / len(m["dram__throughput.avg.pct_of_peak_sustained_elapsed"])))
PY
head -3 gpurun_out/asm_lu_metrics.csv | cut -c1-300
