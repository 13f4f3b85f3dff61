mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from xyce_b200 import workloads as wl
w = wl.ring_oscillator_array(4950, 101)
eng = wl.build_engine(w)
r = eng.tran_run(w["x"], 1e-11, 1e-12, [0])
r = eng.tran_run(w["x"], 2e-11, 1e-12, [0])
print(r["stats"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 900 --csv --log-file gpurun_out/r02_launches_tran_c3_v3.csv python /tmp/one.py > gpurun_out/r02_tran_ncu_v3.log 2>&1
tail -2 gpurun_out/r02_tran_ncu_v3.log | cut -c1-300
