mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from xyce_b200 import workloads as wl
w = wl.ring_oscillator_array(4950, 101)
eng = wl.build_engine(w)
r = eng.tran_run(w["x"], 6e-12, 1e-12, [0])
print(r["stats"])
PY
python -m pytest tests/test_gpu_lu.py -x -q -k "batched or ring_arrays" 2>&1 | grep -v Netlist | tail -3
ncu --set full --clock-control none --import-source on -k regex:batched -s 4 -c 2 -o gpurun_out/r02_lu_batched python /tmp/one.py > gpurun_out/r02_lu_batched_ncu.log 2>&1
tail -3 gpurun_out/r02_lu_batched_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:batched -s 4 -c 6 --csv python /tmp/one.py 2>&1 | grep batched | cut -c1-250
