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
python -m pytest tests/test_gpu_lu.py tests/test_gpu_border.py tests/test_gpu_lu_import.py tests/test_gpu_lu_graph.py tests/test_gpu_tran.py -x -q 2>&1 | grep -v Netlist | tail -6
python -m pytest tests/test_gpu_full_size.py -x -q -k "c3" 2>&1 | grep -v Netlist | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:batched -s 4 -c 4 --csv python /tmp/one.py 2>&1 | grep batched | cut -c60-250
