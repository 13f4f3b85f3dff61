# round-2 ncu captures: batched LU kernels and the assembly kernel inside the C3 .TRAN, the generic ADMS kernel
mkdir -p gpurun_out
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from xyce_b200 import workloads as wl
w = wl.ring_oscillator_array(4950, 101)
eng = wl.build_engine(w)
r = eng.tran_run(w["x"], 1e-11, 1e-12, [0])
r = eng.tran_run(w["x"], 1e-11, 1e-12, [0])
print(r["stats"])
PY
ncu --set full --clock-control none -k regex:"lu_.*batched|assemble_kernel" -s 12 -c 4 -f -o gpurun_out/r02_lu_asm_c3 python /tmp/one.py > gpurun_out/r02_lu_asm_c3_ncu.log 2>&1
tail -2 gpurun_out/r02_lu_asm_c3_ncu.log | cut -c1-200
ncu --set full --clock-control none -k regex:adms_gen_kernel -s 8 -c 1 -f -o gpurun_out/r02_adms_gen python scripts/simple_kernels_timing.py 200000 /tmp/sk.json > gpurun_out/r02_adms_gen_ncu.log 2>&1
tail -2 gpurun_out/r02_adms_gen_ncu.log | cut -c1-200
