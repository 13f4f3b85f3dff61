mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Netlist | tail -3) 2>&1 | tee gpurun_out/pytest_gpu.log
XYCE_B200_NO_CPU=1 python scripts/tran_bench.py 495 2>&1 | grep -v Netlist | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['mosfets'], 'ms/iter', round(d['ms_per_newton_iter'], 4), 'wall', round(d['wall_s'],4), 'run', round(d['run_s_inside'],4))
"
python scripts/lu_big_block_timing.py 500 101 2>&1 | grep -v Netlist | cut -c1-200
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
python -c "
import json; d=json.load(open('gpurun_out/bench_f.json')); print(d['value'], d['ms_per_step'], d['gpu_launches'], d['e2e']['value'], d['roofline']['kernel_ms'], d['tran_c3']['ms_per_newton_iter'])"
