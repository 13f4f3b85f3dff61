mkdir -p gpurun_out
python -m pytest tests/test_gpu_lu.py tests/test_gpu_full_size.py -x -q -k "pivot_monitor or load_host_jr" 2>&1 | grep -v Netlist | tail -15
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -3 gpurun_out/r02_bench_b.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_b.json')); print('ours', d['value'], d['ms_per_step'], d['roofline']['kernel_ms']); print(json.dumps(d['e2e'])); t=d['tran_c3']; t.pop('cpu_baseline',None); print(json.dumps(t))"
