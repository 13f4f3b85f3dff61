mkdir -p gpurun_out
python -m pytest tests/test_gpu_full_size.py -x -q -k "newton_step_host or load_host_jr" 2>&1 | grep -v Netlist | tail -15
python bench.py --steps 20 --warmup 5 --no-tran > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -3 gpurun_out/r02_bench_d.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_d.json')); print('ours', d['value'], d['ms_per_step'], d['roofline']['kernel_ms']); print(json.dumps(d['e2e'])[:1500])"
