mkdir -p gpurun_out
python -m pytest tests/test_gpu_adaptor.py -x -q 2>&1 | grep -v Netlist | tail -8
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -3 gpurun_out/r02_bench_c.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_c.json')); print('ours', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']); t=d['tran_c3']; print(json.dumps(t))"
