mkdir -p gpurun_out
python -m pytest tests/test_gpu_full_size.py tests/test_gpu_devices.py -x -q -k "c3 or models_set" -s 2>&1 | grep -v Netlist | tail -15
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -3 gpurun_out/r02_bench_a.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_a_ref.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_a.json')); print('ours', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']); print(json.dumps(d['tran_c3'])[:1500])
r=json.load(open('gpurun_out/r02_bench_a_ref.json')); print('ref', r['value'], r['ms_per_step'], r['cpu_baseline']['cores']); print(json.dumps(r['tran_c3'])[:600])"
nproc
