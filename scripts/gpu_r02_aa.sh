mkdir -p gpurun_out
python -m pytest tests/test_gpu_full_size.py -x -q -k "pipelined or load_host_jr or newton_step" 2>&1 | grep -v Netlist | tail -8
python bench.py --steps 20 --warmup 5 --no-tran > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; tail -3 gpurun_out/r02_bench_g.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_g.json')); print('ours', d['value'], d['ms_per_step'], d['roofline']['kernel_ms']); print(json.dumps(d['e2e'])[:1800])"
