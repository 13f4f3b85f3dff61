mkdir -p gpurun_out
python -m pytest tests/test_adms_translator.py -m gpu -x -q 2>&1 | grep -v Netlist | tail -5
python scripts/simple_kernels_timing.py 400000 gpurun_out/r02_simple_kernels.json 2>&1 | grep -v Netlist
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -3 gpurun_out/r02_bench_e.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_e.json')); print('ours', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']); print(json.dumps(d['tran_c3'])[:900])"
