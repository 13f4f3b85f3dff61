mkdir -p gpurun_out
python -m pytest tests/test_gpu_bsim4_parity.py -x -q 2>&1 | grep -v Netlist | tail -3
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
python -c "
import json; d=json.load(open('gpurun_out/bench_f.json')); print(d['value'], d['ms_per_step'], d['gpu_launches'], 'e2e', d['e2e']['value'], d['roofline']['kernel_ms'], d['tran_c3']['ms_per_newton_iter'])"
