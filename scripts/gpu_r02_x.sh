mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | grep -v Netlist | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; tail -2 gpurun_out/r02_bench_f.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_f.json')); print('ours', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['tran_c3']['ms_per_newton_iter'])"
ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 1 -f -o gpurun_out/r02_b4_eval_v5_100k python scripts/prof_one.py 50000 > gpurun_out/r02_b4_eval_v5_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_v4.csv python bench.py --steps 6 --warmup 3 --no-tran --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -1 gpurun_out/r02_b4_eval_v5_ncu.log
