# The round's final single-GPU run (under gpurun): GPU test suite, smoke, both bench arms, the ncu launch list of the
# bench command, the full capture of the evaluation kernel and the launch list of the C3 .TRAN.
# Multi-GPU numbers: scripts/gpu_multi.sh N under `gpurun --gpus N`.
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Netlist | tail -3) 2>&1 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Netlist | tail -1
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ref.json')); print('reference arm', d['value'], d['cpu_baseline']['cores'], d['tran_c3']['ms_per_newton_iter_whole_array'])"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('ours', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['fastest_variant'], d['roofline']['frac'], d['tran_c3']['ms_per_newton_iter'], d['cpu_baseline']['value'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-tran > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 1 -f -o gpurun_out/b4_eval_100k python scripts/prof_one.py 50000 > gpurun_out/b4_eval_ncu.log 2>&1
