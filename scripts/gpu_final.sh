mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Netlist | tail -3) 2>&1 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Netlist | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_g.json 2> gpurun_out/bench_ref_g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ref_g.json')); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g.json')); print('ours', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['tran_c3']['ms_per_newton_iter'], d['cpu_baseline']['value'], d['clocks'])"
wc -l gpurun_out/bench_g.json gpurun_out/bench_ref_g.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file gpurun_out/launches_v6.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-tran > gpurun_out/b_ncu_v6.log 2>&1
