python bench.py --no-tran --no-cpu-baseline --inverters 500000 --steps 30 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_1m.json
python -c "
import json; d=json.load(open('gpurun_out/bench_1m.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline'])"
python bench.py --no-tran --steps 50 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_c2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline'], d['cpu_baseline'])"
