mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Netlist | tail -3) 2>&1 | tee gpurun_out/pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_asm_tran.py 2>&1 | grep -v Netlist | tail -2
python scripts/tran_bench.py 2>&1 | grep -v Netlist | cut -c1-260
python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['tran_c3']['ms_per_newton_iter'], d['tran_c3']['wall_s_all_runs'], d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-tran 2>/dev/null | head -c 300
