mkdir -p gpurun_out
nvidia-smi -L | head -4
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | grep -v Netlist | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/multi_gpu_tran.py --rings 4950 --check-single 0 --check-oracle 3 2>&1 | grep -v Netlist | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r02_bench_2gpu.err | tail -1 > gpurun_out/r02_bench_2gpu.json; tail -3 gpurun_out/r02_bench_2gpu.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench_2gpu.json')); print('2gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['config']['parallelism'][:60])"
