# 2-GPU sanity after single-GPU changes: library multi-GPU tests, C3 .TRAN on 2 ranks, the bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_border.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 scripts/multi_gpu_tran.py --rings 4950 --tstop 2e-10 --direct 1 --check-oracle 2 2>&1 | grep -v Netlist | tail -1 | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 20 --warmup 5 --no-tran 2>gpurun_out/bench2.err | tail -1 > gpurun_out/bench_2gpu_final.json; python -c "
import json
d=json.load(open('gpurun_out/bench_2gpu_final.json')); print('2gpu', d['value'], d['ms_per_step'], d['e2e']['value'])"
