mkdir -p gpurun_out
python -m pytest tests/test_gpu_bsim4_parity.py -x -q -k "row_wise" 2>&1 | grep -v Netlist | tail -6
python -m pytest tests/test_gpu_multi.py tests/test_gpu_border.py -x -q 2>&1 | grep -v Netlist | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r02_bench_2gpu_v2.err | tail -1 > gpurun_out/r02_bench_2gpu_v2.json
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_2gpu_v2.json')); print('2gpu', d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:900]); print(json.dumps(d.get('tran_c3'))[:400])"
tail -3 gpurun_out/r02_bench_2gpu_v2.err
