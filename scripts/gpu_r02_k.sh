mkdir -p gpurun_out
true
true
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 scripts/multi_gpu_tran.py --rings 4950 --check-single 0 --check-oracle 2 --direct 1 2>&1 | grep -v Netlist | tail -1 | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-tran 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
