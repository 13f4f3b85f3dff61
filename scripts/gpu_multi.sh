mkdir -p gpurun_out
N=${1:-8}
for cfg in "4950 2e-10" "49500 1e-10"; do
set -- $cfg
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 scripts/multi_gpu_tran.py --rings $1 --tstop $2 --direct 1 --check-oracle 2 2>&1 | grep -v Netlist | tail -1 | cut -c1-6000 | tee -a gpurun_out/r02_multi_gpu_tran_${N}gpu_v4.txt | cut -c1-420
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 20 --warmup 5 --no-tran 2>gpurun_out/r02_bench_${N}gpu_v4.err | tail -1 > gpurun_out/r02_bench_${N}gpu_v4.json; python -c "
import json,sys
d=json.load(open('gpurun_out/r02_bench_${N}gpu_v4.json')); print('${N}gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['fastest_variant'])"
