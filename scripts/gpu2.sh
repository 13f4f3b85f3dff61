mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 900 gpurun_out/bench_2gpu.json; grep -v "Netlist\|OMP_NUM\|\*\*\*" gpurun_out/bench_2gpu.err | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/multi_gpu_newton.py --rings 40 --check 1 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 scripts/multi_gpu_newton.py --rings 4950 --check 0 --iters 5 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 2>/dev/null | tail -c 500
