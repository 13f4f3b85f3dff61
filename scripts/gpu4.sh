mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 scripts/multi_gpu_newton.py --rings 40 --check 1 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -3 | tee gpurun_out/mg4_check.txt
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 scripts/multi_gpu_newton.py --rings 49505 --check 0 --iters 5 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -3) 2>&1 | tee gpurun_out/mg4_c4.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
tail -c 1200 gpurun_out/bench_4gpu.json
