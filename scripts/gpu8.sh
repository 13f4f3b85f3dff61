mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/multi_gpu_newton.py --rings 40 --check 1 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -3 | tee gpurun_out/mg8_check.txt
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 scripts/multi_gpu_newton.py --rings 49505 --check 0 --iters 5 2>&1 | grep -v "OMP_NUM\|\*\*\*" | tail -3) 2>&1 | tee gpurun_out/mg8_c4.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err
tail -c 1200 gpurun_out/bench_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 scripts/multi_gpu_newton.py --rings 4950 --check 0 --iters 5 2>&1 | grep "^{" | tee gpurun_out/mg8_c3.txt
