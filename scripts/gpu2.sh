mkdir -p gpurun_out
python -m pytest tests/test_gpu_devices.py -x -q -m gpu 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 1500 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 2>/dev/null | tail -c 600
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_1gpu.json 2>/dev/null; tail -c 2500 gpurun_out/bench_1gpu.json
