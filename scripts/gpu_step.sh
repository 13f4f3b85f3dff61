mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_1gpu_b.json 2>gpurun_out/bench_b.err; tail -c 1800 gpurun_out/bench_1gpu_b.json; tail -3 gpurun_out/bench_b.err
