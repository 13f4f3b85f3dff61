mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Netlist | tail -3) 2>&1 | tee gpurun_out/pytest_gpu.log
XYCE_B200_BENCH_VERBOSE=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; grep "ms:" gpurun_out/bench_f.err | cut -c1-100
