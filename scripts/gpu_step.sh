mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python scripts/sanitize_shapes.py 2>&1 | grep -v Netlist | tail -3
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Netlist | tail -4
XYCE_B200_BENCH_VERBOSE=1 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; grep "ms:" gpurun_out/bench_e.err | cut -c1-160; tail -c 2400 gpurun_out/bench_e.json
