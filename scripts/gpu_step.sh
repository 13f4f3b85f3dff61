mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | grep -v Netlist | tail -4
python scripts/tran_bench.py 2>&1 | grep -v Netlist | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(r['impl'][:20], r['mosfets'], 'iters', r['newton_iters'], 'ms/iter %.3f' % r['ms_per_newton_iter'], 'wall %.3f' % r['wall_s'], 'launches', r.get('launches'))
"
XYCE_B200_BENCH_VERBOSE=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; grep "ms:" gpurun_out/bench_d.err | cut -c1-200; tail -c 1500 gpurun_out/bench_d.json
