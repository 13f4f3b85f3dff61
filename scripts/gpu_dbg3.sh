export XYCE_B200_NO_CPU=1
timeout 900 python -m pytest tests/test_gpu_tran.py tests/test_gpu_devices.py -x -q 2>&1 | grep -v Netlist | tail -3
for g in 1 2 3 4 5 6; do python scripts/tran_bench.py 495 2>&1 | grep -v Netlist | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['mosfets'], 'ms/iter', round(d['ms_per_newton_iter'], 4), 'wall', round(d['wall_s'],4), 'setup', round(d['setup_s_inside'],4), 'run', round(d['run_s_inside'],4), 'maxwait', round(d['max_readback_wait_s'],4))
"; done
