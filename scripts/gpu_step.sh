mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lu.py tests/test_gpu_tran.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | grep -v Netlist | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.3e' % r['value'], 'tran_c3', r['tran_c3'])"
bash scripts/gpu_tran_profile.sh 2>&1 | head -12
