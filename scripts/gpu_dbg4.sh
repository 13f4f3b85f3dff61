timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_lu_graph.py tests/test_gpu_lu_import.py tests/test_gpu_tran.py tests/test_gpu_full_size.py -x -q 2>&1 | grep -v Netlist | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tran_c3', d['tran_c3']['ms_per_newton_iter'], d['tran_c3']['wall_s_all_runs'])"
bash scripts/gpu_tran_profile.sh 2>&1 | grep "Newton iterations in window\|lu_\|tstop" | cut -c1-150
