mkdir -p gpurun_out
python -m pytest tests/test_gpu_fastmath.py tests/test_gpu_bsim4_parity.py -x -q 2>&1 | grep -v Netlist | tail -4
python scripts/b4_exp_time.py default 128x3,128x4 50000,500000 2>&1 | grep -v Netlist
ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 1 -f -o gpurun_out/r02_b4_eval_v4_100k python scripts/prof_one.py 50000 > gpurun_out/r02_b4_eval_v4_ncu.log 2>&1
tail -2 gpurun_out/r02_b4_eval_v4_ncu.log
