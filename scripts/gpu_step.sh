mkdir -p gpurun_out
python -m pytest tests/test_gpu_bsim4_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python scripts/b4_variants_ls.py 2>&1 | grep "lockstep 0"
