set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests/test_gpu_bsim4_parity.py -x -q -m gpu 2>&1 | tail -15
