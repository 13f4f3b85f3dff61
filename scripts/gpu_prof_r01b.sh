set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 1 -o gpurun_out/prof_b4_fast python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out
