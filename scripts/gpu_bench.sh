set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -3
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
tail -30 gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 2 -o gpurun_out/prof_b4 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
