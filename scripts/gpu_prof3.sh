export XYCE_B200_B4_THREADS=256 XYCE_B200_B4_MINBLOCKS=1 XYCE_B200_B4_UNIFORM=1 XYCE_B200_B4_LOCKSTEP=1
ncu --section WarpStateStats --section SchedulerStats --section SpeedOfLight --clock-control none -k regex:b4_eval -s 3 -c 1 -o gpurun_out/prof_b4_256x1_ls python scripts/prof_one.py 500000 > gpurun_out/p3.log 2>&1
export XYCE_B200_B4_THREADS=512
ncu --section WarpStateStats --section SchedulerStats --section SpeedOfLight --clock-control none -k regex:b4_eval -s 3 -c 1 -o gpurun_out/prof_b4_512x1_ls python scripts/prof_one.py 500000 > gpurun_out/p4.log 2>&1
tail -2 gpurun_out/p3.log gpurun_out/p4.log
