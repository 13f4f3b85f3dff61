export XYCE_B200_B4_THREADS=256 XYCE_B200_B4_MINBLOCKS=1 XYCE_B200_B4_UNIFORM=1 XYCE_B200_B4_LOCKSTEP=1
ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 1 -o gpurun_out/prof_b4_256x1_ls_full python scripts/prof_one.py 500000 > gpurun_out/p5.log 2>&1
export XYCE_B200_B4_LOCKSTEP=0
ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 1 -o gpurun_out/prof_b4_256x1_fm python scripts/prof_one.py 500000 > gpurun_out/p6.log 2>&1
tail -n 2 gpurun_out/p5.log gpurun_out/p6.log
