mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_tran_c3_cur.csv python scripts/prof_tran_c3.py 4e-11 > gpurun_out/tran_under_ncu.log 2>&1
tail -2 gpurun_out/tran_under_ncu.log | cut -c1-300
