mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 1 -o gpurun_out/prof_b4_v3_100k python scripts/prof_one.py 50000 > gpurun_out/p9.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 1 -o /tmp/prof_b4_v3_1m python scripts/prof_one.py 500000 > gpurun_out/p10.log 2>&1
XYCE_B200_B4_SPEC=0 ncu --set full --clock-control none --import-source on -k regex:b4_eval -s 3 -c 1 -o /tmp/prof_b4_v3g_1m python scripts/prof_one.py 500000 > gpurun_out/p11.log 2>&1
(echo "== v3 specialised, 100k instances (auto shape)"; python scripts/ncu_summarize.py gpurun_out/prof_b4_v3_100k.ncu-rep; echo "== v3 specialised, 1M instances (auto shape)"; python scripts/ncu_summarize.py /tmp/prof_b4_v3_1m.ncu-rep; echo "== v3 generic (b4_spec=0), 1M instances"; python scripts/ncu_summarize.py /tmp/prof_b4_v3g_1m.ncu-rep) > gpurun_out/prof_v3_summary.txt 2>&1
for f in gpurun_out/prof_b4_v3_100k.ncu-rep /tmp/prof_b4_v3_1m.ncu-rep /tmp/prof_b4_v3g_1m.ncu-rep; do ncu -i $f --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; v=rows[2]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.per_cycle_active','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.per_cycle_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__grid_size','launch__block_size','smsp__sass_inst_executed_op_local_ld.sum','smsp__sass_inst_executed_op_local_st.sum']
for i,k in enumerate(h):
    if k in keys: print(k,u[i],v[i])
print()
"; done >> gpurun_out/prof_v3_summary.txt
rm -f gpurun_out/prof_b4_v3_100k.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-tran > gpurun_out/b_ncu_v3.log 2>&1
cat gpurun_out/prof_v3_summary.txt
