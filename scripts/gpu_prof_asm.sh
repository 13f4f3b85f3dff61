mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 0 -c 1 -o gpurun_out/prof_asm_v5 python scripts/prof_one.py 50000 > gpurun_out/p12.log 2>&1
python scripts/ncu_summarize.py gpurun_out/prof_asm_v5.ncu-rep 2>&1 | cut -c1-900 | tee gpurun_out/prof_asm_v5.txt
ncu -i gpurun_out/prof_asm_v5.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; v=rows[2]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','launch__grid_size','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__average_warp_latency_per_inst_issued.ratio','smsp__issue_active.avg.per_cycle_active','lts__t_sector_hit_rate.pct','launch__waves_per_multiprocessor','sm__cycles_active.avg','sm__cycles_elapsed.max','smsp__inst_executed.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']
for i,k in enumerate(h):
    if k in keys: print(k,u[i],v[i])
" | tee -a gpurun_out/prof_asm_v5.txt
ncu -i gpurun_out/prof_asm_v5.ncu-rep --page source --csv 2>/dev/null > gpurun_out/prof_asm_v5_source.csv; wc -l gpurun_out/prof_asm_v5_source.csv
