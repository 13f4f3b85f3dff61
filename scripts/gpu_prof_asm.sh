mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-tran > gpurun_out/b_ncu_v4.log 2>&1
grep -v "^==" gpurun_out/launches_v4.csv | cut -d, -f5,15 | cut -c1-120 | tail -12
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 0 -c 1 -o gpurun_out/prof_asm_v4 python scripts/prof_one.py 50000 > gpurun_out/p12.log 2>&1
ncu -i gpurun_out/prof_asm_v4.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; v=rows[2]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__grid_size','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__average_warp_latency_per_inst_issued.ratio','smsp__issue_active.avg.per_cycle_active','lts__t_sector_hit_rate.pct','launch__waves_per_multiprocessor']
for i,k in enumerate(h):
    if k in keys: print(k,u[i],v[i])
" | tee gpurun_out/prof_asm_v4.txt
