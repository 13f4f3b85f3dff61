mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_g.json 2> gpurun_out/bench_ref_g.err; tail -c 700 gpurun_out/bench_ref_g.json; echo
python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -c 300 gpurun_out/bench_g.json; echo; grep -c . gpurun_out/bench_g.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 24 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-tran > gpurun_out/b_ncu_v5.log 2>&1
grep -v "^==" gpurun_out/launches_v5.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,100-140 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Netlist | tail -2
