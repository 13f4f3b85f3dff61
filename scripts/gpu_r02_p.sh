mkdir -p gpurun_out
python scripts/dbg_case.py igc2_v461__tran_iter1 2>&1 | grep -v Netlist
for tag in base expinl loginl dexpsel helpers; do
  XYCE_B200_LIB=$PWD/xyce_b200/lib/exp/libxyce_b200_$tag.so python scripts/b4_exp_time.py $tag 128x3,128x4 50000,500000 2>&1 | grep -v Netlist | tee -a gpurun_out/r02_b4_exp_p.jsonl
done
