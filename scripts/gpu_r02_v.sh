mkdir -p gpurun_out
for tag in base v1 v2 v3; do
  XYCE_B200_LIB=$PWD/xyce_b200/lib/exp/libxyce_b200_$tag.so python scripts/b4_exp_time.py $tag 128x3,128x4 50000,500000 2>&1 | grep -v Netlist | tee -a gpurun_out/r02_b4_exp_v.jsonl
done
