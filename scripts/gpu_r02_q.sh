mkdir -p gpurun_out
for tag in base ls8 ls30 ls80; do
  XYCE_B200_LIB=$PWD/xyce_b200/lib/exp/libxyce_b200_$tag.so python scripts/b4_exp_time.py $tag 128x3,128x4,384x1,512x1,256x1 50000,500000 2>&1 | grep -v Netlist | tee -a gpurun_out/r02_b4_exp_q.jsonl
done
