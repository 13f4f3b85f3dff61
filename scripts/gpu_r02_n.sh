mkdir -p gpurun_out
: > gpurun_out/r02_b4_exp_n.jsonl
for tag in base div1 sq div1sq div1sqes; do
  XYCE_B200_LIB=$PWD/xyce_b200/lib/exp/libxyce_b200_$tag.so python scripts/b4_exp_time.py $tag 128x3,128x4,128x5,128x6,352x1 50000,500000 2>&1 | grep -v Netlist | tee -a gpurun_out/r02_b4_exp_n.jsonl
done
