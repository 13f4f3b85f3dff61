mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 scripts/multi_gpu_tran.py --rings 40 --stages 31 --tstop 3e-10 2>&1 | grep -v Netlist | tail -2 | cut -c1-3000
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | grep -v Netlist | tail -4
