mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lu.py tests/test_gpu_lu_graph.py tests/test_gpu_tran.py -x -q 2>&1 | grep -v Netlist | tail -4
python scripts/lu_big_block_timing.py 500 101 2>&1 | grep -v Netlist | tee gpurun_out/lu_big_block.txt
python scripts/lu_big_block_timing.py 2000 101 2>&1 | grep -v Netlist | tee -a gpurun_out/lu_big_block.txt
