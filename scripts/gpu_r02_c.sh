mkdir -p gpurun_out
python -m pytest tests/test_gpu_lu.py tests/test_gpu_lu_import.py tests/test_gpu_lu_graph.py tests/test_gpu_tran.py -x -q 2>&1 | grep -v Netlist | tail -15
python -m pytest tests/test_gpu_full_size.py -x -q -k "c3" -s 2>&1 | grep -v Netlist | tail -5
python scripts/tran_bench.py 2>&1 | grep -v Netlist | tail -12
