python -m pytest tests/test_adms_translator.py tests/test_gpu_adaptor.py -m gpu -x -q 2>&1 | grep -v Netlist | tail -6
python scripts/simple_kernels_timing.py 200000 gpurun_out/r02_simple_kernels_v2.json 2>&1 | grep -v Netlist | grep adms | cut -c1-260
