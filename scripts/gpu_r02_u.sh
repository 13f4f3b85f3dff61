mkdir -p gpurun_out
python -m pytest tests/test_adms_translator.py tests/test_gpu_devices.py -m gpu -x -q 2>&1 | grep -v Netlist | tail -5
python scripts/simple_kernels_timing.py 400000 gpurun_out/r02_simple_kernels.json 2>&1 | grep -v Netlist | grep adms
