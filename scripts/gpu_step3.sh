mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v Netlist | tail -12) 2>&1 | tee gpurun_out/pytest_gpu.log
