timeout 900 python -m pytest tests/test_gpu_devices.py tests/test_gpu_adaptor.py -x -q 2>&1 | tail -8
