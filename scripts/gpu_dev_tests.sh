# device-level GPU parity tests (small devices, adaptors, translated ADMS models)
timeout 900 python -m pytest tests/test_adms_translator.py tests/test_gpu_adaptor.py tests/test_gpu_devices.py -x -q -m gpu 2>&1 | tail -6
