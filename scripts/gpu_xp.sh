timeout 900 python -m pytest tests/test_adms_translator.py tests/test_gpu_adaptor.py tests/test_gpu_devices.py -x -q -m gpu 2>&1 | tail -8
python scripts/simple_kernels_timing.py 200000 gpurun_out/sk_v4.json > /dev/null 2>&1; python -c "
import json
for r in json.load(open('gpurun_out/sk_v4.json')): print(r['device'], round(r['eval_ms']*1e3,1), '%.3g' % r['evals_per_s'])"
