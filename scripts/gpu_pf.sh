python scripts/simple_kernels_timing.py 200000 gpurun_out/sk_base.json > /dev/null 2>&1
XYCE_B200_LIB=$PWD/xyce_b200/lib/exp/libxyce_b200_pf.so python scripts/simple_kernels_timing.py 200000 gpurun_out/sk_pf.json > /dev/null 2>&1
python - <<'PY'
import json
a=json.load(open('gpurun_out/sk_base.json')); b=json.load(open('gpurun_out/sk_pf.json'))
for x,y in zip(a,b): print(x['device'], round(x['eval_ms']*1e3,1), round(y['eval_ms']*1e3,1))
PY
