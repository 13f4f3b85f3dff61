for v in "" asm3 asm4; do
  if [ -n "$v" ]; then export XYCE_B200_LIB=$PWD/xyce_b200/lib/exp/libxyce_b200_$v.so; else unset XYCE_B200_LIB; fi
  for r in 1 2; do python bench.py --no-tran --no-cpu-baseline --steps 50 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['value'], d['e2e']['value'])"; done
done
