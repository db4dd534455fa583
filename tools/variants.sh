#!/bin/bash
# development aid: bench every build/libb2d_*.so variant (mixed_262144, device arm only)
for f in build/libb2d_*.so; do
  echo "== $f"
  B2D_LIB=$PWD/$f python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    j=json.loads(t); print('ms/step %.3f  solver %.3f ms  frac %.3f' % (j['ms_per_step'], j['roofline']['kernel_ms'], j['roofline']['frac']))
except Exception as e: print('FAILED', t[-300:])
"
done
