#!/bin/bash
# Sweep tile-shape builds x CTAs/SM on the headline workload; prints per-kernel GB/s.
for v in default u8b2 u4b2 u6b2 u2b4 u4b3; do
  if [ "$v" = default ]; then so=""; else so="$PWD/build/variants/lib_$v.so"; fi
  for bps in 1 2 3 4; do
    LBFGSB200_SO=$so LBFGSB200_BLOCKS_PER_SM=$bps python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v bps=$bps value=%.2f e2e=%.2f' % (d['value'], d['e2e']['value']), d['iteration']['profile_pass']['kernel_GBps'])
"
  done
done
