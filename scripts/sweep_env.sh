#!/bin/bash
# Sweep launch-shape env knobs of liblbfgsb200.so on the headline workload; prints per-kernel GB/s.
for bps in 2 3 4 6 8; do for st in 1 0; do
  LBFGSB200_BLOCKS_PER_SM=$bps LBFGSB200_STREAMING=$st python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bps=$bps streaming=$st value=%.2f' % d['value'], d['iteration']['profile_pass']['kernel_GBps'])
"
done; done
