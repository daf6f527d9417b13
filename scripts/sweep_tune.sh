#!/bin/bash
# usage: sweep_tune.sh "<variant:bps> ..." [repeats]
reps=${2:-2}
for r in $(seq $reps); do
for cfg in $1; do
  v=${cfg%%:*}; bps=${cfg##*:}
  if [ "$v" = default ]; then so=""; else so="$PWD/build/variants/lib_$v.so"; fi
  TUNE_TAG="$v bps=$bps" LBFGSB200_SO=$so LBFGSB200_BLOCKS_PER_SM=$bps python scripts/tune_kernels.py 2>&1 | tail -1
done; done
