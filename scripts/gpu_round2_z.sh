#!/bin/bash
# round-2 GPU call Z (1 GPU): occupancy of the FP64-bound multi-step probe (CTAs per SM: 1, 2 = shipped, 3, 4)
mkdir -p gpurun_out
: > gpurun_out/z_sweep.log
for rep in 1 2; do
for cfg in default:2 pm-1:1 pm-1:2 pm-3:3 pm-4:4 default:3 default:4; do
  v=${cfg%%:*}; bps=${cfg##*:}
  if [ "$v" = default ]; then so=""; else so="$PWD/build/variants/lib_$v.so"; fi
  echo "== $v grid $bps x SMs" >> gpurun_out/z_sweep.log
  LBFGSB200_TRIAL_BLOCKS_PER_SM=$bps LBFGSB200_SO=$so timeout 300 python scripts/tune_compact.py 100000000 6 10 2>&1 | grep compact >> gpurun_out/z_sweep.log
done; done
cat gpurun_out/z_sweep.log
