#!/bin/bash
# round-2 GPU call N (1 GPU): how many ring vectors pass B should keep open per sub-pass (m = 20 at 2^28, m = 12 / 32 at 1e8)
mkdir -p gpurun_out
: > gpurun_out/n_sweep.log
for split in 0 28 20 14 10; do
  LBFGSB200_COMPACT_SPLIT=$split timeout 300 python scripts/tune_compact.py 268435456 20 8 2>&1 | grep compact >> gpurun_out/n_sweep.log
done
for split in 0 20 14; do
  LBFGSB200_COMPACT_SPLIT=$split timeout 300 python scripts/tune_compact.py 100000000 12 8 2>&1 | grep compact >> gpurun_out/n_sweep.log
  LBFGSB200_COMPACT_SPLIT=$split timeout 300 python scripts/tune_compact.py 100000000 32 8 2>&1 | grep compact >> gpurun_out/n_sweep.log
done
timeout 300 python scripts/tune_compact.py 268435456 20 8 2>&1 | grep two_loop >> gpurun_out/n_sweep.log
cat gpurun_out/n_sweep.log
