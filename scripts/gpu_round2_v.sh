#!/bin/bash
# round-2 GPU call V (1 GPU): pass A tile (U = 2 vs 3) without the commit fusion at m = 6, and at m = 20; commit fusion at U = 3
mkdir -p gpurun_out
: > gpurun_out/v_sweep.log
for rep in 1 2; do
for v in default gram-3; do
  if [ "$v" = default ]; then so=""; else so="$PWD/build/variants/lib_$v.so"; fi
  echo "== $v" >> gpurun_out/v_sweep.log
  LBFGSB200_COMMIT_GRAM=0 LBFGSB200_SO=$so timeout 300 python scripts/tune_compact.py 100000000 6 10 2>&1 | grep compact >> gpurun_out/v_sweep.log
  LBFGSB200_SO=$so timeout 300 python scripts/tune_compact.py 268435456 20 8 2>&1 | grep compact >> gpurun_out/v_sweep.log
done; done
echo "== default, commit fused" >> gpurun_out/v_sweep.log
timeout 300 python scripts/tune_compact.py 100000000 6 10 2>&1 | grep compact >> gpurun_out/v_sweep.log
cat gpurun_out/v_sweep.log
