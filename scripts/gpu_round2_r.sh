#!/bin/bash
# round-2 GPU call R (1 GPU): tile shape of the commit fused with pass A (default U=2, 1 CTA/SM vs U=1, 2 CTAs/SM vs U=3, 1 CTA/SM)
mkdir -p gpurun_out
: > gpurun_out/r_sweep.log
for rep in 1 2; do
for v in default cg-1-2 cg-3-1; do
  if [ "$v" = default ]; then so=""; else so="$PWD/build/variants/lib_$v.so"; fi
  echo "== $v" >> gpurun_out/r_sweep.log
  LBFGSB200_SO=$so timeout 300 python scripts/tune_compact.py 100000000 6 10 2>&1 | grep compact >> gpurun_out/r_sweep.log
done; done
cat gpurun_out/r_sweep.log
