#!/bin/bash
# round-2 GPU call AJ (1 GPU): smoke() and the compact suite on the rebuilt library (header comment change only)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/aj_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/aj_smoke.log
( timeout 600 python -m pytest tests/test_gpu_compact.py tests/test_gpu_primitives.py -x -q -m gpu ) > gpurun_out/aj_tests.log 2>&1; echo "rc=$?" >> gpurun_out/aj_tests.log
cat gpurun_out/aj_smoke.log; tail -n 3 gpurun_out/aj_tests.log
