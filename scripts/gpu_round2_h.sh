#!/bin/bash
# quick 1-GPU check of the cluster kernel variant: parity tests + latency
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver.py tests/test_gpu_bitexact.py -x -q -m gpu -k "small or speculative or p2 or p3 or owlqn or damping" > gpurun_out/h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/h_tests.log
timeout 300 python scripts/diag_small_n.py > gpurun_out/h_small.log 2>&1
tail -n 4 gpurun_out/h_tests.log; cat gpurun_out/h_small.log
