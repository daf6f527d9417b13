#!/bin/bash
# round-2 GPU call S (1 GPU): compact direction in one cluster launch (k_compact_small) — tests, latency; commit_gram tile sweep
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_compact.py -x -q -m gpu ) > gpurun_out/s_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s_tests.log
timeout 600 python scripts/diag_small_n.py > gpurun_out/s_small.log 2>&1
bash scripts/gpu_round2_r.sh > /dev/null 2>&1
tail -n 12 gpurun_out/s_tests.log; cat gpurun_out/s_small.log; cat gpurun_out/r_sweep.log
