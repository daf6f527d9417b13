#!/bin/bash
# round-2 GPU call K (1 GPU): the compact search direction — parity tests, then the bench line with its block
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_compact.py -x -q -m gpu -s --durations=10 ) > gpurun_out/k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/k_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
tail -n 40 gpurun_out/k_tests.log; tail -c 3000 gpurun_out/k_bench.json; tail -n 5 gpurun_out/k_bench.err
