#!/bin/bash
# round-2 GPU call L (1 GPU): compact direction with the pipelined generic pass B (m > 8) — tests, small-n latency, bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_compact.py -x -q -m gpu ) > gpurun_out/l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/l_tests.log
timeout 600 python scripts/diag_small_n.py > gpurun_out/l_small.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
tail -n 8 gpurun_out/l_tests.log; cat gpurun_out/l_small.log; tail -c 1800 gpurun_out/l_bench.json; tail -n 5 gpurun_out/l_bench.err
