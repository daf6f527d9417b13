#!/bin/bash
# round-2 GPU call J (1 GPU): the whole -m gpu suite as the driver runs it, smoke(), the default bench line
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=15 ) > gpurun_out/j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/j_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/j_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/j_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
tail -n 30 gpurun_out/j_tests.log; cat gpurun_out/j_smoke.log; cat gpurun_out/j_bench.json
