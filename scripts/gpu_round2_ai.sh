#!/bin/bash
# round-2 GPU call AI (1 GPU): last check of the final tree — the whole -m gpu suite as the driver runs it, smoke(), `python bench.py` with no flags
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/ai_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ai_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/ai_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/ai_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/ai_bench_default.json 2> gpurun_out/ai_bench.err
tail -n 6 gpurun_out/ai_tests.log; cat gpurun_out/ai_smoke.log; cut -c1-330 gpurun_out/ai_bench_default.json; tail -n 4 gpurun_out/ai_bench.err
