#!/bin/bash
# round-2 GPU call AG (1 GPU): scripts/bench_configs.py after the warm-up solves were added (reduced sizes)
mkdir -p gpurun_out
timeout 600 python scripts/bench_configs.py > gpurun_out/ag_configs.log 2> gpurun_out/ag_configs.err; echo "rc=$?" >> gpurun_out/ag_configs.log
grep -c "^{" gpurun_out/ag_configs.log; tail -n 4 gpurun_out/ag_configs.log | cut -c1-300; tail -n 3 gpurun_out/ag_configs.err
