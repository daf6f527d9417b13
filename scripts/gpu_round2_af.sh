#!/bin/bash
# round-2 GPU call AF (1 GPU): BASELINE configs[2] at full size, the solver's share read against the objective at the solution
mkdir -p gpurun_out
timeout 900 python scripts/bench_configs.py --full --only cfg3 > gpurun_out/af_cfg3.log 2> gpurun_out/af_cfg3.err
grep "^{" gpurun_out/af_cfg3.log; tail -n 3 gpurun_out/af_cfg3.err
