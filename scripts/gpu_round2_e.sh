#!/bin/bash
# round-2 GPU call E (1 GPU): speculation + LJ + small-n checks, latency, bench, ncu of the small / LJ kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_bitexact.py tests/test_gpu_primitives.py tests/test_gpu_solver.py -x -q -m gpu -k "not n1e8" > gpurun_out/e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/e_tests.log
timeout 300 python scripts/diag_small_n.py > gpurun_out/e_small.log 2>&1
timeout 300 python scripts/prof_lj.py > gpurun_out/e_lj.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
LJ_REPS=1 python scripts/prof_lj.py > gpurun_out/e_plain_lj.log 2>&1 &&
LJ_REPS=1 ncu --set full --clock-control none --import-source on -k regex:k_lj_lanes -c 4 -o gpurun_out/e_prof_lj python scripts/prof_lj.py > gpurun_out/e_ncu3.log 2>&1
python scripts/prof_small.py > gpurun_out/e_plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_two_loop_small -s 8 -c 2 -o gpurun_out/e_prof_small python scripts/prof_small.py > gpurun_out/e_ncu4.log 2>&1
tail -n 4 gpurun_out/e_tests.log; cat gpurun_out/e_small.log gpurun_out/e_lj.log
