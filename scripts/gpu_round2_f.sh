#!/bin/bash
# round-2 GPU call F (1 GPU): the rewritten cluster kernel + speculation, the whole suite once more, latency
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f_tests.log
timeout 300 python scripts/diag_small_n.py > gpurun_out/f_small.log 2>&1
python scripts/prof_small.py > gpurun_out/f_plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_two_loop_small -s 8 -c 2 -o gpurun_out/f_prof_small python scripts/prof_small.py > gpurun_out/f_ncu4.log 2>&1
tail -n 6 gpurun_out/f_tests.log; cat gpurun_out/f_small.log
