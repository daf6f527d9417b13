#!/bin/bash
# round-2 GPU call D (1 GPU): whole -m gpu suite, bench, secondary configs, small-n latency, ncu captures
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/d_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
timeout 300 python scripts/diag_small_n.py > gpurun_out/d_small.log 2>&1
timeout 600 python scripts/prof_lj.py > gpurun_out/d_lj.log 2>&1
timeout 900 python scripts/bench_configs.py --only cfg1,cfg4 > gpurun_out/d_cfg_small.log 2>&1
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config5"
$BENCH > gpurun_out/d_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d_launches_bench.csv $BENCH > gpurun_out/d_ncu1.log 2>&1
$BENCH > gpurun_out/d_plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_rosenbrock_probe|k_rosenbrock_commit|k_backward" -s 40 -c 14 -o gpurun_out/d_prof_hot $BENCH > gpurun_out/d_ncu2.log 2>&1
LJ_REPS=1 python scripts/prof_lj.py > gpurun_out/d_plain_lj.log 2>&1 &&
LJ_REPS=1 ncu --set full --clock-control none --import-source on -k regex:k_lj_lanes -c 4 -o gpurun_out/d_prof_lj python scripts/prof_lj.py > gpurun_out/d_ncu3.log 2>&1
python scripts/prof_small.py > gpurun_out/d_plain_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_two_loop_small -s 8 -c 2 -o gpurun_out/d_prof_small python scripts/prof_small.py > gpurun_out/d_ncu4.log 2>&1
tail -n 4 gpurun_out/d_tests.log; cat gpurun_out/d_small.log gpurun_out/d_lj.log; ls -la gpurun_out/*.ncu-rep
