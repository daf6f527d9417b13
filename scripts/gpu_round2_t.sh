#!/bin/bash
# round-2 GPU call T (1 GPU), final build: the whole -m gpu suite as the driver runs it, smoke(), the default bench line,
# the reference arm (short), the ncu launch list of the bench command, the secondary configurations
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu --durations=8 ) > gpurun_out/t_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/t_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/t_smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/t_bench_ref.json 2> gpurun_out/t_bench_ref.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/t_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/t_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/t_ncu.log 2>&1
timeout 900 python scripts/bench_configs.py --full > gpurun_out/t_configs_1gpu_full.log 2> gpurun_out/t_configs.err
tail -n 14 gpurun_out/t_tests.log; cat gpurun_out/t_smoke.log; tail -c 1500 gpurun_out/t_bench.json; cat gpurun_out/t_bench_ref.json; tail -n 3 gpurun_out/t_bench.err gpurun_out/t_configs.err; cat gpurun_out/t_configs_1gpu_full.log
