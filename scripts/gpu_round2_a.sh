#!/bin/bash
# round-2 GPU call A: new parity tests, the whole -m gpu suite, bench, tile sweep for probe / commit
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_smi.txt; free -g >> gpurun_out/a_smi.txt; nproc >> gpurun_out/a_smi.txt
timeout 1500 python -m pytest tests/test_gpu_fused_steps.py tests/test_gpu_bitexact.py -x -q -m gpu > gpurun_out/a_tests_new.log 2>&1; echo "rc=$?" >> gpurun_out/a_tests_new.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_fused_steps.py --deselect tests/test_gpu_bitexact.py > gpurun_out/a_tests_rest.log 2>&1; echo "rc=$?" >> gpurun_out/a_tests_rest.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench2.json 2>> gpurun_out/a_bench.err
timeout 900 bash scripts/sweep_tune.sh "default:1 x-8-4-8:1 x-8-3-8:1 x-8-2-4:1 x-8-6-12:1 x-8-5-10:1" 2 > gpurun_out/a_sweep.log 2>&1
tail -3 gpurun_out/a_tests_new.log gpurun_out/a_tests_rest.log; cat gpurun_out/a_sweep.log
