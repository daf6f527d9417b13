#!/bin/bash
# round-2 GPU call B: small-n path + perturbed solves, bench in its new format, small-n latency, probe variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_solver.py tests/test_gpu_bitexact.py tests/test_gpu_fused_steps.py -x -q -m gpu -k "not init_direction and not backward_step and not forward_step and not damp_y and not history_update" > gpurun_out/b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/b_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
timeout 300 python scripts/diag_small_n.py > gpurun_out/b_small.log 2>&1
timeout 900 bash scripts/sweep_tune.sh "default:1 x-8-4-8:1 x-8-4-10:1 p-8-4-8:1 p-8-4-6:1 p-8-4-4:1 p-8-4-10:1" 2 > gpurun_out/b_sweep.log 2>&1
export LBFGSB200_TRIAL_BLOCKS_PER_SM=2
timeout 600 bash scripts/sweep_tune.sh "default:1 x-8-4-4:1 x-8-4-5:1 x-8-4-8:1" 1 > gpurun_out/b_sweep_t2.log 2>&1
tail -n 5 gpurun_out/b_tests.log; cat gpurun_out/b_small.log gpurun_out/b_sweep.log gpurun_out/b_sweep_t2.log
