#!/bin/bash
# round-2 GPU call C (2 GPUs): N > 1 path (peer-aware probe / commit), GLM hardening tests, bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/c_tests_multi.log 2>&1; echo "rc=$?" >> gpurun_out/c_tests_multi.log
timeout 900 python -m pytest tests/test_gpu_glm.py tests/test_gpu_primitives.py -x -q -m gpu > gpurun_out/c_tests_glm.log 2>&1; echo "rc=$?" >> gpurun_out/c_tests_glm.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c_bench2.json 2> gpurun_out/c_bench2.err
tail -n 6 gpurun_out/c_tests_multi.log gpurun_out/c_tests_glm.log; tail -c 600 gpurun_out/c_bench2.err
