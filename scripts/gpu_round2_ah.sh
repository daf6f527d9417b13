#!/bin/bash
# round-2 GPU call AH (8 GPUs): the secondary configurations sharded, after the warm-up solves were added
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 scripts/bench_configs.py --full --only cfg3,cfg4 ) > gpurun_out/ah_configs_8gpu_full.log 2> gpurun_out/ah_configs.err
grep "^{" gpurun_out/ah_configs_8gpu_full.log | cut -c1-600; tail -n 3 gpurun_out/ah_configs.err
