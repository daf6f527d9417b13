#!/bin/bash
# round-2 GPU call G (8 GPUs): the driver's multi-GPU bench line, and the sharded secondary configs at full size
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/g_bench8.json 2> gpurun_out/g_bench8.err
timeout 900 $TR scripts/bench_configs.py --full --only cfg4,cfg3 > gpurun_out/g_cfg_8gpu.log 2> gpurun_out/g_cfg_8gpu.err
nvidia-smi topo -m > gpurun_out/g_topo.txt 2>&1
tail -c 1500 gpurun_out/g_bench8.json; tail -c 400 gpurun_out/g_bench8.err; cut -c1-300 gpurun_out/g_cfg_8gpu.log; tail -c 400 gpurun_out/g_cfg_8gpu.err
