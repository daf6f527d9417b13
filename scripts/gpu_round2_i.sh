#!/bin/bash
# store / load cache-policy sweep on the shipped tiles (the commit writes four streams at 74 % DRAM utilisation)
mkdir -p gpurun_out
timeout 900 bash scripts/sweep_tune.sh "default:1 s-1-0:1 s-1-2:1 s-0-0:1 x-8-4-6:1" 2 > gpurun_out/i_sweep.log 2>&1
cat gpurun_out/i_sweep.log
