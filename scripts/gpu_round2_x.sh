#!/bin/bash
# round-2 GPU call X (1 GPU): the whole -m gpu suite after the multi-step probe landed
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -q -m gpu --durations=6 ) > gpurun_out/t_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t_tests.log
tail -n 14 gpurun_out/t_tests.log
