#!/bin/bash
# round-2 GPU call AE (1 GPU): the C++ mirror test and the compact suite on the final build
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_cxx_builder.py tests/test_gpu_compact.py -q -m gpu ) > gpurun_out/ae_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ae_tests.log
tail -n 6 gpurun_out/ae_tests.log
