#!/bin/bash
# round-2 GPU call Q (1 GPU): commit fused with pass A — compact tests, fused-ops related tests, bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_compact.py tests/test_gpu_bitexact.py tests/test_cxx_builder.py -x -q -m gpu ) > gpurun_out/q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/q_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
timeout 300 python scripts/tune_compact.py 100000000 6 10 > gpurun_out/q_tune.log 2>&1
LBFGSB200_COMMIT_GRAM=0 timeout 300 python scripts/tune_compact.py 100000000 6 10 >> gpurun_out/q_tune.log 2>&1
timeout 300 python scripts/tune_compact.py 268435456 20 8 >> gpurun_out/q_tune.log 2>&1
tail -n 8 gpurun_out/q_tests.log; cat gpurun_out/q_tune.log
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/q_bench.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "compact", {k:d["compact_direction"].get(k) for k in ("value","ms_per_step","algorithmic_GBps_per_gpu","profile_pass_kernel_GBps","profile_pass_kernel_ms_per_iteration")})
print("config5", d["config5"]["value"], "compact", {k:d["config5"]["compact_direction"].get(k) for k in ("value","ms_per_step")})
PY
tail -n 3 gpurun_out/q_bench.err
