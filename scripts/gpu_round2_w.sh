#!/bin/bash
# round-2 GPU call W (1 GPU): several line-search trials per pass (probe_multi) — transparency, the bit-exact and compact suites, bench
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_gpu_bitexact.py tests/test_gpu_compact.py tests/test_cxx_builder.py -x -q -m gpu --durations=5 ) > gpurun_out/w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/w_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
LBFGSB200_MULTI_PROBE_MAX=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/w_bench_k1.json 2>> gpurun_out/w_bench.err
timeout 300 python scripts/tune_compact.py 100000000 6 10 > gpurun_out/w_tune.log 2>&1
tail -n 14 gpurun_out/w_tests.log; cat gpurun_out/w_tune.log
python - <<PY
import json
for f in ("gpurun_out/w_bench.json", "gpurun_out/w_bench_k1.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    c=d["compact_direction"]
    print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "syncs", d["iteration"]["host_syncs"], "parity", d["parity"] and d["parity"]["timed_trajectory"]["bar_met"],
          "| compact", c.get("value"), c.get("ms_per_step"), (c.get("e2e") or {}).get("value"), c.get("profile_pass_kernel_ms_per_iteration"))
    if d.get("config5"): print("   config5", d["config5"]["value"], d["config5"]["ms_per_step"], "compact", d["config5"]["compact_direction"].get("value"))
    print("   kernel ms timed region", d["iteration"]["kernel_ms_timed_region"], d["iteration"]["profile_pass"]["kernel_GBps"])
PY
tail -n 3 gpurun_out/w_bench.err
