#!/bin/bash
# round-2 GPU call AB (1 GPU): up to six trial points per pass — the transparency / bit-exact / compact suites and the bench line
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests/test_gpu_bitexact.py tests/test_gpu_compact.py tests/test_gpu_solver.py -x -q -m gpu -k "not n1e8" ) > gpurun_out/ab_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ab_tests.log
( time timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err
tail -n 6 gpurun_out/ab_tests.log
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/ab_bench.json") if l.startswith("{")][-1])
c=d["compact_direction"]
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "passes", d["iteration"]["line_search_passes_per_iteration"], "| compact", c["value"], c["ms_per_step"], "| config5", d["config5"]["value"], d["config5"]["compact_direction"]["value"], d["config5"]["line_search_passes_per_iteration"])
PY
