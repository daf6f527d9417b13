#!/bin/bash
# round-2 GPU call M (1 GPU): steady-state (ring full) m = 20 at 2^28: two-loop vs compact, unrolled vs pipelined pass B
mkdir -p gpurun_out
for mode in 0 1; do
  LBFGSB200_COMPACT_PIPELINED=$mode timeout 600 python bench.py --n 268435456 --m 20 --steps 10 --warmup 22 --no-config5 --no-cpu-baseline > gpurun_out/m_bench_m20_pipelined$mode.json 2> gpurun_out/m_bench.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/m_bench_m20_pipelined$mode.json").read().strip().splitlines()[-1])
print("pipelined=$mode two-loop ms/it", d["ms_per_step"], "compact", {k:d["compact_direction"][k] for k in ("ms_per_step","algorithmic_GBps_per_gpu","profile_pass_kernel_GBps","profile_pass_kernel_ms_per_iteration")})
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-config5 --no-cpu-baseline > gpurun_out/m_bench_m6.json 2>> gpurun_out/m_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/m_bench_m6.json").read().strip().splitlines()[-1])
print("m=6 two-loop", d["value"], "compact", {k:d["compact_direction"][k] for k in ("value","ms_per_step","algorithmic_GBps_per_gpu","profile_pass_kernel_GBps","profile_pass_kernel_ms_per_iteration")})
PY
tail -n 3 gpurun_out/m_bench.err
