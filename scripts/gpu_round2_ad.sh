#!/bin/bash
# round-2 GPU call AD (4 GPUs): the C++ mirror test on the final build, then the 4-GPU bench line
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_cxx_builder.py -q -m gpu ) > gpurun_out/ad_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ad_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 20 --warmup 3 ) > gpurun_out/ad_bench4.json 2> gpurun_out/ad_bench4.err
tail -n 3 gpurun_out/ad_tests.log
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/ad_bench4.json") if l.startswith("{")][-1])
c=d["compact_direction"]
print("4 GPUs value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "| compact", c["value"], c["ms_per_step"], (c.get("e2e") or {}).get("value"), "| config5", d["config5"]["value"], d["config5"]["compact_direction"]["value"], d["parity"]["timed_trajectory"]["bar_met"], d["parity"]["nondegenerate_sharded_solve"]["bar_met"])
PY
tail -n 3 gpurun_out/ad_bench4.err
