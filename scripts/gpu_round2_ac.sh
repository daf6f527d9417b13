#!/bin/bash
# round-2 GPU call AC (2 GPUs): the N > 1 tests and the 2-GPU bench line of the final build
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu ) > gpurun_out/ac_tests.log 2>&1; echo "rc=$?" >> gpurun_out/ac_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/ac_bench2.json 2> gpurun_out/ac_bench2.err
tail -n 5 gpurun_out/ac_tests.log
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/ac_bench2.json") if l.startswith("{")][-1])
c=d["compact_direction"]
print("2 GPUs value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "| compact", c["value"], c["ms_per_step"], (c.get("e2e") or {}).get("value"), "| config5", d["config5"]["value"], d["config5"]["compact_direction"]["value"], d["parity"]["nondegenerate_sharded_solve"]["bar_met"])
PY
