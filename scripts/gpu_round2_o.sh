#!/bin/bash
# round-2 GPU call O (2 GPUs): the N > 1 tests (with the compact-direction cases) and the 2-GPU bench line
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu ) > gpurun_out/o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/o_tests.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 ) > gpurun_out/o_bench2.json 2> gpurun_out/o_bench2.err
tail -n 12 gpurun_out/o_tests.log
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/o_bench2.json") if l.startswith("{")][-1])
print("2 GPUs: value", d["value"], "e2e", d["e2e"]["value"], "compact", {k:d["compact_direction"].get(k) for k in ("value","ms_per_step","allreduces","parity")})
print("config5", d["config5"]["value"], "compact", {k:d["config5"]["compact_direction"].get(k) for k in ("value","ms_per_step","parity")})
print(d["parity"])
PY
tail -n 5 gpurun_out/o_bench2.err
