#!/bin/bash
# round-2 GPU call U (8 GPUs), final build: the bench line at N = 8 (driver's arguments) and the secondary configurations sharded
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 ) > gpurun_out/u_bench8.json 2> gpurun_out/u_bench8.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 scripts/bench_configs.py --full --only cfg3,cfg4 ) > gpurun_out/u_configs_8gpu_full.log 2> gpurun_out/u_configs.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/u_bench8.json") if l.startswith("{")][-1])
print("8 GPUs: value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
c=d["compact_direction"]; print("compact", {k:c.get(k) for k in ("value","ms_per_step","allreduces","parity","e2e")})
print("config5", {k:d["config5"].get(k) for k in ("value","ms_per_step","parity")}, "compact", {k:d["config5"]["compact_direction"].get(k) for k in ("value","ms_per_step","parity")})
print(d["parity"])
PY
tail -n 4 gpurun_out/u_bench8.err; grep "^{" gpurun_out/u_configs_8gpu_full.log
