#!/bin/bash
# round-2 GPU call Y (2 GPUs): A/B of the multi-step probe on the sharded path (mailbox exchange carries at most 3 trial points)
mkdir -p gpurun_out
for k in 4 1 4 1; do
  LBFGSB200_MULTI_PROBE_MAX=$k timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2954$k bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/y_bench2_k$k.json 2> gpurun_out/y_bench2.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/y_bench2_k$k.json") if l.startswith("{")][-1])
print("kmax=$k value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "| compact", round(d["compact_direction"]["value"],1), round(d["compact_direction"]["ms_per_step"],3), "| config5", round(d["config5"]["ms_per_step"],2), "compact", round(d["config5"]["compact_direction"]["ms_per_step"],2), d["config5"]["compact_direction"]["allreduces"])
PY
done
