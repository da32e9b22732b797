#!/bin/bash
# 2-GPU check of the data-parallel CUDA-graph modes: eager (0), one all-reduce in the graph (1), early slices in the graph (2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --cpu-baseline 0"
for g in 0 1 2; do
    timeout 55 $T --dp-graph $g > gpurun_out/dp2_g$g.json 2> gpurun_out/dp2_g$g.err
    echo "g=$g rc=$?"
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/dp2_g$g.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d.get("final_loss"), d["e2e"]["value"])
except Exception as ex:
    print("no json", ex)
PY
    tail -3 gpurun_out/dp2_g$g.err | cut -c1-300
done
