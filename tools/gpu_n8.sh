#!/usr/bin/env bash
# 8-GPU bench line (under gpurun --gpus 8).  Usage: bash tools/gpu_n8.sh [tag]
TAG="${1:-n8}"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n8.json 2> gpurun_out/${TAG}_bench_n8.err
python - <<PY
import json
l = json.loads(open("gpurun_out/${TAG}_bench_n8.json").read().strip().splitlines()[-1])
print("n_gpus", l["n_gpus"], "value", l["value"], "ms_per_step", l["ms_per_step"], "e2e", l["e2e"]["value"])
PY
