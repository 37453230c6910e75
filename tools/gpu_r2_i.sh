#!/usr/bin/env bash
# round-2 GPU call I: full GPU suite + headline bench on the build with the per-lane stage inputs, then the ncu profiles
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/i_tests.txt 2>&1
tail -6 gpurun_out/i_tests.txt | cut -c1-300
python bench.py --steps 5 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
python - <<PY
import json
l = json.loads(open("gpurun_out/i_bench.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["e2e"]["value"], (l["cpu_baseline"] or {}).get("value"), {k: round(v["ms"], 2) for k, v in l["roofline"]["kernels"].items()})
PY
bash tools/gpu_ncu_r02.sh
