#!/usr/bin/env bash
# round-2 GPU call G: per-environment parameters + the whole GPU suite + headline bench on the build with PE kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_perenv.py tests/test_gpu_api.py tests/test_replay_export.py -m gpu -q > gpurun_out/g_perenv.txt 2>&1
tail -30 gpurun_out/g_perenv.txt
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/g_tests.txt 2>&1
tail -5 gpurun_out/g_tests.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
python - <<PY
import json
l = json.loads(open("gpurun_out/g_bench.json").read().strip().splitlines()[-1])
print(l["value"], l["ms_per_step"], l["e2e"]["value"], (l["cpu_baseline"] or {}).get("value"), {k: round(v["ms"], 2) for k, v in l["roofline"]["kernels"].items()})
PY
