#!/usr/bin/env bash
# round-2 GPU call H: insertion front-end test, all bench arms on the final build (lines kept for profiles/r02_bench_lines.md)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_callers.py -m gpu -q -x > gpurun_out/h_tests.txt 2>&1
tail -12 gpurun_out/h_tests.txt | cut -c1-300
for w in push push_fwd dclaw insertion stepsim; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/h_bench_$w.json 2> gpurun_out/h_bench_$w.err
  echo "== $w: $(python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/h_bench_$w.json").read().strip().splitlines()[-1])
    print(l["value"], l["unit"], "ms/step", l["ms_per_step"], "e2e", l["e2e"]["value"], "cpu", (l.get("cpu_baseline") or {}).get("value"), {k: round(v["ms"], 2) for k, v in l.get("roofline", {}).get("kernels", {}).items()}, l.get("roofline", {}).get("fp64", {}).get("frac"), l.get("roofline", {}).get("traffic"))
except Exception as e:
    print("FAILED", e)
PY
)"
  tail -2 gpurun_out/h_bench_$w.err | cut -c1-300
done
