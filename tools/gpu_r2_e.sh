#!/usr/bin/env bash
# round-2 GPU call E: cooperative step loop (lookup table + pipelined sums): timing, clock64 profile, parity, all bench arms
mkdir -p gpurun_out
V=tactilesimulation_b200/_variants
( bash tools/gpu_variants.sh 200 3 tactilesimulation_b200/libtactilesim_b200.so ) > gpurun_out/e_variants.txt 2>&1
cat gpurun_out/e_variants.txt
( TSIM_B200_LIB=$PWD/$V/prof.so PT=100 python tools/cycle_profile.py 2>&1 | tail -24 ) > gpurun_out/e_profile.txt 2>&1
cat gpurun_out/e_profile.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py tests/test_gpu_fullsize_properties.py tests/test_gpu_reference_callers.py -m gpu -q -x > gpurun_out/e_tests.txt 2>&1
tail -4 gpurun_out/e_tests.txt
for w in push push_fwd dclaw insertion stepsim; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/e_bench_$w.json 2> gpurun_out/e_bench_$w.err
  echo "== $w: $(python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/e_bench_$w.json").read().strip().splitlines()[-1])
    print(l["value"], l["unit"], "ms/step", l["ms_per_step"], "e2e", l["e2e"]["value"], "cpu", (l.get("cpu_baseline") or {}).get("value"), {k: round(v["ms"], 2) for k, v in l.get("roofline", {}).get("kernels", {}).items()})
except Exception as e:
    print("FAILED", e)
PY
)"
  tail -3 gpurun_out/e_bench_$w.err
done
