#!/usr/bin/env bash
# round-2 final GPU call (1 GPU): smoke, full GPU suite, driver-style bench + reference arm, all workload arms, ncu profiles
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fin_smoke.txt 2>&1; tail -1 gpurun_out/fin_smoke.txt
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/fin_tests.txt 2>&1
tail -3 gpurun_out/fin_tests.txt | cut -c1-200
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/fin_ref.json 2> gpurun_out/fin_ref.err; tail -c 700 gpurun_out/fin_ref.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
for w in push_fwd dclaw insertion stepsim; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/h_bench_$w.json 2> gpurun_out/h_bench_$w.err
done
python tools/make_r02_bench_lines.py | sed -n 3,12p | cut -c1-330
bash tools/gpu_ncu_r02.sh
