#!/usr/bin/env bash
# round-2 GPU call F: eager batched line search A/B on the straggler-bound configs, stepsim arm, full GPU tests
mkdir -p gpurun_out
V=tactilesimulation_b200/_variants
for lib in $V/noeager.so tactilesimulation_b200/libtactilesim_b200.so; do
  echo "== $lib"
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --case dclaw8x6_episodic_s0 --B 2048 --T 200 --lanes 16 --reps 2 --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms\|adjoint [0-9.]* ms\|newton mean.*" | paste -sd' '
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --case insertion20x20_episodic_s0 --B 1024 --T 45 --lanes 16 --reps 2 --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms\|adjoint [0-9.]* ms\|newton mean.*" | paste -sd' '
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --B 4096 --T 200 --lanes 8 --reps 3 --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms\|adjoint [0-9.]* ms" | paste -sd' '
done > gpurun_out/f_variants.txt 2>&1
cat gpurun_out/f_variants.txt
timeout 900 python bench.py --workload stepsim --steps 2 --warmup 1 > gpurun_out/f_bench_stepsim.json 2> gpurun_out/f_bench_stepsim.err
tail -c 1500 gpurun_out/f_bench_stepsim.json; tail -3 gpurun_out/f_bench_stepsim.err
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/f_tests.txt 2>&1
tail -6 gpurun_out/f_tests.txt
