#!/usr/bin/env bash
# round-2 GPU call L: sub-tile cooperative line-search trials (TS_LS_SUB=4 on the 16-lane variants) A/B + parity of the 16-lane scenes
mkdir -p gpurun_out
V=tactilesimulation_b200/_variants
for lib in $V/lssub1.so tactilesimulation_b200/libtactilesim_b200.so; do
  echo "== $lib"
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --case dclaw8x6_episodic_s0 --B 2048 --T 200 --lanes 16 --reps 2 --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms\|adjoint [0-9.]* ms\|newton mean.*" | paste -sd' '
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --case insertion20x20_episodic_s0 --B 1024 --T 45 --lanes 16 --reps 2 --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms\|adjoint [0-9.]* ms\|newton mean.*" | paste -sd' '
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --case stable_grasp_episodic_s0 --B 1024 --T 50 --lanes 16 --reps 2 --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms\|adjoint [0-9.]* ms\|newton mean.*" | paste -sd' '
done > gpurun_out/l_variants.txt 2>&1
cat gpurun_out/l_variants.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_properties.py tests/test_gpu_perenv.py tests/test_gpu_api.py -m gpu -q -x > gpurun_out/l_tests.txt 2>&1
tail -4 gpurun_out/l_tests.txt | cut -c1-300
