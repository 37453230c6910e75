#!/usr/bin/env bash
# round-2 GPU call M: word-level cull of large point sets A/B (16-lane scenes), rolling ball probe, then full suite
mkdir -p gpurun_out
V=tactilesimulation_b200/_variants
for lib in $V/nowordcull.so tactilesimulation_b200/libtactilesim_b200.so; do
  echo "== $lib"
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --case dclaw8x6_episodic_s0 --B 2048 --T 200 --lanes 16 --reps 3 --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms" | paste -sd' '
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --case dclaw_episodic_s0 --B 2048 --T 100 --lanes 16 --reps 2 --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms" | paste -sd' '
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --case insertion20x20_episodic_s0 --B 1024 --T 45 --lanes 16 --reps 3 --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms" | paste -sd' '
done > gpurun_out/m_variants.txt 2>&1
cat gpurun_out/m_variants.txt
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/m_tests.txt 2>&1
tail -4 gpurun_out/m_tests.txt | cut -c1-300
