#!/usr/bin/env bash
# A/B kernel build variants: tools/gpu_variants.sh <T> <reps> lib1.so lib2.so ...
T=$1; shift; R=$1; shift
for lib in "$@"; do
  echo "== $lib"
  TSIM_B200_LIB=$PWD/$lib python tools/perf_probe.py --B 4096 --T $T --lanes 8 --reps $R --grad-only 2>&1 | grep -o "fwd+tape [0-9.]* ms\|adjoint [0-9.]* ms" | paste -sd' '
done
