#!/usr/bin/env bash
# round-2 GPU call R: A/B of library builds on the 16-lane scenes (LIBS = "stock" and / or paths of tools/build_variant.py
# libraries) + the whole GPU suite on TESTLIB (default: the stock library)
mkdir -p gpurun_out
for lib in ${LIBS:-stock}; do
  echo "== lib $lib"
  if [ "$lib" = stock ]; then unset TSIM_B200_LIB; else export TSIM_B200_LIB=$lib; fi
  for c in "dclaw8x6_episodic_s0 2048 200" "insertion20x20_episodic_s0 1024 45" "stable_grasp_episodic_s0 1024 100"; do
    set -- $c
    echo "-- $1 B=$2 T=$3"
    timeout 900 python tools/perf_probe.py --case $1 --B $2 --T $3 --lanes 16 --reps 2 2>&1 | tail -2
  done
done 2>&1 | tee gpurun_out/r_ab.txt
unset TSIM_B200_LIB
if [ -n "$TESTLIB" ]; then export TSIM_B200_LIB=$TESTLIB; fi
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r_tests.txt 2>&1
tail -5 gpurun_out/r_tests.txt | cut -c1-300
