#!/usr/bin/env bash
# round-2 GPU call R: block-wide line-search service (-DTS_LS_SERVICE=1) A/B on the 16-lane scenes + parity
mkdir -p gpurun_out
SVC=$PWD/tactilesimulation_b200/_variants/svc.so
for lib in ${LIBS:-"" "$SVC"}; do
  echo "== lib ${lib:-stock}"
  for c in "dclaw8x6_episodic_s0 2048 200" "insertion20x20_episodic_s0 1024 45" "stable_grasp_episodic_s0 1024 100"; do
    set -- $c
    echo "-- $1 B=$2 T=$3"
    TSIM_B200_LIB=$lib timeout 900 python tools/perf_probe.py --case $1 --B $2 --T $3 --lanes 16 --reps 2 2>&1 | tail -2
  done
done 2>&1 | tee gpurun_out/r_ab.txt
TSIM_B200_LIB=$SVC timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r_tests.txt 2>&1
tail -5 gpurun_out/r_tests.txt | cut -c1-300
