#!/usr/bin/env bash
# round-2 GPU call D: clock64 breakdown of the cooperative step loop
mkdir -p gpurun_out
V=tactilesimulation_b200/_variants
for v in prof profgp; do
  echo "== $v"
  TSIM_B200_LIB=$PWD/$V/$v.so PT=100 python tools/cycle_profile.py 2>&1 | tail -22
done > gpurun_out/d_profile.txt 2>&1
cat gpurun_out/d_profile.txt
