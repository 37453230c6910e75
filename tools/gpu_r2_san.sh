#!/usr/bin/env bash
# compute-sanitizer on the cooperative step loop and the other kernels (small batch)
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_probe.py 2>&1 | tail -25
  echo "== $tool, DClaw 8x6 (variant 16)"
  PCASE=dclaw8x6_episodic_s0 PB=28 PT=8 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_probe.py 2>&1 | tail -25
done > gpurun_out/san.txt 2>&1
cat gpurun_out/san.txt | cut -c1-300
