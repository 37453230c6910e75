#!/usr/bin/env bash
# warp-state stall breakdown of the forward kernel at several batch sizes.  Usage: gpu_stalls.sh <tag> B1 B2 ...
set -u
TAG="${1:-st}"; shift
OUT=gpurun_out; mkdir -p $OUT
for B in "$@"; do
  ncu --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section InstructionStats --clock-control none -k regex:'fwd_kernel' -c 1 --csv --page raw --log-file $OUT/${TAG}_B${B}.csv \
      python tools/perf_probe.py --B $B --T 20 --lanes 8 --reps 1 --grad-only > $OUT/${TAG}_B${B}.log 2>&1
  echo "== B=$B rc=$?"
done
