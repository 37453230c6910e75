#!/usr/bin/env bash
# ncu full capture of the forward (and adjoint) kernel at the bench batch size, short horizon.  Usage: gpu_ncu.sh <tag> [T]
set -u
TAG="${1:-prof}"; T="${2:-20}"
OUT=gpurun_out; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:'fwd_kernel|bwd_kernel' -c 2 -f -o $OUT/${TAG}_prof \
    python tools/perf_probe.py --B 4096 --T $T --lanes 8 --reps 1 --grad-only > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -2 $OUT/${TAG}_ncu_full.log
