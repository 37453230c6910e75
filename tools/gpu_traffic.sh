#!/usr/bin/env bash
# DRAM traffic of the two kernels per env-step at several batch sizes (ncu, T=20).  Usage: gpu_traffic.sh <tag> B1 B2 ...
set -u
TAG="${1:-traffic}"; shift
OUT=gpurun_out; mkdir -p $OUT
for B in "$@"; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:'fwd_kernel|bwd_kernel' -c 2 --csv --log-file $OUT/${TAG}_B${B}.csv \
      python tools/perf_probe.py --B $B --T 20 --lanes 8 --reps 1 --grad-only > $OUT/${TAG}_B${B}.log 2>&1
  echo "== B=$B"; grep -E "fwd_kernel|bwd_kernel" $OUT/${TAG}_B${B}.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}'
done
