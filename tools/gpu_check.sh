#!/usr/bin/env bash
# One GPU-box pass: parity tests, smoke, bench (both arms), ncu launch list + full capture of the two kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
set -u
TAG="${1:-r01}"
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench ref rc=$?"
python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/${TAG}_bench.json
# launch list of the same command (short): kernel SHARES of the step
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --horizon 40 --e2e-steps 1 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# full capture of the two kernels (small horizon keeps the ~40 replays short)
ncu --set full --clock-control none --import-source on -k regex:'fwd_kernel|tape_kernel|tac_kernel|vjp_kernel|bwd_kernel' -c 6 -f -o $OUT/${TAG}_prof \
    python tools/perf_probe.py --B 4096 --T 20 --lanes 8 --reps 1 --grad-only > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
