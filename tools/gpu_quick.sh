#!/usr/bin/env bash
# quick GPU pass: parity tests + short bench
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -3
