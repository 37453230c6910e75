#!/usr/bin/env bash
# round-2 GPU call A: A/B of the round pacing variants, bench line, full GPU test suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/a_smi.txt 2>&1
V=tactilesimulation_b200/_variants
( bash tools/gpu_variants.sh 200 2 $V/base.so $V/k1t.so $V/k2t.so tactilesimulation_b200/libtactilesim_b200.so $V/k4t.so $V/k6t.so $V/k3w.so ) > gpurun_out/a_variants.txt 2>&1
cat gpurun_out/a_variants.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -c 3000 gpurun_out/a_bench.json
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.txt 2>&1
tail -15 gpurun_out/a_tests.txt
