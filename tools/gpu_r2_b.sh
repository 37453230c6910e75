#!/usr/bin/env bash
# round-2 GPU call B: A/B of the side-by-side contact point evaluation (TS_GP_ILP), then the full GPU test suite
mkdir -p gpurun_out
V=tactilesimulation_b200/_variants
( bash tools/gpu_variants.sh 200 2 $V/ilp1.so tactilesimulation_b200/libtactilesim_b200.so $V/ilp3.so $V/ilp4.so ) > gpurun_out/b_variants.txt 2>&1
cat gpurun_out/b_variants.txt
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/b_tests.txt 2>&1
tail -25 gpurun_out/b_tests.txt
