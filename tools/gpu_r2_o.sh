#!/usr/bin/env bash
# round-2 GPU call O: LU elimination through shared scratch vs shuffles (both with the merged count/vote barrier) + parity
mkdir -p gpurun_out
V=tactilesimulation_b200/_variants
( bash tools/gpu_variants.sh 200 3 $V/lushfl.so tactilesimulation_b200/libtactilesim_b200.so ) > gpurun_out/o_variants.txt 2>&1
cat gpurun_out/o_variants.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_properties.py tests/test_gpu_bench_parity.py tests/test_gpu_perenv.py -m gpu -q -x > gpurun_out/o_tests.txt 2>&1
tail -3 gpurun_out/o_tests.txt | cut -c1-200
