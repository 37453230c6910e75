#!/usr/bin/env bash
# round-2 GPU call C: block-cooperative contact points vs the serial point loop; parity tests on the cooperative build
mkdir -p gpurun_out
V=tactilesimulation_b200/_variants
( bash tools/gpu_variants.sh 200 3 $V/nocoop.so tactilesimulation_b200/libtactilesim_b200.so ) > gpurun_out/c_variants.txt 2>&1
cat gpurun_out/c_variants.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_parity.py tests/test_gpu_fullsize_properties.py tests/test_gpu_reference_callers.py -m gpu -q -x > gpurun_out/c_tests.txt 2>&1
tail -8 gpurun_out/c_tests.txt
