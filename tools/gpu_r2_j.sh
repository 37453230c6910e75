#!/usr/bin/env bash
# round-2 GPU call J: env front-end tests (insertion, dclaw), the gd.py-style training example (10 epochs, log kept)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_callers.py -m gpu -q > gpurun_out/j_tests.txt 2>&1
grep -n "^E  \|passed\|failed" gpurun_out/j_tests.txt | head -20 | cut -c1-300
timeout 900 python examples/train_push_gd.py --batch 1024 --epochs 10 > gpurun_out/j_train_push_gd.log 2>&1
tail -14 gpurun_out/j_train_push_gd.log
