#!/usr/bin/env bash
# round-2 GPU call P: cooperative area limited to 22 slots (164 KB shared-memory configuration): timing, parity; and the
# overflow path forced by a 2-slot build (parity + bitwise invariance tests)
mkdir -p gpurun_out
( bash tools/gpu_variants.sh 200 3 tactilesimulation_b200/libtactilesim_b200.so ) > gpurun_out/p_variants.txt 2>&1
cat gpurun_out/p_variants.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_properties.py tests/test_gpu_bench_parity.py -m gpu -q > gpurun_out/p_tests.txt 2>&1
tail -3 gpurun_out/p_tests.txt | cut -c1-200
echo "== 2-slot build (overflow path)"
TSIM_B200_LIB=$PWD/tactilesimulation_b200/_variants/slots2.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_properties.py tests/test_gpu_bench_parity.py -m gpu -q > gpurun_out/p_tests_slots2.txt 2>&1
tail -6 gpurun_out/p_tests_slots2.txt | cut -c1-300
ncu --metrics launch__shared_mem_config_size,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum -k regex:fwd_kernel -c 1 --csv python tools/perf_probe.py --B 4096 --T 20 --lanes 8 --reps 1 --grad-only 2>/dev/null | grep fwd_kernel | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
