set -u
mkdir -p gpurun_out
bash tools/gpu_variants.sh 40 2 tactilesimulation_b200/libtactilesim_b200.so gpurun_variants_nosync.so 2>&1 | tee gpurun_out/x1_variants.log
bash tools/gpu_traffic.sh x1 4096 2>&1 | tee gpurun_out/x1_traffic.log
python tools/round_stats.py 2>&1 | tee gpurun_out/x1_rounds.log
