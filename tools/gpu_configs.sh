#!/usr/bin/env bash
# Throughput of the other BASELINE configs at their stated sizes (development probes, not bench lines).
# Usage (under gpurun): bash tools/gpu_configs.sh [tag]
set -u
TAG="${1:-cfg}"
OUT=gpurun_out; mkdir -p $OUT
{
echo "== configs[0] RollingBall (test_sim_speed.py logic; reference CPU line first)"
python tools/ref_config0.py 2>&1 | tail -2
timeout 900 python tools/rolling_ball_probe.py --B 1024 --noise 0 2>&1 | tail -4
timeout 900 python tools/rolling_ball_probe.py --B 1024 --noise 0.02 2>&1 | tail -2
echo "== configs[1] TactilePush 32x13 forward-only, B=1024, T=200"
python tools/perf_probe.py --B 1024 --T 200 --lanes 8 --reps 2 2>&1 | tail -2
echo "== configs[3] DClaw B=2048 T=200 fwd+adjoint"
timeout 900 python tools/perf_probe.py --case dclaw_episodic_s0 --B 2048 --T 200 --lanes 16 --reps 2 --grad-only 2>&1 | tail -2
echo "== configs[4] TactileInsertion forward rollout B=8192 T=45"
timeout 900 python tools/perf_probe.py --case insertion_episodic_s0 --B 8192 --T 45 --lanes 16 --reps 2 2>&1 | tail -2
} 2>&1 | tee $OUT/${TAG}_configs.log
