set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/x2_pytest.log
timeout 600 python tools/rolling_ball_probe.py --B 256 1024 2>&1 | tail -12 | tee gpurun_out/x2_rb.log
python tools/ref_config0.py 2>&1 | tail -1 | tee gpurun_out/x2_ref0.log
python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/x2_bench.json
