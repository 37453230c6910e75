#!/usr/bin/env bash
# round-2 8-GPU lines for profiles/: headline (configs[2]) and configs[4] (TactileInsertion rollout, 8192 environments over 8 GPUs)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/n8_push.json 2> gpurun_out/n8_push.err
tail -c 600 gpurun_out/n8_push.json; tail -2 gpurun_out/n8_push.err
$TR bench.py --gpus 8 --workload insertion --steps 3 --warmup 3 > gpurun_out/n8_insertion.json 2> gpurun_out/n8_insertion.err
tail -c 600 gpurun_out/n8_insertion.json; tail -2 gpurun_out/n8_insertion.err
$TR bench.py --gpus 8 --workload dclaw --steps 3 --warmup 3 > gpurun_out/n8_dclaw.json 2> gpurun_out/n8_dclaw.err
tail -c 400 gpurun_out/n8_dclaw.json; tail -2 gpurun_out/n8_dclaw.err
