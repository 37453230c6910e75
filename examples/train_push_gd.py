#!/usr/bin/env python
"""gd.py-style analytic policy-gradient training of TactilePush on the B200 path (R/algorithms/gd.py:133-264,
R/examples/TactilePushExp/cfg/gd_tactile.yaml): batched rollouts through BatchedTactilePushEnv, loss = -mean episode
reward, one backward through the simulator, ONE all-reduce of the policy gradient across ranks, clip, Adam with
linear LR decay.

    python examples/train_push_gd.py --scene <pusher.xml | golden .npz> --batch 1024 --epochs 20
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_push_gd.py --batch 4096 ...
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

from tactilesimulation_b200.distributed import allreduce_gradients, rank_seed  # noqa: E402
from tactilesimulation_b200.envs import BatchedTactilePushEnv  # noqa: E402
from tactilesimulation_b200.layout import scene_from_blob  # noqa: E402
from tactilesimulation_b200.redmax import Simulation  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default=os.path.join(ROOT, "tests", "golden", "pusher13x10_episodic_s0.npz"))
    ap.add_argument("--batch", type=int, default=1024, help="environments per GPU (= episodes per epoch per GPU)")
    ap.add_argument("--epochs", type=int, default=10)
    ap.add_argument("--horizon", type=int, default=100, help="gym steps per episode (R/envs/__init__.py:9-13)")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--max-grad-norm", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if a.scene.endswith(".npz"):
        g = np.load(a.scene)
        scene, q_init = scene_from_blob(g["ibuf"], g["dbuf"]), g["q0"]
    else:
        from tactilesimulation_b200.scene import compile_scene
        scene, q_init = compile_scene(a.scene), None
    sim = Simulation(scene, batch=a.batch, device=dev)
    if q_init is not None:
        sim.set_q_init(np.tile(q_init, (a.batch, 1)))
    M = sim.ndof_tactile // 3
    rows = 13 if M == 130 else 32
    env = BatchedTactilePushEnv(sim, "tactile_flatten", gradient=True, tactile_rows=rows, tactile_cols=M // rows,
                                seed=rank_seed(a.seed, rank))
    torch.manual_seed(a.seed)            # identical policy replicas
    actor = torch.nn.Sequential(torch.nn.Linear(3 + 3 * M, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                torch.nn.Linear(64, 3)).double().to(dev)
    opt = torch.optim.Adam(actor.parameters(), lr=a.lr)
    for epoch in range(a.epochs):
        for gparam in opt.param_groups:
            gparam["lr"] = a.lr * (1.0 - epoch / float(a.epochs))            # gd.py linear decay
        t0 = time.time()
        obs, total = env.reset(), 0.0
        for _ in range(a.horizon):
            obs, r, done, info = env.step(actor(obs))
            total = total + r
        loss = -total.sum()                                                   # 1/(global batch) applied after the reduce
        opt.zero_grad(set_to_none=True)
        loss.backward()
        stats = allreduce_gradients(actor.parameters(), global_batch=a.batch * world,
                                    extra=torch.stack([total.detach().sum(), info["final_pos_error"].sum()]))
        torch.nn.utils.clip_grad_norm_(actor.parameters(), a.max_grad_norm)
        opt.step()
        sim.clearBackwardCache()
        torch.cuda.synchronize()
        if rank == 0:
            n = a.batch * world
            print(f"epoch {epoch}: mean episode reward {stats[0].item() / n:.3f}  final pos error {stats[1].item() / n:.4f}  "
                  f"{n * a.horizon * env.frame_skip / (time.time() - t0):.3e} env-steps/s", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
