"""BASELINE configs[0] on the GPU (development probe, not a bench line): examples/RollingBallExp/test_sim_speed.py
logic -- 350 steps of the rolling-ball scene under the script's action schedule, the 200x200 tactile field read every
5 steps -- (a) one environment through the drop-in Simulation (the script's own FPS figure), (b) a batch of
environments with perturbed actions through BatchedSim."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from tactilesimulation_b200.layout import scene_from_blob  # noqa: E402
from tactilesimulation_b200.redmax import Simulation  # noqa: E402
from tactilesimulation_b200.sim import BatchedSim  # noqa: E402
from tests import rolling_ball as rb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, nargs="+", default=[256, 1024])
    ap.add_argument("--res", type=int, default=200)
    ap.add_argument("--lanes", type=int, default=16)
    ap.add_argument("--noise", type=float, default=0.02, help="std of the per-environment perturbation of the x/y pad force")
    ap.add_argument("--max-newton", type=int, default=0, help="TSIM_OPT_MAX_NEWTON (0 = the reference's cap)")
    a = ap.parse_args()
    g = np.load(os.path.join(ROOT, "tests", "golden", "rollingball_bdf2_s0.npz"))
    ib, db = rb.full_resolution_blob(g["ibuf"], g["dbuf"], a.res)
    T = g["u"].shape[0]
    # (a) the script, one environment
    sim = Simulation(scene_from_blob(ib, db))
    for rep in range(2):
        sim.reset(False)
        torch.cuda.synchronize()
        t0 = time.time()
        for t in range(T):
            sim.set_u(g["u"][t])
            sim.forward(1)
            if t % 5 == 0:
                tac = sim.get_tactile_force_vector()
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(f"compat Simulation, 1 env, {a.res}x{a.res} markers: time elapsed = {dt:.3f} s, FPS = {T / dt:.1f}", flush=True)
    # (b) batched
    core = BatchedSim((ib, db), "cuda:0", lanes=a.lanes)
    if a.max_newton:
        core.set_option(1, a.max_newton)
    dev = core.device
    rows = rb.tactile_rows(T, 5)
    rng = np.random.default_rng(0)
    for B in a.B:
        u = np.tile(g["u"][:, None, :], (1, B, 1))
        u[:, :, :2] += a.noise * rng.normal(size=(1, B, 2))
        ut = torch.tensor(u, device=dev)
        for rep in range(2):
            q = torch.zeros((B, core.ndof_r), dtype=torch.float64, device=dev)
            qd = torch.zeros_like(q)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            out = core.forward(q, qd, ut, T, tac_rows=rows, want_status=True, want_traj=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            st = out["status"]
            print(f"batched B={B} T={T} tactile every 5 steps ({a.res}x{a.res}): {ms:.1f} ms, {B * T / ms * 1e3:.3e} env-steps/s, "
                  f"newton mean {(st & 255).double().mean().item():.2f} max {(st & 255).max().item()} flags {(st >> 16).max().item()} "
                  f"steps with >20 iterations {((st & 255) > 20).sum().item()}", flush=True)
            del out


if __name__ == "__main__":
    main()
