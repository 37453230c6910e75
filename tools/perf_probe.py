"""Quick throughput probe (development tool): fwd(+tape) and adjoint over T steps at batch B."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from tactilesimulation_b200.sim import BatchedSim  # noqa: E402


def inputs(g, B, T, dev, seed=0):
    rng = np.random.default_rng(seed)
    n, nu = len(g["q0"]), g["u"].shape[1]
    q0 = np.tile(g["q0"], (B, 1))
    if n == 10:      # DClaw rotate-cap (BASELINE configs[3]): U(-1,1)^9 with the middle joints closing on the cap
        u = rng.uniform(-1, 1, (T, B, nu))
        u[:, :, 1::3] = 0.6 + 0.4 * u[:, :, 1::3]
        q0[:, :9] += rng.uniform(-0.05, 0.05, (B, 9))
        return (torch.tensor(q0, device=dev), torch.zeros((B, n), dtype=torch.float64, device=dev), torch.tensor(u, device=dev))
    if n != 7:       # TactileInsertion / StableGrasp: the golden action schedule (cycled), perturbed per environment
        gu = g["u"]
        u = np.stack([gu[t % len(gu)] for t in range(T)])[:, None, :].repeat(B, axis=1)
        u[:, :, :4] += 2e-4 * rng.normal(size=(T, B, 4))
        return (torch.tensor(q0, device=dev), torch.zeros((B, n), dtype=torch.float64, device=dev), torch.tensor(u, device=dev))
    q0[:, 1] = -0.001
    q0[:, 4] = rng.uniform(-0.02, 0.02, B)
    u = np.zeros((T, B, nu))
    u[:, :, :3] = np.tanh(rng.normal(size=(T, B, 3)))
    ext = np.zeros((B, 2))
    for t in range(T):
        if t % 10 == 0:
            ext = np.where(rng.uniform(size=(B, 1)) >= 0.5, rng.uniform(-1, 1, (B, 2)), 0.0)
        u[t, :, 3:5] = ext
    return (torch.tensor(q0, device=dev), torch.zeros((B, n), dtype=torch.float64, device=dev),
            torch.tensor(u, device=dev))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--case", default="pusher32x13_episodic_s0")
    ap.add_argument("--lanes", type=int, nargs="+", default=[8, 16, 32])
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--zero-u", action="store_true", help="zero actions: the pad never touches the box")
    ap.add_argument("--no-gp", action="store_true", help="timing experiment: drop the pad-box contact force")
    ap.add_argument("--grad-only", action="store_true", help="skip the no-grad forward (for ncu captures)")
    ap.add_argument("--max-newton", type=int, default=0, help="TSIM_OPT_MAX_NEWTON (0 = the reference's cap)")
    ap.add_argument("--vjp-pass", type=int, default=1, help="TSIM_OPT_VJP_PASS (1 = readout pull-backs in a balanced pass of their own)")
    ap.add_argument("--tac-pass", type=int, default=1, help="TSIM_OPT_TAC_PASS (1 = tactile readout in a balanced pass of its own)")
    ap.add_argument("--tape-pass", type=int, default=1, help="TSIM_OPT_TAPE_PASS (1 = G0 / G1 blocks of the tape in a balanced pass of their own)")
    ap.add_argument("--identical", type=int, default=-1, help="timing experiment: every environment gets the inputs of this one (perfect balance)")
    a = ap.parse_args()
    g = np.load(os.path.join(ROOT, "tests", "golden", a.case + ".npz"))
    for lanes in a.lanes:
        ib = g["ibuf"].copy()
        if a.no_gp:
            ib[8] = 0
        if int(ib[3]) > 8 and lanes < 16:
            continue
        sim = BatchedSim((ib, g["dbuf"]), "cuda:0", lanes=lanes)
        dev = sim.device
        if a.max_newton:
            sim.set_option(1, a.max_newton)
        sim.set_option(2, a.vjp_pass)
        sim.set_option(3, a.tac_pass)
        sim.set_option(4, a.tape_pass)
        q0, qd0, u = inputs(g, a.B, a.T, dev)
        if a.zero_u:
            u = torch.zeros_like(u)
        if a.identical >= 0:
            q0 = q0[a.identical:a.identical + 1].repeat(a.B, 1).contiguous()
            u = u[:, a.identical:a.identical + 1].repeat(1, a.B, 1).contiguous()
        for rep in range(a.reps):
            q, qd = q0.clone(), qd0.clone()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            torch.cuda.synchronize()
            e[0].record()
            if not a.grad_only:
                out = sim.forward(q, qd, u, a.T, grad=False, want_status=True)
            e[1].record()
            q, qd = q0.clone(), qd0.clone()
            out = sim.forward(q, qd, u, a.T, grad=True, want_status=True)
            e[2].record()
            dq = torch.ones_like(out["q_traj"])
            dv = torch.ones_like(out["var"]) if out["var"] is not None else None
            dt = torch.full_like(out["tactile"], 1e-3)
            bw = sim.backward(out, u, a.T, dq, dv, dt, want_q0=True)
            e[3].record()
            torch.cuda.synchronize()
            t_ng, t_f, t_b = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])
            st = out["status"]
            steps = a.B * a.T
            print(f"lanes={lanes} B={a.B} T={a.T} rep={rep}: fwd-nograd {t_ng:.1f} ms ({steps / t_ng * 1e3:.3e} steps/s) | "
                  f"fwd+tape {t_f:.1f} ms ({steps / t_f * 1e3:.3e}) | adjoint {t_b:.1f} ms ({steps / t_b * 1e3:.3e}) | "
                  f"fwd+adjoint {steps / (t_f + t_b) * 1e3:.3e} env-steps/s | newton mean {(st & 255).double().mean().item():.2f} "
                  f"max {(st & 255).max().item()} ls mean {((st >> 8) & 255).double().mean().item():.2f} flags {(st >> 16).max().item()} "
                  f"nan {torch.isnan(bw['df_du']).any().item()}", flush=True)
        print("  kernel ms:", {k: (None if v is None else round(v, 2)) for k, v in sim.kernel_times().items()}, flush=True)


if __name__ == "__main__":
    main()
