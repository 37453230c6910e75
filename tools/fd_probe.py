import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from bench import make_inputs
from tactilesimulation_b200.sim import BatchedSim
g = np.load("tests/golden/pusher32x13_episodic_s0.npz")
sim = BatchedSim((g["ibuf"], g["dbuf"]), device="cuda:0")
dev = sim.device
def run(contact, Ts, eps, Bs=256):
    q0, qd0, u, _ = make_inputs(g["q0"], Bs, Ts, 7)
    if contact:
        q0[:, 1] = 0.0005
        u[:, :, 0] = 0.5 + 0.4 * u[:, :, 0]
    rng = np.random.default_rng(11)
    du, dq0 = rng.normal(size=u.shape), rng.normal(size=q0.shape) * 1e-2
    du[:, :, 5] = 0.0
    tq0, tqd0, tu = torch.tensor(q0, device=dev), torch.tensor(qd0, device=dev), torch.tensor(u, device=dev)
    out = sim.forward(tq0.clone(), tqd0.clone(), tu, Ts, grad=True)
    gen = torch.Generator(device=dev).manual_seed(5)
    mk = lambda t, s: (s * torch.randn(t.shape, generator=gen, device=dev, dtype=torch.float64)).contiguous()
    wq, wv, wt = mk(out["q_traj"], 1.0), mk(out["var"], 1.0), mk(out["tactile"], 1e-3)
    loss = lambda o: (o["q_traj"] * wq).sum(dim=(0, 2)) + (o["var"] * wv).sum(dim=(0, 2)) + (o["tactile"] * wt).sum(dim=(0, 2))
    bw = sim.backward(out, tu, Ts, wq, wv, wt, want_q0=True)
    lin = (bw["df_du"] * torch.tensor(du, device=dev)).sum(dim=(0, 2)) + (bw["df_dq0"] * torch.tensor(dq0, device=dev)).sum(dim=1)
    ls = []
    for s in (+1.0, -1.0):
        o = sim.forward(torch.tensor(q0 + s * eps * dq0, device=dev), tqd0.clone(), torch.tensor(u + s * eps * du, device=dev), Ts)
        ls.append(loss(o))
    fd = (ls[0] - ls[1]) / (2 * eps)
    rel = ((fd - lin).abs() / (fd.abs() + 1e-9)).cpu().numpy()
    touching = float((out["tactile"].abs().amax(dim=(0, 2)) > 0).double().mean())
    print(f"contact={contact} T={Ts} eps={eps:g}: median {np.median(rel):.2e} p90 {np.percentile(rel,90):.2e} frac<1e-3 {(rel<1e-3).mean():.2f} touching {touching:.2f} |fd| med {fd.abs().median().item():.2e}")
for contact in (0, 1):
    for Ts in (5, 30):
        for eps in (1e-3, 1e-4, 1e-5, 1e-6):
            run(contact, Ts, eps)
