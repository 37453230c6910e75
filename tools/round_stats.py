"""Development probe: distribution of evaluation rounds per env / warp / block over a trajectory."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from bench import make_inputs
from tactilesimulation_b200.sim import BatchedSim
g = np.load(os.path.join(ROOT, "tests", "golden", "pusher32x13_episodic_s0.npz"))
sim = BatchedSim((g["ibuf"], g["dbuf"]), "cuda:0", lanes=8)
dev = sim.device
B, T = 4096, 200
q0, qd0, u, goal = make_inputs(g["q0"], B, T, 1234)
q, qd, ut = torch.tensor(q0, device=dev), torch.tensor(qd0, device=dev), torch.tensor(u, device=dev)
out = sim.forward(q, qd, ut, T, grad=True, want_status=True, want_contacts=True, want_tactile=False)
st = out["status"].cpu().numpy()
ls = (st >> 8) & 255
rounds = ls + 2                      # initial eval + trials + G0 (rare re-evals ignored)
cm = out["contact_masks"].cpu().numpy().astype(np.uint32)
incontact = (cm[:, :, 1:4] != 0).any(axis=2)
per_env = rounds.sum(axis=0)
print("rounds/step mean", rounds.mean(), "per-env total: mean", per_env.mean(), "max", per_env.max(), "p99", np.percentile(per_env, 99))
w = rounds.reshape(T, B // 4, 4).max(axis=2)          # warp = 4 consecutive envs: max per step
per_warp = w.sum(axis=0)
print("per-warp total: mean", per_warp.mean(), "max", per_warp.max(), "p99", np.percentile(per_warp, 99))
nb = B // 28
blk = per_warp[: nb * 7].reshape(nb, 7).max(axis=1)
print("per-block (7 warps) max-warp total: mean", blk.mean(), "max", blk.max())
print("contact fraction", incontact.mean(), " per-env contact-steps: max", incontact.sum(axis=0).max())
cw = incontact.reshape(T, B // 4, 4).any(axis=2)
print("warp has a contact env: frac of warp-steps", cw.mean())
print("rounds/step when in contact", rounds[incontact].mean(), "when not", rounds[~incontact].mean())
# worst envs
order = np.argsort(-per_env)[:8]
for e in order:
    it = (st[:, e] & 255)
    print("env", int(e), "block", int(e) // 28, "total rounds", int(per_env[e]), "newton iters: max", int(it.max()), "sum", int(it.sum()),
          "ls max", int(ls[:, e].max()), "flags", int((st[:, e] >> 16).max()), "contact steps", int(incontact[:, e].sum()))
print("per-block totals sorted (top 8):", np.sort(blk)[-8:], "median", np.median(blk))
print("histogram of per-env totals:", np.histogram(per_env, bins=[0, 600, 700, 800, 900, 1000, 1200, 1500, 2000, 5000, 100000])[0])
