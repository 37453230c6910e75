"""Development probe: evaluation rounds per environment / block of the bench workload (from the status words)."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from bench import make_inputs
from tactilesimulation_b200.sim import BatchedSim
g = np.load(os.path.join(ROOT, "tests", "golden", "pusher32x13_episodic_s0.npz"))
B, T = 4096, 200
sim = BatchedSim((g["ibuf"], g["dbuf"]), "cuda:0")
q0, qd0, u, goal = make_inputs(g["q0"], B, T, 1234)
dev = sim.device
out = sim.forward(torch.tensor(q0, device=dev), torch.tensor(qd0, device=dev), torch.tensor(u, device=dev), T, grad=True,
                  want_status=True, want_contacts=True, want_tactile=False)
st = out["status"].cpu().numpy()
it, ls = st & 255, (st >> 8) & 255
rounds = (ls + 1).sum(axis=0)                    # per env: one evaluation at x0 + one per line-search evaluation (lower bound)
cm = out["contact_masks"].cpu().numpy()
touch = (cm[:, :, 1:4] != 0).any(axis=2)
print("rounds per env: mean %.0f  p50 %.0f  p90 %.0f  p99 %.0f  max %d" % (rounds.mean(), *np.percentile(rounds, [50, 90, 99]), rounds.max()))
blk = rounds[: (B // 28) * 28].reshape(-1, 28)
bm = blk.max(axis=1)
print("per-block max: mean %.0f  p50 %.0f  p90 %.0f  max %d ; per-block mean of envs: %.0f" % (bm.mean(), *np.percentile(bm, [50, 90]), bm.max(), blk.mean()))
top = np.argsort(-rounds)[:12]
print("top envs:", [(int(e), int(rounds[e]), int(it[:, e].max()), int(touch[:, e].sum())) for e in top])
print("contact steps per env: mean %.1f ; corr(rounds, contact steps) %.2f" % (touch.sum(axis=0).mean(), np.corrcoef(rounds, touch.sum(axis=0))[0, 1]))
# rounds of contact steps vs free steps
print("ls+1 per step: in contact %.2f, free %.2f ; fraction of steps in contact %.3f" % ((ls + 1)[touch].mean(), (ls + 1)[~touch].mean(), touch.mean()))
