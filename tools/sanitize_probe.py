"""Small forward + backward on the pusher scene for compute-sanitizer runs (racecheck / memcheck / synccheck)."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from bench import make_inputs
from tactilesimulation_b200.sim import BatchedSim
CASE = os.environ.get("PCASE", "pusher32x13_episodic_s0")       # PCASE: another scene (16-lane variants: DClaw ...)
g = np.load(os.path.join(ROOT, "tests", "golden", CASE + ".npz"))
B, T = int(os.environ.get("PB", 56)), int(os.environ.get("PT", 12))
sim = BatchedSim((g["ibuf"], g["dbuf"]), "cuda:0")
dev = sim.device
if CASE.startswith("pusher"):
    q0, qd0, u, goal = make_inputs(g["q0"], B, T, 7)
    q0[:, 1] = 0.0005            # start in contact: the cooperative point phase runs from the first round
    u[:, :, 0] = 0.9
    tq, tqd, tu = torch.tensor(q0, device=dev), torch.tensor(qd0, device=dev), torch.tensor(u, device=dev)
else:
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from perf_probe import inputs
    tq, tqd, tu = inputs(g, B, T, dev)
out = sim.forward(tq, tqd, tu, T, grad=True, want_status=True, want_contacts=True)
bw = sim.backward(out, tu, T, torch.ones_like(out["q_traj"]), torch.ones_like(out["var"]), torch.full_like(out["tactile"], 1e-3), want_q0=True)
torch.cuda.synchronize()
print("ok", float(out["tactile"].abs().max()), int((out["contact_masks"][:, :, 1:] != 0).any(dim=2).sum()), float(bw["df_du"].abs().max()))
