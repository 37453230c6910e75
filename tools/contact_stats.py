"""Development probe: contact statistics of the bench workload + throughput vs batch size."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from bench import make_inputs
from tactilesimulation_b200.sim import BatchedSim
g = np.load(os.path.join(ROOT, "tests", "golden", "pusher32x13_episodic_s0.npz"))
sim = BatchedSim((g["ibuf"], g["dbuf"]), "cuda:0", lanes=8)
dev = sim.device
B, T = 1024, 200
q0, qd0, u, goal = make_inputs(g["q0"], B, T, 1234)
q, qd, ut = torch.tensor(q0, device=dev), torch.tensor(qd0, device=dev), torch.tensor(u, device=dev)
out = sim.forward(q, qd, ut, T, grad=False, want_status=True, want_contacts=True)
cm = out["contact_masks"].cpu().numpy().astype(np.uint32)
pop = np.vectorize(lambda x: bin(int(x)).count("1"))
gp = pop(cm[:, :, 1]) + pop(cm[:, :, 2]) + pop(cm[:, :, 3])
gr = pop(cm[:, :, 0])
mb = out["marker_body"].cpu().numpy()
st = out["status"].cpu().numpy()
print("gp contact: frac env-steps with any", (gp > 0).mean(), "mean active pts when any", gp[gp > 0].mean(), "overall mean", gp.mean())
print("by time quartile:", [(gp[i * 50:(i + 1) * 50] > 0).mean() for i in range(4)])
print("ground active pts mean", gr.mean())
print("markers in contact: mean", (mb >= 0).sum(axis=2).mean(), "frac steps any", ((mb >= 0).sum(axis=2) > 0).mean())
print("newton iters mean", (st & 255).mean(), "hist", np.bincount((st & 255).ravel())[:12], "ls mean", ((st >> 8) & 255).mean())
for Bb in (4096, 16384):
    q0, qd0, u, goal = make_inputs(g["q0"], Bb, 20, 1)
    q, qd, ut = torch.tensor(q0, device=dev), torch.tensor(qd0, device=dev), torch.tensor(u, device=dev)
    for rep in range(2):
        qq, qqd = q.clone(), qd.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = sim.forward(qq, qqd, ut, 20, grad=True); e1.record(); torch.cuda.synchronize()
        print("B", Bb, "fwd+tape T=20 ms", e0.elapsed_time(e1), "env-steps/s %.3e" % (Bb * 20 / e0.elapsed_time(e1) * 1e3))
    del o
