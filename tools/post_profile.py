"""Development probe: clock64 split of the Newton bookkeeping of the forward kernel (library built with
tools/experiments/post_profile.patch applied and -DTS_PROFILE -DTS_PROFILE_POST, pointed to by TSIM_B200_LIB).
PCASE / PLANES / PB / PT as in tools/cycle_profile.py."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench import make_inputs
from tactilesimulation_b200.sim import BatchedSim
CASE = os.environ.get("PCASE", "pusher32x13_episodic_s0")
LANES = int(os.environ.get("PLANES", 8))
g = np.load(os.path.join(ROOT, "tests", "golden", CASE + ".npz"))
sim = BatchedSim((g["ibuf"], g["dbuf"]), "cuda:0", lanes=LANES)
dev = sim.device
B, T = int(os.environ.get("PB", 4096)), int(os.environ.get("PT", 100))
NT = 224 // LANES
nthreads = ((B * LANES + 223) // 224) * 224
prof = torch.zeros((nthreads, 16), dtype=torch.int64, device=dev)
getattr(sim.lib, "tsim_debug_set_prof_v%d" % (8 if LANES == 8 else 16))(ctypes.c_void_p(prof.data_ptr()))
if os.environ.get("PWORK"):          # the inputs of a bench workload (bench.workload_inputs)
    import bench
    q0, qd0, u, goal = bench.workload_inputs(os.environ["PWORK"], g, B, T, 1234)
    ut = torch.tensor(u, device=dev)
elif LANES == 8:
    q0, qd0, u, goal = make_inputs(g["q0"], B, T, 1234)
    ut = torch.tensor(u, device=dev)
else:
    from perf_probe import inputs
    q0, qd0, ut = inputs(g, B, T, dev)
    q0, qd0 = q0.cpu().numpy(), qd0.cpu().numpy()
for rep in range(2):
    q, qd = torch.tensor(q0, device=dev), torch.tensor(qd0, device=dev)
    out = sim.forward(q, qd, ut, T, grad=os.environ.get("PGRAD", "1") == "1")
    torch.cuda.synchronize()
p = prof.cpu().numpy().reshape(-1, NT, LANES, 16)[:, :, 0, :].astype(float)
tot = p[:, :, 7]
names = {0: "kinematics", 1: "ground", 2: "gp", 3: "inward", 4: "vote wait", 5: "round total", 6: "epilogue", 7: "kernel",
         8: "eval_columns: sync + stage inputs", 15: "eval_g total", 9: "column extraction", 10: "post: line-search logic",
         14: "post: batched line search", 11: "post: tape H store", 12: "post: norm + transpose + LU", 13: "post: dx store"}
b = int(tot.max(axis=1).argmax())
tb = int(p[b, :, 5].argmax())
print("slowest block", b, "busiest tile", tb, "kernel %.4g" % tot[b].max())
for i in [7, 4, 5, 15, 0, 1, 2, 3, 8, 9, 10, 14, 11, 12, 13, 6]:
    v = p[:, :, i]
    print(f"{names[i]:36s} mean {v.mean():.4g} ({100 * v.mean() / tot.mean():5.1f}% of kernel) | busiest tile of the slowest block {p[b, tb, i]:.4g} ({100 * p[b, tb, i] / tot[b].max():5.1f}%)")
