"""Development probe: clock64 breakdown of the forward kernel (needs a -DTS_PROFILE build of the library,
pointed to by TSIM_B200_LIB)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from bench import make_inputs
from tactilesimulation_b200.sim import BatchedSim
CASE = os.environ.get("PCASE", "pusher32x13_episodic_s0")       # PCASE / PLANES: another scene (16-lane variants)
LANES = int(os.environ.get("PLANES", 8))
g = np.load(os.path.join(ROOT, "tests", "golden", CASE + ".npz"))
sim = BatchedSim((g["ibuf"], g["dbuf"]), "cuda:0", lanes=LANES)
dev = sim.device
B, T = int(os.environ.get('PB', 4096)), int(os.environ.get('PT', 50))
NT = 224 // LANES
nthreads = ((B * LANES + 223) // 224) * 224
prof = torch.zeros((nthreads, 16), dtype=torch.int64, device=dev)
getattr(sim.lib, "tsim_debug_set_prof_v%d" % (8 if LANES == 8 else 16))(ctypes.c_void_p(prof.data_ptr()))
if LANES == 8:
    q0, qd0, u, goal = make_inputs(g["q0"], B, T, 1234)
    ut = torch.tensor(u, device=dev)
else:
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from perf_probe import inputs
    q0, qd0, ut = inputs(g, B, T, dev)
    q0, qd0 = q0.cpu().numpy(), qd0.cpu().numpy()
for rep in range(3):
    q, qd = torch.tensor(q0, device=dev), torch.tensor(qd0, device=dev)
    out = sim.forward(q, qd, ut, T, grad=True)
    torch.cuda.synchronize()
p = prof.cpu().numpy().reshape(-1, NT, LANES, 16)[:, :, 0, :]      # [block, tile, slot] lane 0 of each tile
names = ["kinematics+dyn", "ground", "gp", "inward", "vote wait", "step_round total", "epilogue", "kernel total"]
tot = p[:, :, 7].astype(float)
bt = tot.max(axis=1)
print("per-block kernel cycles percentiles (min/25/50/75/90/99/max):", np.percentile(bt, [0, 25, 50, 75, 90, 99, 100]).round(-3))
print("blocks", p.shape[0], "kernel cycles: mean %.3g max %.3g (%.2f ms at 1.965 GHz)" % (tot.mean(), tot.max(), tot.max() / 1.965e6))
for i, nm in enumerate(names):
    v = p[:, :, i].astype(float)
    print(f"{nm:18s} mean {v.mean():.4g} ({100 * v.mean() / tot.mean():5.1f}% of kernel)  max-tile {v.max():.4g}")
# slowest block
b = int(tot.mean(axis=1).argmax())
print("slowest block", b, {nm: float(p[b, :, i].mean()) for i, nm in enumerate(names)})

v = p[:, :, 15].astype(float)
print("batched line search mean %.4g (%5.1f%% of kernel) max-tile %.4g; in the slowest block: mean %.4g max-tile %.4g of %.4g" % (
    v.mean(), 100 * v.mean() / tot.mean(), v.max(), v[b].mean(), v[b].max(), tot[b].max()))
# cooperative contact-point phase (slots 8..14)
if p[:, :, 9].sum() > 0:
    ph = p[:, 0, 9].astype(float)          # phases with items, per block (tile 0)
    print("coop: phases with items per block mean %.0f, items per phase %.1f, batches per phase %.2f" % (
        ph.mean(), p[:, 0, 8].sum() / ph.sum(), p[:, 0, 10].sum() / ph.sum()))
    for i, nm in ((11, "coop publish"), (12, "coop compute"), (13, "coop barrier wait"), (14, "coop accumulate"), (15, "batched line search")):
        v = p[:, :, i].astype(float)
        print(f"{nm:18s} mean {v.mean():.4g} ({100 * v.mean() / tot.mean():5.1f}% of kernel)  per phase {v.mean() / ph.mean():.0f} cycles")
