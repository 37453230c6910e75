"""Parity of the CUDA path with the UNMODIFIED reference on the MEASURED workload itself: bench.make_inputs(seed=1234)
-- TactilePush 32x13, B = 4096, T = 200, tanh(N(0,1)) actions, random pushes -- which is what bench.py times.

One B = 4096 run of the CUDA path (forward with tape + adjoint with the bench's cotangents: gradient of the TactilePush
reward on q / var, 1e-3 on the tactile field) is compared row by row with the reference C++ (oracle/_ref/redmax_py,
which travels to the GPU box as a prebuilt .so; oracle/_ref/redmax_probe for contact index sets and Newton counters) run
LIVE on eight environments of the batch: five fixed ones, env 2826 (runs into the reference's cap of 140 Newton
iterations, DH/Simulation.cpp:1155), and the two environments with the most pad-box contact steps in the CUDA run.

Gates: q, qdot, var <= 1e-9, tactile <= 1e-8 (print_error metric of DH/Utils.h:315-319), contact
index sets and marker->body ids identical, Newton iterations and line-search evaluations per step identical (status
bits 0-15), df_du / df_dq0 / df_dqdot0 <= 1e-6.
"""
import os
import sys

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN, REF_DIR, rel_err

pytestmark = pytest.mark.gpu

B, T = 4096, 200
XML = os.path.join(REF_DIR, "assets", "pusher", "pusher_32x13.xml")
FIXED = (0, 1, 17, 1950, 2826, 4095)


def _have_ref():
    return os.path.exists(XML) and os.path.isdir(REF_DIR) and \
        any(f.startswith("redmax_probe") and f.endswith(".so") for f in os.listdir(REF_DIR))


def _ids(words):
    out = []
    for w, x in enumerate(words):
        x = int(x) & 0xffffffff
        for b in range(32):
            if (x >> b) & 1:
                out.append(32 * w + b)
    return out


@pytest.fixture(scope="module")
def run():
    if not _have_ref():
        pytest.fail("oracle/_ref (reference build + assets) is missing on this box: run oracle/build_ref.sh before gpurun")
    from bench import make_inputs, reward_cotangents
    from tactilesimulation_b200.sim import BatchedSim
    g = np.load(os.path.join(GOLDEN, "pusher32x13_episodic_s0.npz"))
    sim = BatchedSim((g["ibuf"], g["dbuf"]), device="cuda:0")
    q0, qd0, u, goal = make_inputs(g["q0"], B, T, 1234)
    dev = sim.device
    tq, tqd, tu = torch.tensor(q0, device=dev), torch.tensor(qd0, device=dev), torch.tensor(u, device=dev)
    out = sim.forward(tq, tqd, tu, T, grad=True, want_status=True, want_contacts=True)
    scale = 1.0 / B
    dq, dv = reward_cotangents(out["q_traj"], out["var"], torch.tensor(goal, device=dev), scale)
    dtac = torch.full((T, B, sim.ndof_tactile), 1e-3, dtype=torch.float64, device=dev)
    bw = sim.backward(out, tu, T, dq, dv, dtac, want_q0=True)
    torch.cuda.synchronize()
    del dtac
    cm = out["contact_masks"].cpu().numpy()                    # [T,B,4]: word 0 ground, words 1..3 pad-box
    touch_steps = (cm[:, :, 1:4] != 0).any(axis=2).sum(axis=0)
    busiest = [int(e) for e in np.argsort(-touch_steps) if int(e) not in FIXED][:2]
    envs = list(FIXED) + busiest
    sel = torch.tensor(envs, device=dev)
    res = dict(envs=envs, touch_steps=touch_steps, q0=q0, u=u, goal=goal, scale=scale, cm=cm[:, envs],
               q=out["q_traj"][:, sel].cpu().numpy(), qd=out["qd_traj"][:, sel].cpu().numpy(),
               var=out["var"][:, sel].cpu().numpy(), tac=out["tactile"][:, sel].cpu().numpy(),
               mb=out["marker_body"][:, sel].cpu().numpy(), status=out["status"][:, sel].cpu().numpy(),
               dq=dq[:, sel].cpu().numpy(), dv=dv[:, sel].cpu().numpy(),
               df_du=bw["df_du"][:, sel].cpu().numpy(), df_dq0=bw["df_dq0"][sel].cpu().numpy(),
               df_dqdot0=bw["df_dqdot0"][sel].cpu().numpy(), ntac=sim.ndof_tactile)
    del out, bw
    torch.cuda.empty_cache()
    return res


def test_workload_has_the_hard_cases(run):
    assert int(run["touch_steps"][run["envs"][-2]]) >= 150, "no environment with >= 150 pad-box contact steps"
    i = run["envs"].index(2826)
    assert int((run["status"][:, i] & 0xff).max()) == 140          # the environment that hits the iteration cap


@pytest.mark.parametrize("slot", range(8))
def test_bench_workload_rows_match_the_live_reference(run, slot):
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import redmax_probe
    import redmax_py
    e = run["envs"][slot]
    sim = redmax_py.Simulation(XML)
    probe = redmax_probe.ProbeSimulation(XML)
    n, nu, nt = sim.ndof_r, sim.ndof_u, sim.ndof_tactile
    assert nt == run["ntac"]
    for s in (sim, probe):
        s.set_state_init(run["q0"][e], np.zeros(n))
        s.reset(True)
    probe.newton_counts()
    worst = dict(q=0.0, qd=0.0, var=0.0, tac=0.0)
    for t in range(T):
        for s in (sim, probe):
            s.set_u(run["u"][t, e])
            s.forward(1)
        it, ls = probe.newton_counts()
        st = int(run["status"][t, slot])
        assert (st & 0xff, (st >> 8) & 0xff) == (it & 0xff, ls & 0xff), (e, t, st & 0xff, (st >> 8) & 0xff, it, ls)
        worst["q"] = max(worst["q"], rel_err(run["q"][t, slot], sim.get_q()))
        worst["qd"] = max(worst["qd"], rel_err(run["qd"][t, slot], sim.get_qdot()))
        worst["var"] = max(worst["var"], rel_err(run["var"][t, slot], sim.get_variables()))
        worst["tac"] = max(worst["tac"], rel_err(run["tac"][t, slot], sim.get_tactile_force_vector()))
        cs = probe.contact_sets()
        assert _ids(run["cm"][t, slot, 0:1]) == [int(x) for x in cs["ground"][0]], (e, t)
        assert _ids(run["cm"][t, slot, 1:4]) == [int(x) for x in cs["gp"][0]], (e, t)
        assert np.array_equal(run["mb"][t, slot], np.asarray(cs["marker_body"][0], dtype=np.int32)), (e, t)
    assert worst["q"] <= 1e-9 and worst["qd"] <= 1e-9 and worst["var"] <= 1e-9, (e, worst)
    assert worst["tac"] <= 1e-8, (e, worst)
    bi = sim.backward_info
    bi.set_flags(True, True, False, True)
    bi.df_dq = run["dq"][:, slot].reshape(-1)
    bi.df_dvar = run["dv"][:, slot].reshape(-1)
    bi.df_dtactile = np.full(nt * T, 1e-3)
    bi.df_dq0, bi.df_dqdot0, bi.df_du = np.zeros(n), np.zeros(n), np.zeros(nu * T)
    sim.backward()
    br = sim.backward_results
    assert rel_err(run["df_du"][:, slot], np.array(br.df_du).reshape(T, nu)) <= 1e-6, e
    assert rel_err(run["df_dq0"][slot], np.array(br.df_dq0)) <= 1e-6, e
    assert rel_err(run["df_dqdot0"][slot], np.array(br.df_dqdot0)) <= 1e-6, e
    print(f"env {e}: touch steps {int(run['touch_steps'][e])}, worst rel err {worst}")
