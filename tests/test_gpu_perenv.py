"""Per-ENVIRONMENT parameters (tsim_scene_set_env_scenes): the reference randomises every
environment object on its own before reset() -- dclaw_rotate_env.py:162-178 (cap damping / radius / end-effector / joint
location), tactile_insertion_env.py:232-275 (contact and tactile coefficients).  Here ONE batched Simulation takes the
same update_* calls with a leading batch dimension; every environment of the batch must reproduce the rollout of the
reference Simulation that was given its parameters (fixture: tests/golden/make_golden.py::perenv_case)."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _dclaw(g, batch):
    from tactilesimulation_b200.layout import scene_from_blob
    from tactilesimulation_b200.redmax import Simulation
    sc = scene_from_blob(g["dclaw_ibuf"], g["dclaw_dbuf"])
    sc.joint_names = [str(x) for x in g["dclaw_joint_names"]]
    sc.body_names = [str(x) for x in g["dclaw_body_names"]]
    for e, name in zip(sc.end_effectors, g["dclaw_ee_names"]):
        e["name"] = str(name)
    return Simulation(sc, batch=batch)


def _dclaw_updates(sim, damping, radius, dxy):
    """the calls of dclaw_rotate_env.py:173-178, values scalar (one env) or [B, ...]"""
    radius, dxy = np.asarray(radius), np.asarray(dxy)
    sim.update_joint_damping("cap", damping)
    sim.update_body_size("cap", np.stack([np.full_like(radius, 0.03), radius], axis=-1))
    sim.update_endeffector_position("cap", np.stack([radius, np.zeros_like(radius), np.zeros_like(radius)], axis=-1))
    sim.update_joint_location("cap", np.concatenate([dxy, np.full(dxy.shape[:-1] + (1,), 0.075)], axis=-1))


def test_dclaw_batch_with_per_environment_parameters_matches_the_reference():
    g = np.load(os.path.join(GOLDEN, "perenv_updates_s0.npz"))
    K, T = g["dclaw_q"].shape[0], g["dclaw_q"].shape[1]
    sim = _dclaw(g, K)
    _dclaw_updates(sim, g["dclaw_damping"], g["dclaw_radius"], g["dclaw_dxy"])
    dev = sim.device
    sim.set_state_init(torch.tensor(g["dclaw_q0"]), torch.zeros((K, 10), dtype=torch.float64))
    sim.reset(False)
    rows = [(t // 5 if t % 5 == 4 else -1) for t in range(T)]
    out = sim.forward_t(T, torch.tensor(g["dclaw_u"], device=dev), tac_rows=rows, want_outputs=True)
    q, var, tac = out["q_traj"].cpu().numpy(), out["var"].cpu().numpy(), out["tactile"].cpu().numpy()
    assert float(np.abs(g["dclaw_tactile"]).max()) > 0
    for e in range(K):
        for t in range(T):
            assert rel_err(q[t, e], g["dclaw_q"][e, t]) <= 1e-9, (e, t)
            assert rel_err(var[t, e], g["dclaw_var"][e, t]) <= 1e-9, (e, t)
        assert rel_err(tac[:, e], g["dclaw_tactile"][e]) <= 1e-8, e
    # the parameters matter: environment 0 run with environment 1's parameters leaves its reference rollout
    sim1 = _dclaw(g, 1)
    _dclaw_updates(sim1, float(g["dclaw_damping"][1]), g["dclaw_radius"][1], g["dclaw_dxy"][1])
    sim1.set_state_init(g["dclaw_q0"][0], np.zeros(10))
    sim1.reset(False)
    o1 = sim1.forward_t(T, torch.tensor(g["dclaw_u"][:, 0:1], device=dev).contiguous(), want_outputs=True)
    assert rel_err(o1["q_traj"][-1, 0].cpu().numpy(), g["dclaw_q"][0, -1]) > 1e-6


def test_per_environment_batch_equals_one_simulation_per_environment_including_the_adjoint():
    """forward + backward() of the batch with per-environment parameters = each environment alone in a batch-1 Simulation
    that got the same values as scalars (the reference's way), outputs and gradients."""
    g = np.load(os.path.join(GOLDEN, "perenv_updates_s0.npz"))
    K, T = g["dclaw_q"].shape[0], 20
    rng = np.random.default_rng(5)
    dq, dv = rng.normal(size=(T, K, 10)), rng.normal(size=(T, K, 12))
    dt = 1e-3 * rng.normal(size=(T, K, 2718))
    sim = _dclaw(g, K)
    _dclaw_updates(sim, g["dclaw_damping"], g["dclaw_radius"], g["dclaw_dxy"])
    dev = sim.device
    u = torch.tensor(g["dclaw_u"][:T], device=dev).contiguous()
    sim.set_state_init(torch.tensor(g["dclaw_q0"]), torch.zeros((K, 10), dtype=torch.float64))
    sim.reset(True)
    out = sim.forward_t(T, u, want_outputs=True)
    df_du, dq0, dqd0 = sim.backward_t(torch.tensor(dq, device=dev), torch.tensor(dv, device=dev), torch.tensor(dt, device=dev))
    for e in range(K):
        s1 = _dclaw(g, 1)
        _dclaw_updates(s1, float(g["dclaw_damping"][e]), g["dclaw_radius"][e], g["dclaw_dxy"][e])
        s1.set_state_init(g["dclaw_q0"][e], np.zeros(10))
        s1.reset(True)
        o1 = s1.forward_t(T, u[:, e:e + 1].contiguous(), want_outputs=True)
        g1 = s1.backward_t(torch.tensor(dq[:, e:e + 1], device=dev), torch.tensor(dv[:, e:e + 1], device=dev),
                           torch.tensor(dt[:, e:e + 1], device=dev))
        assert torch.equal(o1["q_traj"][:, 0], out["q_traj"][:, e]), e
        assert torch.equal(o1["tactile"][:, 0], out["tactile"][:, e]), e
        assert rel_err(g1[0][:, 0].cpu().numpy(), df_du[:, e].cpu().numpy()) <= 1e-12, e
        assert rel_err(g1[1][0].cpu().numpy(), dq0[e].cpu().numpy()) <= 1e-12, e
    # dropping the per-environment updates returns to the scene's own parameters for every environment
    sim.clear_env_parameters()
    sim.reset(False)
    o2 = sim.forward_t(5, u[:5].contiguous(), want_outputs=True)
    base = _dclaw(g, K)
    base.set_state_init(torch.tensor(g["dclaw_q0"]), torch.zeros((K, 10), dtype=torch.float64))
    base.reset(False)
    o3 = base.forward_t(5, u[:5].contiguous(), want_outputs=True)
    assert torch.equal(o2["q_traj"], o3["q_traj"])


def test_insertion_batch_with_per_environment_contact_and_tactile_coefficients():
    from tactilesimulation_b200.layout import scene_from_blob
    from tactilesimulation_b200.redmax import Simulation
    g = np.load(os.path.join(GOLDEN, "perenv_updates_s0.npz"))
    K, T = g["ins_q"].shape[0], g["ins_q"].shape[1]
    sc = scene_from_blob(g["ins_ibuf"], g["ins_dbuf"])
    sc.body_names = [str(x) for x in g["ins_body_names"]]
    for s_, name in zip(sc.sensors, g["ins_sensor_names"]):
        s_.name = str(name)
    sim = Simulation(sc, batch=K)
    c, tp = g["ins_cpar"], g["ins_tpar"]
    for pad in ("tactile_pad_left", "tactile_pad_right"):         # tactile_insertion_env.py:254-275
        sim.update_contact_parameters(pad, "box", c[:, 0], c[:, 1], c[:, 2], c[:, 3])
        sim.update_tactile_parameters(pad, tp[:, 0], tp[:, 1], tp[:, 2], tp[:, 3])
    dev = sim.device
    sim.set_state_init(torch.tensor(np.tile(g["ins_q0"], (K, 1))), torch.zeros((K, 12), dtype=torch.float64))
    sim.reset(False)
    rows = [(t // 5 if t % 5 == 4 else -1) for t in range(T)]
    u = torch.tensor(np.tile(g["ins_u"][:, None, :], (1, K, 1)), device=dev).contiguous()
    out = sim.forward_t(T, u, tac_rows=rows, want_outputs=True)
    q, tac = out["q_traj"].cpu().numpy(), out["tactile"].cpu().numpy()
    assert float(np.abs(g["ins_tactile"]).max()) > 0
    for e in range(K):
        for t in range(T):
            assert rel_err(q[t, e], g["ins_q"][e, t]) <= 1e-9, (e, t)
        assert rel_err(tac[:, e], g["ins_tactile"][e]) <= 1e-8, e
    assert rel_err(g["ins_q"][0, -1], g["ins_q"][1, -1]) > 1e-7           # the coefficients do change the rollouts
