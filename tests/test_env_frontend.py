"""Batched TactilePush front-end: its observation / reward are checked against a loop
restatement of the reference's per-environment formulas (R/envs/tactile_push_env.py:83-107, :203-211) on CPU
tensors; the GPU test runs a gd.py-style analytic policy-gradient epoch through it."""
import math
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN


def _ref_obs(q, goal, privilege):
    """tactile_push_env.py:83-107, one environment."""
    rot = q[0:1]
    c, s = math.cos(-rot[0]), math.sin(-rot[0])
    R = np.array([[c, -s], [s, c]])
    obj = R @ q[3:5] - q[1:3]
    gl = R @ goal[0:2] - q[1:3]
    state = np.concatenate([gl, goal[2:3] - rot])
    return np.concatenate([obj, q[6:7] - rot, state]) if privilege else state


def _ref_reward(q, var, u, goal):
    """tactile_push_env.py:203-211, one environment."""
    r_pos = -np.sum(((q[3:5] - goal[0:2]) / 0.01) ** 2) * 0.01
    r_rot = -(((q[6] - goal[2]) / (np.pi / 36.0)) ** 2) * 0.1
    r_touch = -np.sum((var[0:3] - var[3:6]) ** 2) / (0.02 ** 2)
    return r_pos + r_rot + r_touch - np.sum(u ** 2) * 0.1


def test_observation_and_reward_match_the_reference_formulas():
    from tactilesimulation_b200.envs.tactile_push import push_observation, push_reward
    rng = np.random.default_rng(0)
    B = 17
    q, var, u, goal = rng.normal(size=(B, 7)), rng.normal(size=(B, 6)), rng.normal(size=(B, 3)), rng.normal(size=(B, 3))
    tq, tv, tu, tg = (torch.tensor(x) for x in (q, var, u, goal))
    for priv in (False, True):
        o = push_observation(tq, tg, privilege=priv).numpy()
        for e in range(B):
            assert np.allclose(o[e], _ref_obs(q[e], goal[e], priv), rtol=1e-13, atol=1e-13)
    r, info = push_reward(tq, tv, tu, tg)
    for e in range(B):
        assert np.isclose(r[e].item(), _ref_reward(q[e], var[e], u[e], goal[e]), rtol=1e-13)
    assert set(info) == {"reward_pos", "reward_rot", "reward_touch", "reward_action"}


@pytest.mark.gpu
def test_gd_style_epoch_through_the_batched_env():
    """obs -> MLP actor -> env.step chained over a short horizon; loss = -mean episode reward (gd.py:258 with
    num_episodes = B); backward through StepSimFunction; the policy gradient is finite, non-zero and matches a
    central finite difference along a random parameter direction."""
    from tactilesimulation_b200.envs import BatchedTactilePushEnv
    from tactilesimulation_b200.layout import scene_from_blob
    from tactilesimulation_b200.redmax import Simulation
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz"))
    B, steps = 64, 6
    sim = Simulation(scene_from_blob(g["ibuf"], g["dbuf"]), batch=B)
    sim.set_q_init(np.tile(g["q0"], (B, 1)))
    env = BatchedTactilePushEnv(sim, observation_type="tactile_flatten", gradient=True, tactile_rows=13, tactile_cols=10, seed=3)
    torch.manual_seed(0)
    actor = torch.nn.Sequential(torch.nn.Linear(3 + 390, 32), torch.nn.Tanh(), torch.nn.Linear(32, 3)).double().to(sim.device)

    def epoch():
        env.gen.manual_seed(3)
        obs = env.reset()
        assert obs.shape == (B, 393)
        total = 0.0
        for _ in range(steps):
            obs, r, done, info = env.step(actor(obs))
            total = total + r
        return -(total.mean())

    loss = epoch()
    loss.backward()
    grads = torch.cat([p.grad.reshape(-1) for p in actor.parameters()])
    assert torch.isfinite(grads).all() and float(grads.abs().max()) > 0
    # directional finite difference in parameter space (free motion over 6 gym steps: smooth)
    direction = [torch.randn_like(p) for p in actor.parameters()]
    lin = sum((p.grad * d).sum() for p, d in zip(actor.parameters(), direction)).item()
    eps, vals = 1e-5, []
    for sgn in (+1.0, -1.0):
        with torch.no_grad():
            for p, d in zip(actor.parameters(), direction):
                p.add_(sgn * eps * d)
            vals.append(epoch().item())
            for p, d in zip(actor.parameters(), direction):
                p.sub_(sgn * eps * d)
    fd = (vals[0] - vals[1]) / (2 * eps)
    assert abs(fd - lin) <= 1e-3 * max(abs(fd), 1e-9), (fd, lin)
