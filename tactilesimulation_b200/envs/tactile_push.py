"""Batched TactilePush front-end: the task logic of ``R/envs/tactile_push_env.py`` (observation, reward, reset
randomisation, random external pushes) as batched torch ops on the device, over the batched ``StepSimFunction``.

The reference runs one environment per python object and crosses tensor -> numpy -> C++ per sub-step
(``tactile_push_env.py:174-232``, ``gd.py:224-259``); after the simulation kernel is fast that loop is the
bottleneck by orders of magnitude.  Here B environments advance in one kernel launch per gym
step and observation / reward are elementwise torch ops on [B, ...] tensors, so a ``gd.py``-style analytic
policy-gradient epoch is: ``obs = env.reset(); loop: u = actor(obs); obs, r, done, info = env.step(u)``, then
``(-r_sum.mean()).backward()`` -- the same loss as ``gd.py:258`` with ``num_episodes = B``.

Semantics kept from the reference (line numbers of ``R/envs/tactile_push_env.py``):
  * reset (:133-172): ``q[1] = -0.001``, ``q[4] ~ U(-0.02, 0.02)``; goal xy ~ U([0.15,-0.2],[0.25,0.2]), goal rot ~
    U(y*pi - pi/16, y*pi + pi/16); initial tactile read-out;
  * step (:174-232): ``action = tanh(u)`` on the three gripper controls, external force on the box resampled every 10
    steps (p = 0.5 zero, else U(-1,1)^2), ``frame_skip = 5`` sub-steps with tactile / variables on the last one,
    reward = pos + rot + touch + action terms with the reference's scales;
  * observation (:72-131): goal pose in the gripper's local frame (+ object pose for "privilege"), concatenated with the
    flattened tactile field ("tactile_flatten") or returned next to the [3, rows, cols] map ("tactile_map").
The per-environment random streams are torch generators on the device (the reference uses one numpy RandomState
per env; draws are i.i.d. with the same distributions, not the same numbers).
"""
import math
from typing import Optional, Tuple

import torch

from ..redmax import Simulation
from ..torch_functions import StepSimFunction


def push_observation(q: torch.Tensor, goal: torch.Tensor, privilege: bool = False) -> torch.Tensor:
    """State observation of ``tactile_push_env.py:83-107`` for q [B,7], goal [B,3] -> [B,3] (or [B,6])."""
    rot = q[:, 0:1]
    c, s = torch.cos(-rot), torch.sin(-rot)
    gripper = q[:, 1:3]

    def to_local(p):
        return torch.cat([c * p[:, 0:1] - s * p[:, 1:2], s * p[:, 0:1] + c * p[:, 1:2]], dim=1) - gripper
    goal_local = torch.cat([to_local(goal[:, 0:2]), goal[:, 2:3] - rot], dim=1)
    if not privilege:
        return goal_local
    return torch.cat([to_local(q[:, 3:5]), q[:, 6:7] - rot, goal_local], dim=1)


def push_reward(q: torch.Tensor, var: torch.Tensor, u: torch.Tensor, goal: torch.Tensor):
    """Reward terms of ``tactile_push_env.py:203-211`` for q [B,7], var [B,6], raw action u [B,3], goal [B,3]."""
    r_pos = -(((q[:, 3:5] - goal[:, 0:2]) / 0.01) ** 2).sum(dim=1) * 0.01
    r_rot = -(((q[:, 6] - goal[:, 2]) / (math.pi / 36.0)) ** 2) * 0.1
    r_touch = -((var[:, 0:3] - var[:, 3:6]) ** 2).sum(dim=1) / (0.02 ** 2)
    r_action = -(u ** 2).sum(dim=1) * 0.1
    return r_pos + r_rot + r_touch + r_action, dict(reward_pos=r_pos, reward_rot=r_rot, reward_touch=r_touch,
                                                    reward_action=r_action)


class BatchedTactilePushEnv:
    """B TactilePush environments on one GPU.  ``sim`` is a ``Simulation`` of the pusher scene with ``batch=B``."""
    frame_skip = 5
    max_episode_steps = 100       # R/envs/__init__.py:9-13

    def __init__(self, sim: Simulation, observation_type: str = "tactile_flatten", gradient: bool = True,
                 tactile_rows: Optional[int] = None, tactile_cols: Optional[int] = None, seed: int = 0):
        if observation_type not in ("tactile_flatten", "tactile_map", "privilege", "no_tactile"):
            raise NotImplementedError(observation_type)
        if sim.ndof_r != 7 or sim.ndof_u != 6 or sim.ndof_var != 6:
            raise ValueError("BatchedTactilePushEnv needs the TactilePush scene (7 dofs, 6 controls, 2 end-effectors)")
        self.sim, self.B, self.device = sim, sim.batch, sim.device
        self.observation_type, self.gradient = observation_type, gradient
        M = sim.ndof_tactile // 3
        if tactile_rows is None:
            ip = sim.scene.sensors[0].image_pos
            tactile_rows = int(ip[:, 0].max()) + 1 if len(ip) and ip.max() > 0 else M
            tactile_cols = M // tactile_rows
        self.tactile_rows, self.tactile_cols = tactile_rows, tactile_cols
        self.dt = sim.options.h * self.frame_skip
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.q_init = sim._q_init[0].clone()
        self.goal = torch.zeros((self.B, 3), dtype=torch.float64, device=self.device)
        self.external_force = torch.zeros((self.B, 2), dtype=torch.float64, device=self.device)
        self.current_step = 0
        self.state_q = None
        self.tactile = None

    def _uniform(self, shape, lo, hi):
        r = torch.rand(shape, generator=self.gen, device=self.device, dtype=torch.float64)
        return lo + (hi - lo) * r

    def _obs(self):
        state = push_observation(self.state_q, self.goal, privilege=self.observation_type == "privilege")
        if self.observation_type in ("privilege", "no_tactile"):
            return state
        tac = self.tactile.reshape(self.B, self.tactile_rows, self.tactile_cols, 3)
        if self.observation_type == "tactile_flatten":
            return torch.cat([state, tac.reshape(self.B, -1)], dim=1)
        return tac.permute(0, 3, 1, 2), state

    def reset(self, q0: Optional[torch.Tensor] = None, goal: Optional[torch.Tensor] = None):
        """q0 [B,7] / goal [B,3]: given initial states / goals instead of the random draws (replay of recorded
        episodes, e.g. of the reference's own environment in tests/test_gpu_reference_callers.py)."""
        B = self.B
        if q0 is None:
            q0 = self.q_init.unsqueeze(0).repeat(B, 1)
            q0[:, 1] = -0.001
            q0[:, 4] = self._uniform((B,), -0.02, 0.02)
        else:
            q0 = q0.to(device=self.device, dtype=torch.float64).clone()
        if goal is None:
            gx = self._uniform((B,), 0.15, 0.25)
            gy = self._uniform((B,), -0.2, 0.2)
            gr = gy * math.pi + self._uniform((B,), -math.pi / 16.0, math.pi / 16.0)
            self.goal = torch.stack([gx, gy, gr], dim=1)
        else:
            self.goal = goal.to(device=self.device, dtype=torch.float64).clone()
        self.sim.set_state_init(q0, torch.zeros_like(q0))
        self.sim.reset(backward_flag=self.gradient)
        self.state_q = q0
        self.tactile = self.sim.get_tactile_force_vector_t()
        self.external_force.zero_()
        self.current_step = 0
        return self._obs()

    def step(self, u: torch.Tensor, external_force: Optional[torch.Tensor] = None) -> Tuple[object, torch.Tensor, bool, dict]:
        """u [B,3] raw policy output (before tanh).  Returns obs, reward [B], done (time limit), info.
        external_force [B,2]: given pushes on the box for this step instead of the random draw."""
        B = self.B
        action = torch.tanh(u.to(torch.float64))
        if external_force is not None:
            self.external_force = external_force.to(device=self.device, dtype=torch.float64)
        elif self.current_step % 10 == 0:
            p = self._uniform((B, 1), 0.0, 1.0)
            self.external_force = torch.where(p < 0.5, self._uniform((B, 2), -1.0, 1.0), torch.zeros((B, 2), dtype=torch.float64, device=self.device))
        robot_action = torch.cat([action, self.external_force, torch.zeros((B, 1), dtype=torch.float64, device=self.device)], dim=1)
        q, var, tactile = StepSimFunction.apply(robot_action, self.frame_skip, self.sim, self.gradient)
        self.state_q, self.tactile = q, tactile
        reward, info = push_reward(q, var, u.to(torch.float64), self.goal)
        info["final_pos_error"] = torch.linalg.norm(q[:, 3:5] - self.goal[:, 0:2], dim=1)
        info["final_rot_error"] = (q[:, 6] - self.goal[:, 2]).abs()
        self.current_step += 1
        return self._obs(), reward, self.current_step >= self.max_episode_steps, info
