"""Batched DClaw rotate-cap front-end: the task logic of ``R/envs/dclaw_rotate_env.py`` (BASELINE configs[3]) as batched
torch ops on the device: every environment of the batch has its OWN cap (joint damping, radius, end-effector position,
joint location drawn per reset -- the reference randomises one Simulation object per environment, :162-178; here the
per-environment ``update_*`` of the drop-in ``Simulation`` put all of them in one batch).

Semantics kept from the reference (line numbers of ``R/envs/dclaw_rotate_env.py``):
  * reset (:162-190): ``q_init`` with the fingers at (-0.5, 0.8) plus N(0, 0.05) on the nine joint angles; cap damping
    U(0.01, 0.7), radius U(0.02, 0.08), joint offset U(-0.02, 0.02)^2 -> ``update_joint_damping / update_body_size /
    update_endeffector_position / update_joint_location``; forward-only;
  * step (:192-221): action clipped to [-1, 1]; position control: relative targets ``q[:9] + action * 0.06`` clipped to
    the joint limits (or absolute, scaled into the limits); torque control: the action itself; ``frame_skip = 5`` sim-steps;
  * observation (:92-122): the nine joint angles, the three fingertip positions (variables) and the three 20 x 20
    tactile flow images (``get_tactile_flow_images``, DH/Robot.cpp:372-387), flattened ("tactile_flatten") or as
    [9, 20, 20] maps ("tactile"); "no_tactile": joint angles + fingertip positions;
  * reward (:124-160): -0.5 per finger whose summed tactile force norm (of the PREVIOUS observation: the reference
    refreshes its tactile buffer after the reward) is below 1, -(min(cap angle - pi/4, 0))^2,
    -0.005 |action|^2, -50 and done when a fingertip rises above the cap's top surface, +50, success and done at pi/4.
"""
import math
from typing import Optional

import numpy as np
import torch

from ..redmax import Simulation

DOF_LIMIT = [[-0.45, 1.35], [-2.0, 2.0], [1.0, 2.0]] * 3


class BatchedDClawRotateEnv:
    frame_skip = 5
    max_episode_steps = 200       # R/envs/__init__.py:15-19
    tactile_rows = tactile_cols = 20

    def __init__(self, sim: Simulation, observation_type: str = "tactile", torque_control: bool = False,
                 relative_control: bool = True, domain_randomization: bool = True, seed: int = 0):
        if observation_type not in ("tactile", "tactile_flatten", "no_tactile"):
            raise NotImplementedError(observation_type)
        if sim.ndof_r != 10 or sim.ndof_u != 9 or sim.ndof_var != 12:
            raise ValueError("BatchedDClawRotateEnv needs a DClaw scene (10 dofs, 9 controls, 4 end-effectors)")
        self.sim, self.B, self.device = sim, sim.batch, sim.device
        dev, f64 = self.device, torch.float64
        self.observation_type, self.is_torque_control = observation_type, torque_control
        self.relative_control, self.domain_randomization = relative_control, domain_randomization
        self.relative_q_scale, self.rot_coef, self.power_coef = 0.06, 1.0, 0.005
        self.dof_limit = torch.tensor(DOF_LIMIT, dtype=f64, device=dev)
        self.cap_top_surface_z = 0.05
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        q_init = sim._q_init[0].detach().clone()
        q_init[[1, 4, 7]] = -0.5
        q_init[[2, 5, 8]] = 0.8
        self.q_init = q_init
        # marker -> pixel of each 20 x 20 flow image (get_tactile_image_pos)
        # (several markers of the fingertip spec share a pixel -- 302 markers on 182 pixels: the reference writes them in
        # marker order, the last one stays, DH/Robot.cpp:383-384)
        self._pix, self._nmark = [], []
        for s_ in sim.scene.sensors:
            ip = np.asarray(s_.image_pos, dtype=np.int64).reshape(-1, 2)
            last = {}
            for j, (r, c) in enumerate(ip):
                last[int(r) * self.tactile_cols + int(c)] = j
            pix = np.array(sorted(last), dtype=np.int64)
            self._pix.append((torch.as_tensor(pix, device=dev), torch.as_tensor(np.array([last[p_] for p_ in pix], dtype=np.int64), device=dev)))
            self._nmark.append(len(ip))
        self.tactile_force_buf = torch.zeros((self.B, 3, self.tactile_rows, self.tactile_cols, 3), dtype=f64, device=dev)
        self.energy_usage = torch.zeros(self.B, dtype=f64, device=dev)

    def _uniform(self, shape, lo, hi):
        return lo + (hi - lo) * torch.rand(shape, generator=self.gen, device=self.device, dtype=torch.float64)

    def _flow_images(self, tactile):
        """[B, 3M] marker-major forces -> [B, sensors, 20, 20, 3] images (zero where a sensor has no marker)."""
        B = tactile.shape[0]
        out = torch.zeros((B, len(self._pix), self.tactile_rows * self.tactile_cols, 3), dtype=tactile.dtype, device=tactile.device)
        off = 0
        for k, (pix, marker) in enumerate(self._pix):
            M = self._nmark[k]
            out[:, k, pix] = tactile[:, off:off + 3 * M].reshape(B, M, 3)[:, marker]
            off += 3 * M
        return out.reshape(B, len(self._pix), self.tactile_rows, self.tactile_cols, 3)

    def _get_obs(self):
        q = self.sim.get_q_t()
        var = self.sim.get_variables_t()
        state = torch.cat([q[:, :9], var[:, :9]], dim=1)
        if self.observation_type == "no_tactile":
            return state
        self.tactile_force_buf = self._flow_images(self.sim.get_tactile_force_vector_t())
        obs = self.tactile_force_buf
        if self.observation_type == "tactile":
            obs = obs.permute(0, 1, 4, 2, 3)
        return torch.cat([state, obs.reshape(self.B, -1)], dim=1)

    def reset(self, q_noise: Optional[torch.Tensor] = None, damping=None, radius=None, dxy=None):
        """Arguments: given draws (q_noise [B,9], damping [B], radius [B], dxy [B,2]) instead of the random ones."""
        B, dev, f64 = self.B, self.device, torch.float64
        q0 = self.q_init.unsqueeze(0).repeat(B, 1)
        if q_noise is None:
            q_noise = torch.randn((B, 9), generator=self.gen, device=dev, dtype=f64) * 0.05
        q0[:, :9] += torch.as_tensor(q_noise, dtype=f64, device=dev)
        if self.domain_randomization:
            damping = self._uniform((B,), 0.01, 0.7) if damping is None else torch.as_tensor(damping, dtype=f64)
            radius = self._uniform((B,), 0.02, 0.08) if radius is None else torch.as_tensor(radius, dtype=f64)
            dxy = self._uniform((B, 2), -0.02, 0.02) if dxy is None else torch.as_tensor(dxy, dtype=f64)
            damping, radius, dxy = damping.cpu().numpy(), radius.cpu().numpy(), dxy.cpu().numpy()
            self.sim.update_joint_damping("cap", damping)
            self.sim.update_body_size("cap", np.stack([np.full(B, 0.03), radius], axis=1))
            self.sim.update_endeffector_position("cap", np.stack([radius, np.zeros(B), np.zeros(B)], axis=1))
            self.sim.update_joint_location("cap", np.concatenate([dxy, np.full((B, 1), 0.075)], axis=1))
        self.sim.set_state_init(q0, torch.zeros_like(q0))
        self.sim.reset(backward_flag=False)
        self.energy_usage.zero_()
        return self._get_obs()

    def step(self, u: torch.Tensor):
        """u [B,9].  Returns obs, reward [B], done [B] (bool), info (success [B])."""
        u = u.detach().to(device=self.device, dtype=torch.float64)
        action = torch.clip(u, -1.0, 1.0)
        if not self.is_torque_control:
            lo, hi = self.dof_limit[:, 0], self.dof_limit[:, 1]
            if self.relative_control:
                action = torch.maximum(torch.minimum(self.sim.get_q_t()[:, :9] + action * self.relative_q_scale, hi), lo)
            else:
                action = 0.5 * (action + 1.0) * (hi - lo) + lo
        self.sim.forward_t(self.frame_skip, action.contiguous())
        # (as in the reference, :208-221: the reward is computed BEFORE the observation refreshes the tactile buffer, so its
        # contact term looks at the tactile images of the previous observation)
        reward, done, success = self._get_reward(u)
        obs = self._get_obs()
        return obs, reward, done, dict(success=success)

    def _get_reward(self, action):
        q = self.sim.get_q_t()
        tips = self.sim.get_variables_t()[:, :9]
        cap_angle = q[:, -1]
        finger_force = self.tactile_force_buf.norm(dim=-1).sum(dim=(-1, -2))            # [B, 3]
        reward = -(finger_force < 1.0).sum(dim=1).to(torch.float64) * 0.5
        max_angle = math.pi / 4
        reward = reward - self.rot_coef * torch.clamp(cap_angle - max_angle, max=0.0) ** 2
        reward = reward - self.power_coef * (action ** 2).sum(dim=1)
        above = (tips[:, 2::3] > self.cap_top_surface_z).any(dim=1)
        reward = reward - 50.0 * above.to(torch.float64)
        success = cap_angle > max_angle
        reward = reward + 50.0 * success.to(torch.float64)
        return reward, above | success, success
