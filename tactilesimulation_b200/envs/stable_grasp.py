"""Batched StableGrasp front-end: the task logic of ``R/envs/stable_grasp_env.py`` as batched torch ops on the device.
Every environment of the batch holds its OWN bar: the densities of the eleven boxes are drawn per environment at reset
(the reference calls ``update_body_density`` on one Simulation per environment, :66-128; here the per-environment
``update_*`` of the drop-in ``Simulation`` put all bars in one batch).

Semantics kept from the reference (line numbers of ``R/envs/stable_grasp_env.py``):
  * ``generate_initial_state`` (:157-178): gripper at the grasp height, fingers open, 500 sim-steps with the feed-forward
    term; run once for the batch;
  * ``reset`` (:66-141): centre of mass uniform along the bar -> densities of the left / middle / right blocks as in the
    reference's construction, total density clipped to [3000, 7000]; first grasp at position 0;
  * ``step`` (:143-165): ``grasp_position += clip(u, -1, 1) * 0.05`` clipped to +-0.11, one grasp;
  * ``grasp`` (:189-262): 180 sim-steps of position targets (move, close, lift, hold, put down, open), tactile frame 60
    (the lifted bar), shear field normalised per environment, success when the bar's rotation vector at the capture frame
    is below 0.02 rad and it left the ground; reward 100 / -10 |angle|; the next grasp starts from the last state.
"""
from typing import Optional

import numpy as np
import torch

from ..redmax import Simulation
from ..torch_functions import EpisodicSimFunction

BOX_IDS = [9, 8, 1, 2, 3, 4, 5, 6, 7, 10, 11]          # boxes from one end of the bar to the other (:72)
NUM_STEPS = [20, 10, 50, 20, 50, 10, 20]
CAPTURE_FRAME = 60


def sample_block_densities(rng: np.random.RandomState) -> np.ndarray:
    """The density construction of ``reset`` (:68-113) for ONE environment, with the reference's sequence of draws."""
    density_range, num_blocks = [600.0, 700.0], 11
    com_y = rng.uniform(1, num_blocks - 1, 1)
    num_left = int(com_y[0])
    num_right = num_blocks - 1 - num_left
    left_ratio = com_y - num_left
    mid = rng.uniform(density_range[0], density_range[1], 1)[0]
    if left_ratio < 0.5:
        right_total = rng.uniform(density_range[0] * num_right, density_range[1] * num_right, 1)[0]
        left_total = right_total + (1 - left_ratio * 2) * mid
    else:
        left_total = rng.uniform(density_range[0] * num_left, density_range[1] * num_left, 1)[0]
        right_total = left_total + (left_ratio * 2 - 1) * mid
    lr = rng.random_sample(num_left) + 0.1
    lr /= lr.sum()
    rr = rng.random_sample(num_right) + 0.1
    rr /= rr.sum()
    dens = (np.asarray(left_total) * lr).reshape(-1).tolist()
    if left_ratio > 0:
        dens.append(mid)
    dens.extend((np.asarray(right_total) * rr).reshape(-1).tolist())
    dens = np.array(dens, dtype=np.float64)
    return dens / dens.sum() * np.clip(dens.sum(), 3000, 7000)


class BatchedStableGraspEnv:
    max_episode_steps = 10        # R/envs/__init__.py:3-7
    tactile_rows, tactile_cols = 13, 10

    def __init__(self, sim: Simulation, observation_type: str = "tactile_map", seed: int = 0):
        if observation_type not in ("tactile_map", "tactile_flatten"):
            raise NotImplementedError(observation_type)
        if sim.ndof_r != 12 or sim.ndof_u != 6 or sim.ndof_tactile != 780:
            raise ValueError("BatchedStableGraspEnv needs the StableGrasp scene (12 dofs, 6 controls, two 13x10 pads)")
        self.sim, self.B, self.device = sim, sim.batch, sim.device
        self.observation_type = observation_type
        self.action_scale, self.grasp_position_bound = 0.05, 0.11
        self.rng = [np.random.RandomState(seed + e) for e in range(self.B)]
        self.qpos_init_reference, self.qvel_init_reference = self.generate_initial_state()
        self.grasp_position = torch.zeros(self.B, dtype=torch.float64, device=self.device)
        self.current_q = None
        self.block_densitys = None
        self.obs_buf = self.reward_buf = self.done_buf = self.is_success = None

    def generate_initial_state(self):
        B, dev, f64 = self.B, self.device, torch.float64
        q = self.sim._q_init.detach().clone()
        q[:, 2], q[:, 4], q[:, 5] = 0.2, -0.03, -0.03
        u = q[:, 0:6].clone()
        u[:, 2] += 0.003                                   # feed-forward term
        masks = torch.zeros(500, dtype=torch.bool)
        masks[-1] = True
        qs, _, _ = EpisodicSimFunction.apply(q, torch.zeros_like(q), u.unsqueeze(0).expand(500, B, 6).contiguous(), masks, self.sim, False)
        return qs[-1].clone(), self.sim.get_qdot_t()

    def reset(self, block_densitys: Optional[np.ndarray] = None):
        """block_densitys [B, 11]: given densities (bar order) instead of the per-environment draws."""
        B = self.B
        if block_densitys is None:
            block_densitys = np.stack([sample_block_densities(r) for r in self.rng])
        self.block_densitys = np.asarray(block_densitys, dtype=np.float64).reshape(B, 11)
        for idx, box_id in enumerate(BOX_IDS):
            self.sim.update_body_density("box_%d" % box_id, self.block_densitys[:, idx])
        self.grasp_position = torch.zeros(B, dtype=torch.float64, device=self.device)
        self.current_q = self.qpos_init_reference.clone()
        self.sim.clearBackwardCache()
        self.grasp()
        return self.obs_buf

    def step(self, u: torch.Tensor):
        """u [B,1].  Returns obs, reward [B], done [B] (bool), info (success [B])."""
        a = torch.clip(u.detach().to(device=self.device, dtype=torch.float64).reshape(self.B, -1), -1.0, 1.0) * self.action_scale
        self.grasp_position = torch.clip(self.grasp_position + a[:, 0], -self.grasp_position_bound, self.grasp_position_bound)
        self.grasp()
        return self.obs_buf, self.reward_buf, self.done_buf, dict(success=self.is_success)

    def grasp(self):
        B, dev, f64 = self.B, self.device, torch.float64
        lift_h, grasp_h, finger = 0.2029862 + 0.03, 0.2029862, -0.008
        q0 = self.current_q.clone()
        q0[:, 1] = self.grasp_position
        gp, z = self.grasp_position, torch.zeros(B, dtype=f64, device=dev)

        def tq(h, f4, f5):
            return torch.stack([z, gp, torch.full_like(z, h), z, f4, f5], dim=1)
        fz = torch.full_like(z, finger)
        targets = [q0[:, :6], tq(grasp_h, fz, fz), tq(grasp_h, fz, fz), tq(lift_h, fz, fz), tq(lift_h, fz, fz),
                   tq(grasp_h, fz, fz), tq(grasp_h, fz, fz), tq(grasp_h, q0[:, 4], q0[:, 5])]
        acts = []
        for stage, ns in enumerate(NUM_STEPS):
            for i in range(ns):
                acts.append((targets[stage + 1] - targets[stage]) / ns * (i + 1) + targets[stage])
        actions = torch.stack(acts)                                           # [180, B, 6]
        masks = torch.zeros(actions.shape[0], dtype=torch.bool)
        masks[CAPTURE_FRAME] = True
        qs, _, tactiles = EpisodicSimFunction.apply(q0, torch.zeros_like(q0), actions.contiguous(), masks, self.sim, False)
        tf = tactiles.reshape(1, B, 2, self.tactile_rows, self.tactile_cols, 3)[..., 0:2].permute(1, 0, 2, 3, 4, 5).contiguous()
        self.tactile_force_buf = tf                                           # [B, 1, 2, rows, cols, 2]
        mx = tf.norm(dim=-1).reshape(B, -1).max(dim=1).values + 1e-5
        obs = tf / (mx / 30.0).view(B, 1, 1, 1, 1, 1)
        if self.observation_type == "tactile_flatten":
            self.obs_buf = obs.reshape(B, -1)
        else:
            self.obs_buf = obs.permute(0, 1, 2, 5, 3, 4).reshape(B, -1, self.tactile_rows, self.tactile_cols)
        abs_angle = qs[CAPTURE_FRAME, :, 9:12].norm(dim=1)
        success = (abs_angle < 0.02) & (qs[CAPTURE_FRAME, :, -4] > 0.005)
        self.reward_buf = torch.where(success, torch.full_like(abs_angle, 100.0), -abs_angle * 10.0)
        self.done_buf, self.is_success = success, success
        self.current_q = qs[-1].clone()
