"""Batched TactileInsertion front-end: the task logic of ``R/envs/tactile_insertion_env.py`` (BASELINE configs[4]: PPO
rollouts) as batched torch ops on the device, over the batched ``EpisodicSimFunction`` (forward-only, tactile read on the
masked frames).

Semantics kept from the reference (line numbers of ``R/envs/tactile_insertion_env.py``):
  * ``generate_initial_pose`` (:126-170): the gripper grasps the box with the scripted position targets (100 + 100 + 300
    sim-steps), the grasp is lifted by the feed-forward height and held for 500 sim-steps; run ONCE for the batch;
  * ``reset`` (:199-230): per-environment position / rotation / grasp-height noise applied to the grasp pose
    (``apply_relative_motion``, :179-197, with the reference's rotation-vector product on q[9:12]), optional domain
    randomisation (:232-281) -- here through the per-environment ``update_*`` of the drop-in ``Simulation`` --, then one
    insertion attempt;
  * ``step`` (:286-339): clipped, scaled action -> relative / accumulative motion of the pre-insertion pose inside the
    working space, one insertion attempt;
  * ``execute_insertion`` (:343-446): ``execution_num_steps = 45`` sim-steps of a position-target ramp (0.0011 down,
    feed-forward 0.003, grasp force on the fingers); tactile frames ``tactile_masks`` (:75-77), taken relative to the
    reference frame; optional noise / normalisation; success test and the "absolute" / "delta" rewards.
Random draws are torch generators on the device (i.i.d. with the reference's distributions, not the same numbers);
``reset`` / ``step`` accept the draws as arguments, which is how the tests replay episodes of the reference's environment.
"""
import math
from typing import Optional

import torch

from ..redmax import Simulation
from ..torch_functions import EpisodicSimFunction


def rotvec_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Batched ``utils/torch_utils.py:18-37`` (composition of rotation vectors, with its small-angle branches): [B,3]."""
    an, bn = a.norm(dim=1, keepdim=True), b.norm(dim=1, keepdim=True)
    au, bu = a / an.clamp_min(1e-300), b / bn.clamp_min(1e-300)
    sa, ca, sb, cb = torch.sin(an / 2), torch.cos(an / 2), torch.sin(bn / 2), torch.cos(bn / 2)
    cn = 2.0 * torch.arccos((ca * cb - ((au * sa) * (bu * sb)).sum(dim=1, keepdim=True)).clamp(-1.0, 1.0))
    cu = (ca * sb * bu + cb * sa * au + torch.cross(au * sa, bu * sb, dim=1)) / torch.sin(cn / 2).clamp_min(1e-300)
    out = cn * cu
    out = torch.where(cn < 1e-7, torch.zeros_like(out), out)
    out = torch.where(bn < 1e-7, a, out)
    out = torch.where(an < 1e-7, b, out)
    return out


class BatchedTactileInsertionEnv:
    """B TactileInsertion environments on one GPU.  ``sim``: a ``Simulation`` of tactile_insertion.xml with ``batch=B``."""
    execution_num_steps = 45
    max_episode_steps = 15        # R/envs/__init__.py:21-24

    def __init__(self, sim: Simulation, observation_type: str = "tactile_flatten", observation_noise: bool = True,
                 normalize_tactile_obs: bool = True, allow_translation: bool = True, allow_rotation: bool = False,
                 num_obs_frames: int = 5, action_xy_scale: float = 0.02, action_rot_scale: float = math.pi / 18.0,
                 action_type: str = "relative", reward_type: str = "absolute", domain_randomization: bool = False,
                 seed: int = 0, tactile_rows: int = 13, tactile_cols: int = 10):
        if observation_type not in ("tactile_flatten", "tactile_map", "privilege"):
            raise NotImplementedError(observation_type)
        if sim.ndof_r != 12 or sim.ndof_u != 6:
            raise ValueError("BatchedTactileInsertionEnv needs the TactileInsertion scene (12 dofs, 6 controls)")
        self.sim, self.B, self.device = sim, sim.batch, sim.device
        dev, f64 = self.device, torch.float64
        self.observation_type, self.observation_noise = observation_type, observation_noise
        self.normalize_tactile_obs, self.domain_randomization = normalize_tactile_obs, domain_randomization
        self.allow_translation, self.allow_rotation = allow_translation, allow_rotation
        self.action_type, self.reward_type = action_type, reward_type
        self.working_space_boundary = torch.tensor([0.015, 0.015], dtype=f64, device=dev)
        self.working_rotation_boundary = math.pi / 12.0
        self.max_error = torch.tensor([0.006, 0.006, math.pi / 18.0], dtype=f64, device=dev)
        self.grasp_force_range = (1.0 / 8.0, 0.8)
        self.tactile_rows, self.tactile_cols = tactile_rows, tactile_cols
        self.tactile_samples = num_obs_frames
        self.tactile_initial_frame = 15 if observation_type != "tactile_map" else 12
        self.obs_frequency = (self.execution_num_steps - self.tactile_initial_frame) // self.tactile_samples
        masks = torch.zeros(self.execution_num_steps, dtype=torch.bool)
        masks[self.tactile_initial_frame + self.obs_frequency - 1::self.obs_frequency] = True      # observation frames
        masks[6] = True                                                                            # reference frame
        self.tactile_masks = masks
        scales = []
        if allow_translation:
            scales += [action_xy_scale, action_xy_scale]
        if allow_rotation:
            scales += [action_rot_scale]
        if not scales:
            raise ValueError("allow_translation or allow_rotation")
        self.ndof_u = len(scales)
        self.action_scale = torch.tensor(scales, dtype=f64, device=dev)
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.grasp_force = torch.ones(self.B, dtype=f64, device=dev)
        self.q_init_reference, self.qdot_init_reference = self.generate_initial_pose()
        self.current_q_init = self.original_q_init = self.prev_object_pose = None
        self.obs_buf = self.reward_buf = self.done_buf = None
        self.info_buf = {}

    # ------------------------------------------------------------------ helpers
    def _uniform(self, shape, lo, hi):
        r = torch.rand(shape, generator=self.gen, device=self.device, dtype=torch.float64)
        return lo + (hi - lo) * r

    def _rollout(self, q0, qd0, actions, masks, want_tactile=True):
        return EpisodicSimFunction.apply(q0, qd0, actions, masks, self.sim, False)

    def generate_initial_pose(self):
        """:126-170 -- identical for every environment (deterministic), computed with the batch."""
        B, dev, f64 = self.B, self.device, torch.float64
        q_init = self.sim._q_init[0].detach().clone()           # sim.get_q_init() of the reference
        grasp_height, initial_object_height = 0.2, 0.026 + 0.003
        q_init[2], q_init[4], q_init[5] = grasp_height, -0.03, -0.03
        targets = [torch.tensor([q_init[0], q_init[1], q_init[2], q_init[4], 0.0, 0.0], dtype=f64, device=dev),
                   torch.tensor([0.0, 0.0, grasp_height, 0.0, 0.0, 0.0], dtype=f64, device=dev),
                   torch.tensor([0.0, 0.0, grasp_height, 0.0, 1.0, 1.0], dtype=f64, device=dev),
                   torch.tensor([0.0, 0.0, grasp_height, 0.0, 1.0, 1.0], dtype=f64, device=dev)]
        us = []
        for stage, ns in enumerate((100, 100, 300)):
            for i in range(ns):
                us.append((targets[stage + 1] - targets[stage]) / ns * (i + 1) + targets[stage])
        u = torch.stack(us).unsqueeze(1).expand(len(us), B, 6).contiguous()
        none = torch.zeros(len(us), dtype=torch.bool)
        none[-1] = True
        qb = q_init.unsqueeze(0).expand(B, 12).contiguous()
        qs, _, _ = self._rollout(qb, torch.zeros_like(qb), u, none)
        initial_q = qs[-1].clone()
        initial_q[:, 2] += initial_object_height
        initial_q[:, 8] += initial_object_height
        hold = initial_q[:, :6].clone()
        hold[:, 4:6] = 1.0
        u2 = hold.unsqueeze(0).expand(500, B, 6).contiguous()
        m2 = torch.zeros(500, dtype=torch.bool)
        m2[-1] = True
        qs2, _, _ = self._rollout(initial_q, torch.zeros_like(initial_q), u2, m2)
        q_fin = qs2[-1].clone()
        qd_fin = self.sim.get_qdot_t()
        return q_fin, qd_fin

    def apply_relative_motion(self, q, relative_position, relative_rotation, grasp_height_noise=None):
        """:179-197, batched: q [B,12], relative_position [B,2] or [B,3], relative_rotation [B]."""
        new_q = q.clone()
        k = relative_position.shape[1]
        new_q[:, 0:k] += relative_position
        new_q[:, 6:6 + k] += relative_position
        if grasp_height_noise is not None:
            new_q[:, 2] += grasp_height_noise
        new_q[:, 3] = new_q[:, 3] + relative_rotation
        rz = torch.zeros((q.shape[0], 3), dtype=q.dtype, device=q.device)
        rz[:, 2] = relative_rotation
        new_q[:, 9:12] = rotvec_mul(q[:, 9:12], rz)
        return new_q

    def do_domain_randomization(self):
        """:232-281: contact and tactile coefficients of both pads, grasp force -- one draw per environment."""
        B = self.B
        c = [self._uniform((B,), 2e3, 14e3), self._uniform((B,), 20.0, 140.0), self._uniform((B,), 0.5, 2.5), torch.full((B,), 1e3, dtype=torch.float64)]
        t = [self._uniform((B,), 50.0, 450.0), self._uniform((B,), 0.2, 2.3), self._uniform((B,), 0.5, 2.5), self._uniform((B,), 0.0, 100.0)]
        for pad in ("tactile_pad_left", "tactile_pad_right"):
            self.sim.update_contact_parameters(pad, "box", *[x.cpu().numpy() for x in c])
            self.sim.update_tactile_parameters(pad, *[x.cpu().numpy() for x in t])
        self.grasp_force = self._uniform((B,), *self.grasp_force_range)

    # ------------------------------------------------------------------ gym-style interface
    def reset(self, position_noise: Optional[torch.Tensor] = None, rotation_noise: Optional[torch.Tensor] = None,
              grasp_height_noise: Optional[torch.Tensor] = None):
        B, dev, f64 = self.B, self.device, torch.float64
        me = self.max_error
        if position_noise is None:
            if self.allow_translation:
                position_noise = torch.stack([self._uniform((B,), -me[0].item(), me[0].item()), self._uniform((B,), -me[1].item(), me[1].item()),
                                              self._uniform((B,), -0.0002, 0.0002)], dim=1)
            else:
                position_noise = torch.zeros((B, 2), dtype=f64, device=dev)
        if rotation_noise is None:
            rotation_noise = self._uniform((B,), -me[2].item(), me[2].item()) if self.allow_rotation else torch.zeros(B, dtype=f64, device=dev)
        if grasp_height_noise is None:
            grasp_height_noise = self._uniform((B,), -0.01, 0.005)
        to = lambda x: torch.as_tensor(x, dtype=f64, device=dev)
        self.current_q_init = self.apply_relative_motion(self.q_init_reference, to(position_noise), to(rotation_noise), to(grasp_height_noise))
        self.original_q_init = self.current_q_init.clone()
        self.prev_object_pose = torch.stack([self.current_q_init[:, 0], self.current_q_init[:, 1], self.current_q_init[:, 3]], dim=1)
        self.sim.clearBackwardCache()
        if self.domain_randomization:
            self.do_domain_randomization()
        self.execute_insertion()
        return self.obs_buf

    def step(self, u: torch.Tensor):
        """u [B, ndof_u] in [-1, 1] (clipped).  Returns obs, reward [B], done [B] (bool), info."""
        action = torch.clip(u.to(device=self.device, dtype=torch.float64), -1.0, 1.0) * self.action_scale
        base = self.current_q_init if self.action_type == "relative" else self.original_q_init
        idx = 0
        if self.allow_translation:
            xy = action[:, 0:2]
            if self.action_type == "relative":
                xy = torch.maximum(torch.minimum(xy, self.working_space_boundary - self.current_q_init[:, 0:2]),
                                   -self.working_space_boundary - self.current_q_init[:, 0:2])
            idx = 2
        else:
            xy = torch.zeros((self.B, 2), dtype=torch.float64, device=self.device)
        if self.allow_rotation:
            rot = action[:, idx]
            if self.action_type == "relative":
                rot = torch.maximum(torch.minimum(rot, torch.full_like(rot, self.working_rotation_boundary)),
                                    -self.working_rotation_boundary - self.current_q_init[:, 3])
        else:
            rot = torch.zeros(self.B, dtype=torch.float64, device=self.device)
        if self.action_type not in ("relative", "accumulative"):
            raise NotImplementedError(self.action_type)
        self.current_q_init = self.apply_relative_motion(base, xy, rot)
        self.execute_insertion()
        return self.obs_buf, self.reward_buf, self.done_buf, dict(self.info_buf)

    def execute_insertion(self, noise: Optional[torch.Tensor] = None):
        """:343-446: one insertion attempt of every environment from ``current_q_init``."""
        B, T, dev, f64 = self.B, self.execution_num_steps, self.device, torch.float64
        init = self.current_q_init[:, :6]
        target = init.clone()
        target[:, 2] -= 0.0011
        steps = torch.arange(1, T + 1, dtype=f64, device=dev).view(T, 1, 1)
        actions = (target - init).unsqueeze(0) / T * steps + init.unsqueeze(0)
        actions[:, :, 2] += 0.003                                   # feed-forward term
        actions[:, :, 4] = self.grasp_force
        actions[:, :, 5] = self.grasp_force
        qs, _, tactiles = self._rollout(self.current_q_init, torch.zeros_like(self.current_q_init), actions.contiguous(), self.tactile_masks)
        tactiles = tactiles - tactiles[0:1]                          # relative to the reference frame
        tactiles = tactiles[1:]                                      # [samples, B, 2 * rows * cols * 3]
        S_ = self.tactile_samples
        tf = tactiles.reshape(S_, B, 2, self.tactile_rows, self.tactile_cols, 3)[..., 0:2].permute(1, 0, 2, 3, 4, 5).contiguous()
        self.tactile_force_buf = tf                                  # [B, samples, 2, rows, cols, 2]
        obs = tf.clone()
        if self.observation_noise:
            if noise is None:
                noise = torch.randn(obs.shape, generator=self.gen, device=dev, dtype=f64) * 0.00001
            obs = obs + noise
        if self.normalize_tactile_obs:
            mx = obs.norm(dim=-1).reshape(B, -1).max(dim=1).values + 1e-5
            obs = obs / (mx / 30.0).view(B, 1, 1, 1, 1, 1)
        if self.observation_type == "tactile_flatten":
            self.obs_buf = obs.reshape(B, -1)
        elif self.observation_type == "tactile_map":
            self.obs_buf = obs.permute(0, 1, 2, 5, 3, 4).reshape(B, -1, self.tactile_rows, self.tactile_cols)
        else:
            self.obs_buf = torch.stack([self.current_q_init[:, 0], self.current_q_init[:, 1]], dim=1)
        cur = torch.stack([self.current_q_init[:, 0], self.current_q_init[:, 1], self.current_q_init[:, 3]], dim=1)
        self.current_object_pose = cur
        if not self.allow_rotation:
            success = (qs[-1, :, 6].abs() <= 0.0022) & (qs[-1, :, 7].abs() <= 0.0022)
        else:
            success = qs[-1, :, 8] < 0.0247
        improve = (self.prev_object_pose / self.max_error).norm(dim=1) > (cur / self.max_error).norm(dim=1)
        if self.reward_type == "absolute":
            reward = -(self.current_q_init[:, 0:2] ** 2).sum(dim=1) * 10000 - (self.current_q_init[:, 3] ** 2) * 20.0
        elif self.reward_type == "delta":
            reward = ((self.prev_object_pose / self.max_error).norm(dim=1) - (cur / self.max_error).norm(dim=1)) * 10.0
            reward = reward + torch.where(success, torch.full_like(reward, 20.0), torch.full_like(reward, -1.0))
        else:
            raise NotImplementedError(self.reward_type)
        self.info_buf = dict(prev_object_pose=self.prev_object_pose, new_object_pose=cur, improve=improve, success=success)
        self.prev_object_pose = cur
        self.reward_buf, self.done_buf = reward, success
        self.last_qs = qs
