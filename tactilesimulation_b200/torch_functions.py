"""Batched twins of the reference's plugin seam ``R/envs/redmax_torch_functions.py``.

Same class names and argument lists as the reference ``torch.autograd.Function``s, with a
leading env dimension and device-resident tensors: the ``tensor -> numpy -> C++ -> numpy ->
tensor`` round trip of the reference (``redmax_torch_functions.py:35-37,129``) does not exist
here -- q, var, tactile and all cotangents stay in HBM, one kernel launch per call.

* ``StepSimFunction.apply(action[B,nu], num_steps, sim, grad_mode) -> q[B,n], var[B,nvar],
  tactile[B,ntac]`` (reference ``:112-174``).  The gradient w.r.t. ``action`` is already summed
  over the sub-steps (the reference returns ``(num_steps, ndof_u)`` and lets autograd's broadcast
  reduction sum it).
* ``EpisodicSimFunction.apply(q0[B,n], qdot0[B,n], actions[T,B,nu], tactile_masks[T], sim,
  grad_mode) -> qs[T,B,n], vars[T,B,nvar], tactiles[Tm,B,ntac]`` (reference ``:11-109``).  As in
  the reference, gradients flow through every step's tactile derivative; in grad mode the masks
  must therefore select every step (the reference's C++ ignores masks in backward and its python
  raises on the size mismatch, ``:69-88``).

``sim`` is a :class:`tactilesimulation_b200.redmax.Simulation` created with ``batch=B``.
"""
from typing import Any, Tuple

import torch
import torch.autograd as autograd

from ._lib import TactileSimError
from .redmax import Simulation


class StepSimFunction(autograd.Function):

    @staticmethod
    def forward(ctx: Any, action: torch.Tensor, num_steps: int, sim: Simulation,
                grad_mode: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        ctx.sim = sim
        ctx.num_steps = num_steps
        ctx.grad_actions = action.requires_grad
        ctx.in_dtype = action.dtype
        u = action.detach().to(device=sim.device, dtype=torch.float64).contiguous()
        if u.dim() == 1:
            u = u.unsqueeze(0).expand(sim.batch, -1).contiguous()
        sim._u = u
        out = sim.forward_t(num_steps, u, save_last_frame_var_only=True, want_outputs=True)
        q = sim.get_q_t()
        var = out["var"][0] if out["var"] is not None else q.new_zeros((sim.batch, 0))
        tactile = out["tactile"][0] if out["tactile"] is not None else q.new_zeros((sim.batch, 0))
        return q.to(ctx.in_dtype), var.to(ctx.in_dtype), tactile.to(ctx.in_dtype)

    @staticmethod
    def backward(ctx: Any, df_dq: torch.Tensor, df_dvar: torch.Tensor, df_dtactile: torch.Tensor):
        sim, ns = ctx.sim, ctx.num_steps
        dev = sim.device

        def cot(c, width):
            if width == 0 or c is None:
                return None
            return c.detach().to(device=dev, dtype=torch.float64).reshape(sim.batch, width)
        # cotangents live on the last sub-step only: the row maps of the kernel say so (no zero-padded [ns, B, 3M] tensors)
        df_du = sim.backward_last_frame_t(ns, cot(df_dq, sim.ndof_r), cot(df_dvar, sim.ndof_var), cot(df_dtactile, sim.ndof_tactile))
        if not ctx.grad_actions:
            return None, None, None, None
        return df_du.sum(dim=0).to(ctx.in_dtype), None, None, None


class EpisodicSimFunction(autograd.Function):

    @staticmethod
    def forward(ctx: Any, q0: torch.Tensor, qdot0: torch.Tensor, actions: torch.Tensor,
                tactile_masks: torch.Tensor, sim: Simulation,
                grad_mode: bool) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        T = actions.shape[0]
        ctx.sim, ctx.T = sim, T
        ctx.grad_q0, ctx.grad_qdot0, ctx.grad_actions = q0.requires_grad, qdot0.requires_grad, actions.requires_grad
        ctx.in_dtype = q0.dtype
        masks = [bool(m) for m in (tactile_masks.tolist() if isinstance(tactile_masks, torch.Tensor) else tactile_masks)]
        if len(masks) != T:
            raise TactileSimError("tactile_masks must have one entry per step")
        if grad_mode and not all(masks):
            raise TactileSimError("grad_mode needs tactile on every step (the reference's backward ignores masks "
                                  "and rejects the shorter cotangent, redmax_torch_functions.py:69-88)")
        rows, k = [], 0
        for m in masks:
            rows.append(k if m else -1)
            k += 1 if m else 0
        sim.set_state_init(q0.detach(), qdot0.detach())
        sim.reset(backward_flag=grad_mode)
        u = actions.detach().to(device=sim.device, dtype=torch.float64)
        if u.dim() == 2:
            u = u.unsqueeze(1).expand(T, sim.batch, -1)
        u = u.contiguous()
        out = sim.forward_t(T, u, save_last_frame_var_only=False, tac_rows=rows, want_outputs=True)
        if grad_mode:
            sim.saveBackwardCache()
        qs = out["q_traj"]
        vars_ = out["var"] if out["var"] is not None else qs.new_zeros((T, sim.batch, 0))
        tac = out["tactile"] if out["tactile"] is not None else qs.new_zeros((k, sim.batch, 0))
        return qs.to(ctx.in_dtype), vars_.to(ctx.in_dtype), tac.to(ctx.in_dtype)

    @staticmethod
    def backward(ctx: Any, df_dq: torch.Tensor, df_dvar: torch.Tensor, df_dtactile: torch.Tensor):
        sim, T = ctx.sim, ctx.T
        dev = sim.device
        sim.popBackwardCache()

        def cot(c, width):
            if width == 0 or c is None:
                return None
            return c.detach().to(device=dev, dtype=torch.float64).reshape(T, sim.batch, width).contiguous()
        df_du, dq0, dqd0 = sim.backward_t(cot(df_dq, sim.ndof_r), cot(df_dvar, sim.ndof_var),
                                          cot(df_dtactile, sim.ndof_tactile), want_q0=True)
        return (dq0.to(ctx.in_dtype) if ctx.grad_q0 else None,
                dqd0.to(ctx.in_dtype) if ctx.grad_qdot0 else None,
                df_du.to(ctx.in_dtype) if ctx.grad_actions else None, None, None, None)
