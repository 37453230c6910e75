"""Multi-GPU plumbing: the environment batch shards trivially (environments are independent), so the
only exchange is ONE all-reduce of the policy gradient per outer optimisation step
(``R/algorithms/gd.py:155-160``: between ``compute_reward_and_grad`` and ``clip_grad_norm_``).

One process per GPU (torchrun), ``torch.distributed`` with the NCCL backend on the GPU box; the same code
runs over gloo on CPU for the tests.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_range(n_envs_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of the global environment batch owned by ``rank``."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_envs_global), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_seed(seed: int, rank: int) -> int:
    """Per-rank RNG stream for the randomised initial states / actions."""
    return int(seed) + int(rank)


def allreduce_gradients(params: Iterable[torch.Tensor], global_batch: int | None = None, extra: torch.Tensor | None = None):
    """Sums ``p.grad`` of every parameter over all ranks with ONE collective on a flat buffer (plus an
    optional ``extra`` 1-D tensor of logging scalars riding along), then scales by 1/global_batch if given.
    Returns the reduced ``extra`` (or None)."""
    plist = [p for p in params if p.grad is not None]
    if not plist and extra is None:
        return extra
    flat = [p.grad.reshape(-1) for p in plist]
    if extra is not None:
        flat.append(extra.reshape(-1).to(flat[0].dtype if flat else extra.dtype))
    buf = torch.cat(flat)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    off = 0
    for p in plist:
        k = p.grad.numel()
        g = buf[off:off + k].view_as(p.grad)
        p.grad.copy_(g / global_batch if global_batch else g)
        off += k
    return buf[off:].clone() if extra is not None else None
