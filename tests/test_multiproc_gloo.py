"""N > 1 path on CPU: world_size-2 gloo.  Environments shard contiguously with no data-path collective;
the only exchange is one all-reduce of the (policy) gradient.  The per-shard gradients come from the
kernel arithmetic compiled for the host (tests/emu), so the test also checks that a sharded run
reproduces the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import GOLDEN


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs(B, T, seed=3):
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz"))
    rng = np.random.default_rng(seed)
    q0 = np.tile(g["q0"], (B, 1))
    q0[:, 1] = 0.0005
    q0[:, 4] = rng.uniform(-0.02, 0.02, B)
    u = np.zeros((T, B, 6))
    u[:, :, :3] = np.tanh(rng.normal(size=(T, B, 3)))
    u[:, :, 0] = 0.8
    return g, q0, u


def _shard_grad(g, q0, u, lo, hi):
    from tests import emu_lib
    T = u.shape[0]
    out = emu_lib.forward(g["ibuf"], g["dbuf"], q0[lo:hi], np.zeros((hi - lo, 7)), u[:, lo:hi], grad=True)
    dq = np.ones((T, hi - lo, 7))
    dv = np.ones((T, hi - lo, 6))
    dt = np.full((T, hi - lo, 390), 1e-3)
    bw = emu_lib.backward(g["ibuf"], g["dbuf"], out, u[:, lo:hi], dq, dv, dt)
    # a stand-in "policy gradient": df_du contracted with a fixed linear policy Jacobian
    W = np.linspace(-1.0, 1.0, 6 * 5).reshape(6, 5)
    return np.einsum("tbu,up->p", bw["df_du"], W)


def _worker(rank, world, port, B, T, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tactilesimulation_b200.distributed import allreduce_gradients, shard_range
    g, q0, u = _inputs(B, T)
    lo, hi = shard_range(B, rank, world)
    p = torch.nn.Parameter(torch.zeros(5, dtype=torch.float64))
    p.grad = torch.tensor(_shard_grad(g, q0, u, lo, hi))
    extra = allreduce_gradients([p], global_batch=B, extra=torch.tensor([float(hi - lo)], dtype=torch.float64))
    if rank == 0:
        ret["grad"] = p.grad.numpy().copy()
        ret["count"] = float(extra[0])
    dist.destroy_process_group()


def test_shard_ranges_partition_the_batch():
    from tactilesimulation_b200.distributed import shard_range
    for B, W in [(4096, 8), (10, 4), (7, 2), (3, 8)]:
        spans = [shard_range(B, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_allreduce_matches_single_process():
    B, T, world = 4, 6, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, T, ret), nprocs=world, join=True)
    g, q0, u = _inputs(B, T)
    ref = _shard_grad(g, q0, u, 0, B) / B
    assert ret["count"] == B
    assert np.allclose(ret["grad"], ref, rtol=1e-12, atol=1e-14)
