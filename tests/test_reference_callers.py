"""The reference's own python callers (staged unmodified under oracle/_ref/py) run on the shims of tests/shims: here,
without a GPU, against the REFERENCE module -- which pins the shims and the checks themselves; the same checks run
against the drop-in module on the GPU box (tests/test_gpu_reference_callers.py)."""
import os

import numpy as np
import pytest

from tests import ref_callers as rc
from tests.conftest import GOLDEN, rel_err

pytestmark = pytest.mark.skipif(not rc.available(), reason="oracle/_ref/py not staged (run oracle/build_ref.sh where /root/reference exists)")
XML = os.path.join(rc.PY_DIR, "envs", "assets", "pusher", "pusher.xml")


@pytest.fixture(scope="module")
def ns():
    return rc.load(rc.reference_module())


def test_unmodified_episodic_function_reproduces_the_golden(ns):
    rc.check_episodic_function(ns, XML, np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz")), rel_err)


def test_unmodified_stepsim_function_reproduces_the_golden(ns):
    rc.check_stepsim_function(ns, XML, np.load(os.path.join(GOLDEN, "pusher13x10_stepsim_s0.npz")), rel_err)


def test_tactile_push_env_and_one_gd_epoch_run(ns, tmp_path):
    obs, rew, grads = rc.run_push_env(ns, steps=10)
    assert obs.shape == (11, 3 + 390) and np.isfinite(obs).all() and np.isfinite(rew).all() and np.isfinite(grads).all()
    assert np.abs(grads).max() > 0
    rewards, lens, gd = rc.run_gd_epoch(ns, str(tmp_path))
    assert lens == [100] and np.isfinite(rewards).all()
    g = np.concatenate([p.grad.reshape(-1).numpy() for p in gd.actor.parameters() if p.grad is not None])
    assert np.isfinite(g).all() and np.abs(g).max() > 0
