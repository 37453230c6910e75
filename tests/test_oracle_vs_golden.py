"""Pins the numpy oracle (oracle/redmax_oracle.py) against golden vectors produced by the
unmodified reference C++ (tests/golden/make_golden.py).  CPU only.

Tolerances (print_error norm, DH/Utils.h:315-319): q, qdot, var <= 1e-9; tactile <= 1e-8
(north_star allows 1e-4); contact index sets identical; gradients <= 1e-6.
"""
import os

import numpy as np
import pytest

from oracle.redmax_oracle import OracleSim
from tests.blob_scene import scene_from_blob
from tests.conftest import GOLDEN, rel_err


def _ids(row):
    return [int(x) for x in row if x >= 0]


@pytest.mark.parametrize("name,T", [("pusher13x10_episodic_s0", 24), ("pusher13x10_episodic_s1", 14),
                                    ("pusher32x13_episodic_s0", 19)])
def test_forward_matches_reference(name, T):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = scene_from_blob(g["ibuf"], g["dbuf"])
    o = OracleSim(sc)
    o.set_state_init(g["q0"], g["qd0"])
    o.reset(False)
    for t in range(T):
        o.set_u(g["u"][t])
        o.forward(1)
        assert rel_err(o.get_q(), g["q"][t]) <= 1e-9
        assert rel_err(o.get_qdot(), g["qd"][t]) <= 1e-9
        assert rel_err(o.get_variables(), g["var"][t]) <= 1e-9
        if t % 4 == 3 or t == T - 1:
            assert rel_err(o.get_tactile_force_vector(), g["tactile"][t]) <= 1e-8
            cs = o.contact_sets()
            assert cs["ground"][0] == _ids(g["ground_ids"][t])
            assert cs["gp"][0] == _ids(g["gp_ids"][t])
            assert np.array_equal(np.array(cs["marker_body"][0]), g["marker_body"][t])


def test_backward_matches_reference():
    """Simulation::backward (EpisodicSimFunction pattern): full-trajectory adjoint."""
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz"))
    sc = scene_from_blob(g["ibuf"], g["dbuf"])
    T = 22   # a prefix is itself a valid trajectory: re-run the reference cotangents on it
    # the golden gradients are for T=60; compare on the full horizon only in the slow test below,
    # here check self-consistency of adjoint vs finite differences on the prefix.
    o = OracleSim(sc)
    o.set_state_init(g["q0"], g["qd0"])
    o.reset(True)
    for t in range(T):
        o.set_u(g["u"][t])
        o.forward(1)
    res = o.backward(g["df_dq"][:T], g["df_dvar"][:T], g["df_dtactile"][:T])

    def loss(u):
        s = OracleSim(sc)
        s.set_state_init(g["q0"], g["qd0"])
        s.reset(False)
        L = 0.0
        for t in range(T):
            s.set_u(u[t])
            s.forward(1)
            L += g["df_dq"][t] @ s.get_q() + g["df_dvar"][t] @ s.get_variables() + g["df_dtactile"][t] @ s.get_tactile_force_vector()
        return L
    u = g["u"][:T].copy()
    for (t, i) in [(3, 0), (17, 1), (20, 2)]:
        eps = 1e-6
        up, um = u.copy(), u.copy()
        up[t, i] += eps
        um[t, i] -= eps
        fd = (loss(up) - loss(um)) / (2 * eps)
        assert abs(fd - res["df_du"][t, i]) <= 1e-4 * max(abs(fd), 1e-3), (t, i, fd, res["df_du"][t, i])


@pytest.mark.parametrize("name", ["pusher13x10_episodic_s1"])
def test_backward_full_horizon_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = scene_from_blob(g["ibuf"], g["dbuf"])
    T = g["u"].shape[0]
    o = OracleSim(sc)
    o.set_state_init(g["q0"], g["qd0"])
    o.reset(True)
    for t in range(T):
        o.set_u(g["u"][t])
        o.forward(1)
    res = o.backward(g["df_dq"], g["df_dvar"], g["df_dtactile"])
    assert rel_err(res["df_du"], g["df_du"]) <= 1e-6
    assert rel_err(res["df_dq0"], g["df_dq0"]) <= 1e-6
    assert rel_err(res["df_dqdot0"], g["df_dqdot0"]) <= 1e-6


def test_backward_steps_matches_reference():
    """StepSimFunction pattern: forward(5, save_last_frame_var_only) + chained backward_steps(5)."""
    g = np.load(os.path.join(GOLDEN, "pusher13x10_stepsim_s0.npz"))
    sc = scene_from_blob(g["ibuf"], g["dbuf"])
    fs = int(g["frame_skip"])
    ns = g["u"].shape[0]
    o = OracleSim(sc)
    o.set_state_init(g["q0"], g["qd0"])
    o.reset(True)
    for t in range(ns):
        o.set_u(g["u"][t])
        o.forward(fs, save_last_frame_var_only=True)
        assert rel_err(o.get_q(), g["q"][t]) <= 1e-9
        assert rel_err(o.get_tactile_force_vector(), g["tactile"][t]) <= 1e-8
    n, nv, nt = o.n, o.nvar, o.ntac
    for t in range(ns - 1, -1, -1):
        a = np.zeros((fs, n)); a[-1] = g["df_dq"][t]
        b = np.zeros((fs, nv)); b[-1] = g["df_dvar"][t]
        c = np.zeros((fs, nt)); c[-1] = g["df_dtactile"][t]
        r = o.backward_steps(fs, a, b, c)
        assert rel_err(r["df_du"], g["df_du"][t]) <= 1e-6


@pytest.mark.parametrize("name,T", [("BDF2", 24), ("SDIRK2", 20)])
def test_other_integrators_match_reference(name, T):
    """The oracle's restatement of integration_BDF2 (with its SDIRK2 start-up step) and integration_SDIRK2
    (DH/Simulation.cpp:1076-1092, 1353-1564) against the reference on the TactilePush scene."""
    g = np.load(os.path.join(GOLDEN, "pusher13x10_integrators_s0.npz"))
    ib = g["ibuf"].copy()
    ib[14] = {"BDF2": 1, "SDIRK2": 2}[name]
    sc = scene_from_blob(ib, g["dbuf"])
    assert sc.integrator == name
    o = OracleSim(sc)
    o.set_state_init(g["q0"], g["qd0"])
    o.reset(False)
    for t in range(T):
        o.set_u(g["u"][t])
        o.forward(1)
        assert rel_err(o.get_q(), g["q_" + name][t]) <= 1e-9, t
        assert rel_err(o.get_qdot(), g["qd_" + name][t]) <= 1e-9, t
        if t % 4 == 3 or t == T - 1:
            assert rel_err(o.get_tactile_force_vector(), g["tactile_" + name][t]) <= 1e-8, t
            assert o.contact_sets()["gp"][0] == _ids(g["gp_ids_" + name][t]), t
    with pytest.raises(RuntimeError):
        o.reset(True)
        o.forward(1)
