"""CPU-side check of the KERNEL ARITHMETIC: the per-lane math that the sm_100a kernels inline
(tactilesimulation_b200/csrc/sim_core.cuh) is compiled by g++ into a test-only harness
(tests/emu) and compared with the golden vectors of the reference.  This is development/test
infrastructure, not a product path (the product has no CPU fallback); the GPU parity tests in
test_gpu_parity.py are the parity tests proper.

Tolerances as in test_gpu_parity.py."""
import os

import numpy as np
import pytest

from tests import emu_lib
from tests.conftest import GOLDEN, rel_err


@pytest.mark.parametrize("name", ["pusher13x10_episodic_s0", "pusher13x10_episodic_s1", "pusher32x13_episodic_s0"])
def test_emulated_kernel_math_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    T = g["u"].shape[0]
    out = emu_lib.forward(g["ibuf"], g["dbuf"], g["q0"], g["qd0"], g["u"][:, None, :], grad=True)
    assert int((out["status"] >> 16).max()) == 0
    # Newton iterations | line-search evaluations << 8 equal the counters of the reference's own newton()
    # (instrumented probe of oracle/build_ref.sh, DH/Simulation.cpp:1171,1189)
    assert np.array_equal(out["status"][:, 0] & 0xff, g["newton"][:, 0])
    assert np.array_equal((out["status"][:, 0] >> 8) & 0xff, g["newton"][:, 1])
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["q"][t]) <= 1e-9
        assert rel_err(out["qd"][t, 0], g["qd"][t]) <= 1e-9
        assert rel_err(out["var"][t, 0], g["var"][t]) <= 1e-9
        assert rel_err(out["tactile"][t, 0], g["tactile"][t]) <= 1e-8
        assert emu_lib.mask_to_ids(out["cmask"][t, 0, 0:1]) == [int(x) for x in g["ground_ids"][t] if x >= 0]
        assert emu_lib.mask_to_ids(out["cmask"][t, 0, 1:4]) == [int(x) for x in g["gp_ids"][t] if x >= 0]
        assert np.array_equal(out["marker_body"][t, 0], g["marker_body"][t])
    bw = emu_lib.backward(g["ibuf"], g["dbuf"], out, g["u"][:, None, :], g["df_dq"][:, None, :],
                          g["df_dvar"][:, None, :], g["df_dtactile"][:, None, :])
    assert rel_err(bw["df_du"][:, 0], g["df_du"]) <= 1e-6
    assert rel_err(bw["df_dq0"][0], g["df_dq0"]) <= 1e-6
    assert rel_err(bw["df_dqdot0"][0], g["df_dqdot0"]) <= 1e-6


def test_emulated_stepsim_chain_matches_reference():
    g = np.load(os.path.join(GOLDEN, "pusher13x10_stepsim_s0.npz"))
    fs, ns = int(g["frame_skip"]), g["u"].shape[0]
    rows = [-1] * (fs - 1) + [0]
    q, qd = g["q0"].copy(), g["qd0"].copy()
    fwds, us = [], []
    for t in range(ns):
        u = np.tile(g["u"][t], (fs, 1, 1))
        o = emu_lib.forward(g["ibuf"], g["dbuf"], q, qd, u, grad=True, var_row=rows, tac_row=rows)
        q, qd = o["q_final"][0], o["qd_final"][0]
        fwds.append(o)
        us.append(u)
        assert rel_err(q, g["q"][t]) <= 1e-9
        assert rel_err(o["var"][0, 0], g["var"][t]) <= 1e-9
        assert rel_err(o["tactile"][0, 0], g["tactile"][t]) <= 1e-8
    carry = None
    for t in range(ns - 1, -1, -1):
        bw = emu_lib.backward(g["ibuf"], g["dbuf"], fwds[t], us[t], g["df_dq"][t][None, None], g["df_dvar"][t][None, None],
                              g["df_dtactile"][t][None, None], rows, rows, rows, carry=carry, want_q0=False)
        carry = bw["carry"]
        assert rel_err(bw["df_du"][:, 0], g["df_du"][t]) <= 1e-6, t


def test_emulated_matches_oracle_on_fresh_seed():
    """CUDA-path arithmetic vs the numpy oracle on inputs that are NOT in the golden set."""
    from oracle.redmax_oracle import OracleSim
    from tests.blob_scene import scene_from_blob
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz"))
    sc = scene_from_blob(g["ibuf"], g["dbuf"])
    rng = np.random.default_rng(77)
    T = 16
    q0 = g["q0"].copy()
    q0[4] = 0.013
    q0[1] = 0.0005          # start in contact
    u = np.zeros((T, 6))
    u[:, :3] = np.tanh(rng.normal(size=(T, 3)))
    u[:, 0] = 0.9
    u[:, 3:5] = rng.uniform(-1, 1, 2)
    o = OracleSim(sc)
    o.set_state_init(q0, np.zeros(7))
    o.reset(True)
    out = emu_lib.forward(g["ibuf"], g["dbuf"], q0, np.zeros(7), u[:, None, :], grad=True)
    for t in range(T):
        o.set_u(u[t])
        o.forward(1)
        assert rel_err(out["q"][t, 0], o.get_q()) <= 1e-9
        assert rel_err(out["tactile"][t, 0], o.get_tactile_force_vector()) <= 1e-8
        assert (out["status"][t, 0] & 255) == o.newton_iters[t]
        # tape: H and the adjoint blocks G0 = -M + hD, G1 = -hM (FORCE motors)
        H, M, D = o.tape["H"][t], o.tape["M"][t], o.tape["D"][t]
        tp = out["tape"][t, 0, :3 * 49].reshape(3, 7, 7)
        assert rel_err(tp[0], H) <= 1e-8
        assert rel_err(tp[1], -M + o.h * D) <= 1e-8
        assert rel_err(tp[2], -o.h * M) <= 1e-8
    dq, dv, dt = rng.normal(size=(T, 7)), rng.normal(size=(T, 6)), 1e-3 * rng.normal(size=(T, 390))
    ref = o.backward(dq, dv, dt)
    bw = emu_lib.backward(g["ibuf"], g["dbuf"], out, u[:, None, :], dq[:, None], dv[:, None], dt[:, None])
    assert rel_err(bw["df_du"][:, 0], ref["df_du"]) <= 1e-6
    assert rel_err(bw["df_dq0"][0], ref["df_dq0"]) <= 1e-6
    assert rel_err(bw["df_dqdot0"][0], ref["df_dqdot0"]) <= 1e-6


@pytest.mark.parametrize("name", ["dclaw_episodic_s0", "dclaw8x6_episodic_s0"])
def test_emulated_dclaw_matches_reference(name):
    """DClaw rotate-cap (BASELINE configs[3], the reference's own asset): 10 reduced dofs (16-dof kernel variant),
    abstract bodies with sampled contact points, cylinder SDF (cap), three abstract 302-marker sensors; and the
    synthetic 3 x (8x6) marker variant BASELINE.json names (oracle/build_ref.sh)."""
    from tests.blob_scene import scene_from_blob
    from tests.multi_force import expected_words
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = scene_from_blob(g["ibuf"], g["dbuf"])
    T = g["u"].shape[0]
    out = emu_lib.forward(g["ibuf"], g["dbuf"], g["q0"], g["qd0"], g["u"][:, None, :], grad=True)
    assert int((out["status"] >> 16).max()) == 0
    assert float(np.abs(g["tactile"]).max()) > 0
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["q"][t]) <= 1e-9, t
        assert rel_err(out["qd"][t, 0], g["qd"][t]) <= 1e-9, t
        assert rel_err(out["var"][t, 0], g["var"][t]) <= 1e-9, t
        assert rel_err(out["tactile"][t, 0], g["tactile"][t]) <= 1e-8, t
        assert np.array_equal(out["cmask"][t, 0].astype(np.uint64), expected_words(sc, g["ground_ids_f"][t], g["gp_ids_f"][t])), t
        assert np.array_equal(out["marker_body"][t, 0], g["marker_body"][t]), t
    bw = emu_lib.backward(g["ibuf"], g["dbuf"], out, g["u"][:, None, :], g["df_dq"][:, None, :],
                          g["df_dvar"][:, None, :], g["df_dtactile"][:, None, :])
    assert rel_err(bw["df_du"][:, 0], g["df_du"]) <= 1e-6
    assert rel_err(bw["df_dq0"][0], g["df_dq0"]) <= 1e-6
    assert rel_err(bw["df_dqdot0"][0], g["df_dqdot0"]) <= 1e-6


@pytest.mark.parametrize("name", ["insertion_episodic_s0", "stable_grasp_episodic_s0", "insertion20x20_episodic_s0"])
def test_emulated_insertion_and_stable_grasp_match_reference(name):
    """TactileInsertion (BASELINE configs[4], the reference's own asset): 12 reduced dofs, position-controlled base
    (PD on the previous state: extra adjoint terms), free3d-euler box, prismatic fingers, ten general-primitive
    contacts, two sensors with 7 candidate bodies each (one of them the other pad, a cylinder).
    StableGrasp: all four motors position-controlled, a bar of eleven boxes on a free3d-euler joint, 44
    general-primitive + 11 ground contacts, 15 candidate bodies per sensor."""
    from tests.blob_scene import scene_from_blob
    from tests.multi_force import expected_words
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = scene_from_blob(g["ibuf"], g["dbuf"])
    T = g["u"].shape[0]
    out = emu_lib.forward(g["ibuf"], g["dbuf"], g["q0"], g["qd0"], g["u"][:, None, :], grad=True)
    assert int((out["status"] >> 16).max()) == 0
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["q"][t]) <= 1e-9, t
        assert rel_err(out["qd"][t, 0], g["qd"][t]) <= 1e-9, t
        assert rel_err(out["tactile"][t, 0], g["tactile"][t]) <= 1e-8, t
        assert np.array_equal(out["cmask"][t, 0].astype(np.uint64), expected_words(sc, g["ground_ids_f"][t], g["gp_ids_f"][t])), t
        assert np.array_equal(out["marker_body"][t, 0], g["marker_body"][t]), t
    bw = emu_lib.backward(g["ibuf"], g["dbuf"], out, g["u"][:, None, :], g["df_dq"][:, None, :], None,
                          g["df_dtactile"][:, None, :])
    assert rel_err(bw["df_du"][:, 0], g["df_du"]) <= 1e-6
    assert rel_err(bw["df_dq0"][0], g["df_dq0"]) <= 1e-6
    assert rel_err(bw["df_dqdot0"][0], g["df_dqdot0"]) <= 1e-6


def test_emulated_rolling_ball_matches_reference():
    """examples/RollingBallExp (BASELINE configs[0], the reference's own scene and action schedule): BDF2 with its
    SDIRK2 start-up step, free3d-exp ball, sphere SDF (ground, pad contact, tactile), 2168 sampled pad points;
    kernel variant 17.  Then the same trajectory in chunks through the multistep state (tsim_forward_multistep),
    and four frames of the real 200x200 tactile field."""
    from tests import rolling_ball as rb
    g = np.load(os.path.join(GOLDEN, "rollingball_bdf2_s0.npz"))
    T = g["u"].shape[0]
    rows = rb.tactile_rows(T, int(g["tactile_every"]))
    u = g["u"][:, None, :]
    out = emu_lib.forward(g["ibuf"], g["dbuf"], g["q0"], g["qd0"], u, tac_row=rows)
    rb.check_trajectory(out["q"][:, 0], out["qd"][:, 0], out["status"][:, 0], out["cmask"][:, 0], out["tactile"][:, 0],
                        out["marker_body"][:, 0], g)
    # chunked: 1 + 4 + 95 steps, history carried by the caller; bit-identical to the single call
    n = len(g["q0"])
    q, qd = g["q0"].copy()[None], g["qd0"].copy()[None]
    hist = (np.zeros((1, n)), np.zeros((1, n)))
    done = 0
    for chunk in (1, 4, 95):
        o = emu_lib.forward(g["ibuf"], g["dbuf"], q, qd, u[done:done + chunk], want_masks=False, hist=hist, steps_done=done)
        assert np.array_equal(o["q"][:, 0], out["q"][done:done + chunk, 0])
        q, qd = o["q_final"], o["qd_final"]
        done += chunk
    assert np.array_equal(hist[0][0], out["q"][done - 2, 0])
    # the real 200x200 sensor at four steps
    ib, db = rb.full_resolution_blob(g["ibuf"], g["dbuf"])
    frames = [int(f) for f in g["frames200"]]
    rows = np.full(T, -1, dtype=np.int32)
    rows[frames] = np.arange(len(frames))
    big = emu_lib.forward(ib, db, g["q0"], g["qd0"], u, tac_row=rows, want_masks=False)
    assert np.array_equal(big["q"], out["q"])
    rb.check_full_resolution_frames(big["tactile"][:, 0], g)


@pytest.mark.parametrize("name,code", [("BDF2", 1), ("SDIRK2", 2)])
def test_emulated_other_integrators_match_reference(name, code):
    """options.integrator = BDF2 / SDIRK2 (DH/Simulation.cpp:1076-1092, 1353-1564) on the TactilePush scene, forward only."""
    g = np.load(os.path.join(GOLDEN, "pusher13x10_integrators_s0.npz"))
    ib = g["ibuf"].copy()
    ib[14] = code
    T = g["u"].shape[0]
    out = emu_lib.forward(ib, g["dbuf"], g["q0"], g["qd0"], g["u"][:, None, :])
    assert int((out["status"] >> 16).max()) == 0
    assert int((g["gp_ids_" + name] >= 0).sum()) > 0
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["q_" + name][t]) <= 1e-9, t
        assert rel_err(out["qd"][t, 0], g["qd_" + name][t]) <= 1e-9, t
        assert rel_err(out["var"][t, 0], g["var_" + name][t]) <= 1e-9, t
        assert rel_err(out["tactile"][t, 0], g["tactile_" + name][t]) <= 1e-8, t
        assert emu_lib.mask_to_ids(out["cmask"][t, 0, 1:4]) == [int(x) for x in g["gp_ids_" + name][t] if x >= 0], t


@pytest.mark.parametrize("name", ["spherical_euler_bdf1_s0", "free2d_plate_bdf1_s0"])
def test_emulated_spherical_euler_chain_matches_reference(name):
    """spherical-euler joints (DH/Joint/JointSphericalEuler.cpp) in a two-link arm, and a free2d joint
    (DH/Joint/JointFree2D.cpp) in a tilted plane carrying a revolute arm -- scenes of our own, no reference asset uses
    these joint types: a pad pressed by the neighbouring link, BDF1 forward + Simulation::backward."""
    from tests.blob_scene import scene_from_blob
    from tests.multi_force import expected_words
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = scene_from_blob(g["ibuf"], g["dbuf"])
    T = g["u"].shape[0]
    out = emu_lib.forward(g["ibuf"], g["dbuf"], g["q0"], g["qd0"], g["u"][:, None, :], grad=True)
    assert int((out["status"] >> 16).max()) == 0
    assert float(np.abs(g["tactile"]).max()) > 0
    assert int((g["ground_ids_f"] >= 0).sum()) > 0 or name.startswith("free2d")
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["q"][t]) <= 1e-9, t
        assert rel_err(out["qd"][t, 0], g["qd"][t]) <= 1e-9, t
        assert rel_err(out["var"][t, 0], g["var"][t]) <= 1e-9, t
        assert rel_err(out["tactile"][t, 0], g["tactile"][t]) <= 1e-8, t
        assert np.array_equal(out["cmask"][t, 0].astype(np.uint64), expected_words(sc, g["ground_ids_f"][t], g["gp_ids_f"][t])), t
        assert np.array_equal(out["marker_body"][t, 0], g["marker_body"][t]), t
    bw = emu_lib.backward(g["ibuf"], g["dbuf"], out, g["u"][:, None, :], g["df_dq"][:, None, :],
                          g["df_dvar"][:, None, :], g["df_dtactile"][:, None, :])
    assert rel_err(bw["df_du"][:, 0], g["df_du"]) <= 1e-6
    assert rel_err(bw["df_dq0"][0], g["df_dq0"]) <= 1e-6
    assert rel_err(bw["df_dqdot0"][0], g["df_dqdot0"]) <= 1e-6


def test_emulated_spherical_exp_chain_matches_reference():
    """spherical-exp elbow (DH/Joint/JointSphericalExp.cpp) under BDF2, forward only (kernel variant 17)."""
    g = np.load(os.path.join(GOLDEN, "spherical_exp_bdf2_s0.npz"))
    T = g["u"].shape[0]
    out = emu_lib.forward(g["ibuf"], g["dbuf"], g["q0"], g["qd0"], g["u"][:, None, :])
    assert int((out["status"] >> 16).max()) == 0
    assert int((g["ground_ids"] >= 0).sum()) > 0
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["q"][t]) <= 1e-9, t
        assert rel_err(out["qd"][t, 0], g["qd"][t]) <= 1e-9, t
        assert rel_err(out["var"][t, 0], g["var"][t]) <= 1e-9, t
        assert rel_err(out["tactile"][t, 0], g["tactile"][t]) <= 1e-8, t
        assert emu_lib.mask_to_ids(out["cmask"][t, 0]) == [int(x) for x in g["ground_ids"][t] if x >= 0], t


def test_emulated_rolling_ball_adjoint_matches_reference():
    """The rolling-ball scene under BDF1 with Simulation::backward(): adjoint through the free3d-exp joint, the sphere SDF
    (ground point, pad contact, tactile field with its hand-written reverse mode) and the 2168-point pad; kernel variant 17."""
    g = np.load(os.path.join(GOLDEN, "rollingball_bdf1_adjoint_s0.npz"))
    T, n = g["u"].shape[0], len(g["q0"])
    out = emu_lib.forward(g["ibuf"], g["dbuf"], g["q0"], g["qd0"], g["u"][:, None, :], grad=True)
    assert int((out["status"] >> 16).max()) == 0
    assert float(np.abs(g["tactile"]).max()) > 0 and int((g["gp_ids"] >= 0).sum()) > 0
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["q"][t]) <= 1e-9, t
        assert rel_err(out["qd"][t, 0], g["qd"][t]) <= 1e-9, t
        assert rel_err(out["tactile"][t, 0], g["tactile"][t]) <= 1e-8, t
        assert emu_lib.mask_to_ids(out["cmask"][t, 0, 1:]) == [int(x) for x in g["gp_ids"][t] if x >= 0], t
        assert np.array_equal(out["marker_body"][t, 0], g["marker_body"][t]), t
    rng = np.random.default_rng(int(g["cot_seed"]))
    df_dq = rng.normal(size=(T, n))
    df_dtac = 1e-3 * rng.normal(size=(T, g["tactile"].shape[1]))
    bw = emu_lib.backward(g["ibuf"], g["dbuf"], out, g["u"][:, None, :], df_dq[:, None, :], None, df_dtac[:, None, :])
    assert rel_err(bw["df_du"][:, 0], g["df_du"]) <= 1e-6
    assert rel_err(bw["df_dq0"][0], g["df_dq0"]) <= 1e-6
    assert rel_err(bw["df_dqdot0"][0], g["df_dqdot0"]) <= 1e-6


def test_emulated_capsule_scene_matches_reference():
    """capsule SDF (DH/Body/BodyCapsule.cpp) in a scene of our own: a pad pressed on a capsule that lies on the ground and
    rolls -- contact primitive, tactile candidate (reverse mode between and beyond the caps), ground contact through
    its sampled points; BDF1 forward + Simulation::backward; kernel variant 17."""
    g = np.load(os.path.join(GOLDEN, "capsule_press_bdf1_s0.npz"))
    T, n = g["u"].shape[0], len(g["q0"])
    out = emu_lib.forward(g["ibuf"], g["dbuf"], g["q0"], g["qd0"], g["u"][:, None, :], grad=True)
    assert int((out["status"] >> 16).max()) == 0
    assert float(np.abs(g["tactile"]).max()) > 0 and int((g["gp_ids"] >= 0).sum()) > 0 and int((g["ground_ids"] >= 0).sum()) > 0
    gw = (g["ground_ids"].shape[1] + 31) // 32
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["q"][t]) <= 1e-9, t
        assert rel_err(out["qd"][t, 0], g["qd"][t]) <= 1e-9, t
        assert rel_err(out["tactile"][t, 0], g["tactile"][t]) <= 1e-8, t
        assert emu_lib.mask_to_ids(out["cmask"][t, 0, :gw]) == [int(x) for x in g["ground_ids"][t] if x >= 0], t
        assert emu_lib.mask_to_ids(out["cmask"][t, 0, gw:]) == [int(x) for x in g["gp_ids"][t] if x >= 0], t
        assert np.array_equal(out["marker_body"][t, 0], g["marker_body"][t]), t
    rng = np.random.default_rng(int(g["cot_seed"]))
    df_dq = rng.normal(size=(T, n))
    df_dtac = 1e-3 * rng.normal(size=(T, g["tactile"].shape[1]))
    bw = emu_lib.backward(g["ibuf"], g["dbuf"], out, g["u"][:, None, :], df_dq[:, None, :], None, df_dtac[:, None, :])
    assert rel_err(bw["df_du"][:, 0], g["df_du"]) <= 1e-6
    assert rel_err(bw["df_dq0"][0], g["df_dq0"]) <= 1e-6
    assert rel_err(bw["df_dqdot0"][0], g["df_dqdot0"]) <= 1e-6


def test_emulated_passes_equal_in_loop_evaluation():
    """The trajectory-only passes (tactile read-out, G0 / G1 tape blocks: calls of 4 steps or more) against the evaluation
    inside the step loop (calls of fewer steps, chained): same trajectory, fields, tape and gradients."""
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz"))
    T = 24
    u = g["u"][:T, None, :]
    one = emu_lib.forward(g["ibuf"], g["dbuf"], g["q0"], g["qd0"], u, grad=True)
    q, qd = g["q0"][None].copy(), g["qd0"][None].copy()
    parts = []
    for t0 in range(0, T, 3):
        o = emu_lib.forward(g["ibuf"], g["dbuf"], q, qd, u[t0:t0 + 3], grad=True)
        q, qd = o["q_final"], o["qd_final"]
        parts.append(o)
    for key, tol in (("q", 0.0), ("tactile", 1e-13), ("tape", 1e-13), ("var", 0.0)):
        cat = np.concatenate([p[key] for p in parts], axis=0)
        if tol == 0.0:
            assert np.array_equal(cat, one[key]), key
        else:
            assert rel_err(cat, one[key]) <= tol, key
    assert float(np.abs(one["tactile"]).max()) > 0
