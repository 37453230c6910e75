"""GPU parity tests proper: CUDA path (through the C ABI) vs golden vectors of the reference
on the same seeded inputs.

Tolerances (print_error norm DH/Utils.h:315-319): q, qdot, var <= 1e-9; tactile <= 1e-8 (north_star
allows 1e-4); contact-point index sets and marker->body ids identical; gradients <= 1e-6."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

CASES = ["pusher13x10_episodic_s0", "pusher13x10_episodic_s1", "pusher32x13_episodic_s0"]


def _ids(words):
    out = []
    for w, word in enumerate(words):
        for b in range(32):
            if ((int(word) & 0xffffffff) >> b) & 1:
                out.append(32 * w + b)
    return out


def _sim(g, lanes=8, variant=8):
    """variant 16 = the 16-dof build of the kernels (csrc/kernel_layout.h), forced through the
    TSIM_B200_VARIANT development knob of csrc/cabi.cpp."""
    from tactilesimulation_b200.sim import BatchedSim
    if variant == 16:
        os.environ["TSIM_B200_VARIANT"] = "16"
    try:
        return BatchedSim((g["ibuf"], g["dbuf"]), device="cuda:0", lanes=lanes)
    finally:
        os.environ.pop("TSIM_B200_VARIANT", None)


@pytest.mark.parametrize("lanes,variant", [(8, 8), (16, 8), (32, 8), (16, 16), (32, 16)])
@pytest.mark.parametrize("name", CASES)
def test_forward_and_adjoint_match_reference(name, lanes, variant):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sim = _sim(g, lanes, variant)
    dev = sim.device
    T = g["u"].shape[0]
    B = 3   # three identical envs: also checks that tiles do not interfere
    q = torch.tensor(np.tile(g["q0"], (B, 1)), device=dev)
    qd = torch.tensor(np.tile(g["qd0"], (B, 1)), device=dev)
    u = torch.tensor(np.tile(g["u"][:, None, :], (1, B, 1)), device=dev).contiguous()
    out = sim.forward(q, qd, u, T, grad=True, want_status=True, want_contacts=True)
    torch.cuda.synchronize()
    qt, qdt = out["q_traj"].cpu().numpy(), out["qd_traj"].cpu().numpy()
    var, tac = out["var"].cpu().numpy(), out["tactile"].cpu().numpy()
    cm, mb = out["contact_masks"].cpu().numpy(), out["marker_body"].cpu().numpy()
    assert int((out["status"] >> 16).max().item()) == 0
    st = out["status"].cpu().numpy()
    for e in range(B):
        # Newton iterations / line-search evaluations per step equal the reference's own counters (status bits 0-15)
        assert np.array_equal(st[:, e] & 0xff, g["newton"][:, 0]), e
        assert np.array_equal((st[:, e] >> 8) & 0xff, g["newton"][:, 1]), e
        for t in range(T):
            assert rel_err(qt[t, e], g["q"][t]) <= 1e-9, (t, e)
            assert rel_err(qdt[t, e], g["qd"][t]) <= 1e-9, (t, e)
            assert rel_err(var[t, e], g["var"][t]) <= 1e-9, (t, e)
            assert rel_err(tac[t, e], g["tactile"][t]) <= 1e-8, (t, e)
            assert _ids(cm[t, e, 0:1]) == [int(x) for x in g["ground_ids"][t] if x >= 0], (t, e)
            assert _ids(cm[t, e, 1:4]) == [int(x) for x in g["gp_ids"][t] if x >= 0], (t, e)
            assert np.array_equal(mb[t, e], g["marker_body"][t]), (t, e)
    assert np.allclose(q.cpu().numpy(), qt[-1])
    dq = torch.tensor(np.tile(g["df_dq"][:, None, :], (1, B, 1)), device=dev).contiguous()
    dv = torch.tensor(np.tile(g["df_dvar"][:, None, :], (1, B, 1)), device=dev).contiguous()
    dt = torch.tensor(np.tile(g["df_dtactile"][:, None, :], (1, B, 1)), device=dev).contiguous()
    bw = sim.backward(out, u, T, dq, dv, dt, want_q0=True)
    torch.cuda.synchronize()
    for e in range(B):
        assert rel_err(bw["df_du"][:, e].cpu().numpy(), g["df_du"]) <= 1e-6
        assert rel_err(bw["df_dq0"][e].cpu().numpy(), g["df_dq0"]) <= 1e-6
        assert rel_err(bw["df_dqdot0"][e].cpu().numpy(), g["df_dqdot0"]) <= 1e-6


def test_stepsim_chain_matches_reference():
    """forward(5, save_last_frame_var_only) per gym step, then chained backward over 5-step chunks
    with the carry -- the StepSimFunction pattern (R/envs/redmax_torch_functions.py:112-174)."""
    g = np.load(os.path.join(GOLDEN, "pusher13x10_stepsim_s0.npz"))
    sim = _sim(g)
    dev = sim.device
    fs, ns = int(g["frame_skip"]), g["u"].shape[0]
    B = 2
    q = torch.tensor(np.tile(g["q0"], (B, 1)), device=dev)
    qd = torch.tensor(np.tile(g["qd0"], (B, 1)), device=dev)
    rows = [-1] * (fs - 1) + [0]
    fwds, us = [], []
    for t in range(ns):
        u = torch.tensor(np.tile(g["u"][t], (B, 1)), device=dev)
        o = sim.forward(q, qd, u, fs, grad=True, var_rows=rows, tac_rows=rows)
        fwds.append(o)
        us.append(u)
        assert rel_err(q[0].cpu().numpy(), g["q"][t]) <= 1e-9
        assert rel_err(o["var"][0, 1].cpu().numpy(), g["var"][t]) <= 1e-9
        assert rel_err(o["tactile"][0, 0].cpu().numpy(), g["tactile"][t]) <= 1e-8
    carry = None
    for t in range(ns - 1, -1, -1):
        dq = torch.tensor(np.tile(g["df_dq"][t], (1, B, 1)), device=dev).contiguous()
        dv = torch.tensor(np.tile(g["df_dvar"][t], (1, B, 1)), device=dev).contiguous()
        dt = torch.tensor(np.tile(g["df_dtactile"][t], (1, B, 1)), device=dev).contiguous()
        bw = sim.backward(fwds[t], us[t], fs, dq, dv, dt, dq_rows=rows, dvar_rows=rows, dtac_rows=rows, carry=carry)
        carry = bw["carry"]
        assert rel_err(bw["df_du"][:, 0].cpu().numpy(), g["df_du"][t]) <= 1e-6, t


def test_readout_matches_step_outputs():
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz"))
    sim = _sim(g)
    dev = sim.device
    t = 40
    q = torch.tensor(g["q"][t][None], device=dev)
    qd = torch.tensor(g["qd"][t][None], device=dev)
    r = sim.readout(q, qd, want_contacts=True)
    assert rel_err(r["tactile"][0].cpu().numpy(), g["tactile"][t]) <= 1e-8
    assert rel_err(r["var"][0].cpu().numpy(), g["var"][t]) <= 1e-9
    assert np.array_equal(r["marker_body"][0].cpu().numpy(), g["marker_body"][t])


def test_batched_line_search_equals_sequential_search():
    """TSIM_OPT_LS_BATCH evaluates the step lengths of a struggling line search in parallel; the accepted
    step, the iterate sequence and the evaluation counts must be those of the sequential search
    (DH/Simulation.cpp:1186-1200).  The bench inputs hold an environment whose Newton runs into the
    iteration cap (140 iterations of up to 20 trials): both searches must agree bit for bit there too."""
    from bench import make_inputs
    g = np.load(os.path.join(GOLDEN, "pusher32x13_episodic_s0.npz"))
    B, T = 4096, 200
    q0, qd0, u, _ = make_inputs(g["q0"], B, T, 1234)
    sel = np.r_[2800:2828, 1940:1968, 0:8]           # the block with the non-converging env, a hard one, easy ones
    outs = []
    for batch in (1, 0):
        sim = _sim(g)
        sim.set_option(0, batch)
        dev = sim.device
        q, qd = torch.tensor(q0[sel], device=dev), torch.tensor(qd0[sel], device=dev)
        ut = torch.tensor(np.ascontiguousarray(u[:, sel]), device=dev)
        o = sim.forward(q, qd, ut, T, grad=True, want_status=True, want_tactile=False)
        torch.cuda.synchronize()
        outs.append((o["q_traj"].cpu().numpy(), o["status"].cpu().numpy(), o["tape"].cpu().numpy()))
    (qa, sa, ta), (qb, sb, tb) = outs
    assert int(((sa >> 8) & 255).max()) >= 3 and int((sa & 255).max()) >= 100, "inputs no longer exercise a struggling search"
    assert np.array_equal(sa, sb)
    assert np.array_equal(qa, qb)
    assert np.array_equal(ta, tb)


MULTI = {"dclaw_episodic_s0": (10, 90, 9, 12, 2718), "insertion_episodic_s0": (12, 78, 6, 0, 780),
         "dclaw8x6_episodic_s0": (10, 90, 9, 12, 432),         # synthetic 3 x (8x6) pads of BASELINE configs[3]
         "insertion20x20_episodic_s0": (12, 78, 6, 0, 2400),   # synthetic 2 x (20x20) pads of BASELINE configs[4]
         "stable_grasp_episodic_s0": (12, 126, 6, 0, 780),
         "spherical_euler_bdf1_s0": (6, 12, 6, 3, 48),        # our own two-link arm on spherical-euler joints
         "free2d_plate_bdf1_s0": (4, 12, 4, 3, 36)}           # our own plate on a free2d joint carrying a revolute arm


@pytest.mark.parametrize("lanes", [16, 32])
@pytest.mark.parametrize("name", sorted(MULTI))
def test_dclaw_and_insertion_match_reference(name, lanes):
    """The reference's own assets of BASELINE configs[3] and [4] on the 16-dof kernel variant.
    DClaw rotate-cap: 10 reduced dofs, abstract bodies, cylinder SDF, three abstract 302-marker sensors, three
    contact forces.  TactileInsertion: 12 reduced dofs, position-controlled base (extra adjoint terms),
    free3d-euler box, prismatic fingers, ground + ten general-primitive contacts, two 13x10 pads with seven
    candidate bodies each.  StableGrasp (the fourth env of the reference): four position-controlled motors, a bar
    of eleven boxes, 44 + 11 contact forces, fifteen candidate bodies per pad."""
    from tests.blob_scene import scene_from_blob
    from tests.multi_force import expected_words
    from tactilesimulation_b200.sim import BatchedSim
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sc = scene_from_blob(g["ibuf"], g["dbuf"])
    sim = BatchedSim((g["ibuf"], g["dbuf"]), device="cuda:0", lanes=lanes)
    assert (sim.ndof_r, sim.ndof_m, sim.ndof_u, sim.ndof_var, sim.ndof_tactile) == MULTI[name]
    dev = sim.device
    T, B = g["u"].shape[0], 3
    q = torch.tensor(np.tile(g["q0"], (B, 1)), device=dev)
    qd = torch.tensor(np.tile(g["qd0"], (B, 1)), device=dev)
    u = torch.tensor(np.tile(g["u"][:, None, :], (1, B, 1)), device=dev).contiguous()
    out = sim.forward(q, qd, u, T, grad=True, want_status=True, want_contacts=True)
    torch.cuda.synchronize()
    qt, qdt = out["q_traj"].cpu().numpy(), out["qd_traj"].cpu().numpy()
    tac = out["tactile"].cpu().numpy()
    cm, mb = out["contact_masks"].cpu().numpy(), out["marker_body"].cpu().numpy()
    assert int((out["status"] >> 16).max().item()) == 0
    assert float(np.abs(g["tactile"]).max()) > 0
    st = out["status"].cpu().numpy()
    for e in range(B):
        if "newton" in g.files:       # Newton iterations / line-search evaluations per step as counted in the reference
            assert np.array_equal(st[:, e] & 0xff, g["newton"][:, 0]), e
            assert np.array_equal((st[:, e] >> 8) & 0xff, g["newton"][:, 1]), e
        for t in range(T):
            assert rel_err(qt[t, e], g["q"][t]) <= 1e-9, (t, e)
            assert rel_err(qdt[t, e], g["qd"][t]) <= 1e-9, (t, e)
            if sim.ndof_var:
                assert rel_err(out["var"][t, e].cpu().numpy(), g["var"][t]) <= 1e-9, (t, e)
            assert rel_err(tac[t, e], g["tactile"][t]) <= 1e-8, (t, e)
            assert np.array_equal(cm[t, e].astype(np.uint32).astype(np.uint64), expected_words(sc, g["ground_ids_f"][t], g["gp_ids_f"][t])), (t, e)
            assert np.array_equal(mb[t, e], g["marker_body"][t]), (t, e)
    dq = torch.tensor(np.tile(g["df_dq"][:, None, :], (1, B, 1)), device=dev).contiguous()
    dv = torch.tensor(np.tile(g["df_dvar"][:, None, :], (1, B, 1)), device=dev).contiguous() if sim.ndof_var else None
    dt = torch.tensor(np.tile(g["df_dtactile"][:, None, :], (1, B, 1)), device=dev).contiguous()
    bw = sim.backward(out, u, T, dq, dv, dt, want_q0=True)
    torch.cuda.synchronize()
    for e in range(B):
        assert rel_err(bw["df_du"][:, e].cpu().numpy(), g["df_du"]) <= 1e-6
        assert rel_err(bw["df_dq0"][e].cpu().numpy(), g["df_dq0"]) <= 1e-6
        assert rel_err(bw["df_dqdot0"][e].cpu().numpy(), g["df_dqdot0"]) <= 1e-6


def test_newton_cap_option_only_touches_steps_that_hit_it():
    """TSIM_OPT_MAX_NEWTON lowers the reference's iteration cap (DH/Simulation.cpp:1155).  Environments whose steps
    all converge below the cap are bit-identical; a step that hits it is flagged not-converged, as in the reference."""
    from bench import make_inputs
    g = np.load(os.path.join(GOLDEN, "pusher32x13_episodic_s0.npz"))
    q0, qd0, u, _ = make_inputs(g["q0"], 4096, 200, 1234)
    sel = np.r_[2800:2828, 0:28]                     # env 2826 runs into the reference cap of 140 iterations
    res = []
    for cap in (0, 25):
        sim = _sim(g)
        sim.set_option(1, cap)
        dev = sim.device
        o = sim.forward(torch.tensor(q0[sel], device=dev), torch.tensor(qd0[sel], device=dev),
                        torch.tensor(np.ascontiguousarray(u[:, sel]), device=dev), 200, want_status=True, want_tactile=False)
        torch.cuda.synchronize()
        res.append((o["q_traj"].cpu().numpy(), o["status"].cpu().numpy()))
    (qa, sa), (qb, sb) = res
    assert int((sa & 255).max()) == 140 and int((sb & 255).max()) == 25
    easy = (sa & 255).max(axis=0) < 25               # environments that never reach the lower cap
    assert easy.sum() >= 40 and not easy.all()
    assert np.array_equal(qa[:, easy], qb[:, easy]) and np.array_equal(sa[:, easy], sb[:, easy])
    hit = (sb & 255) == 25
    assert ((sb[hit] >> 16) & 1).all()


@pytest.mark.parametrize("lanes", [16, 32])
def test_rolling_ball_matches_reference(lanes):
    """examples/RollingBallExp (BASELINE configs[0]): BDF2 with the SDIRK2 start-up step, free3d-exp ball, sphere SDF,
    2168 sampled pad points, kernel variant 17; tolerances in tests/rolling_ball.py.  Three identical environments,
    then the same trajectory in chunks through the compat Simulation (forward(1) per step, as test_sim_speed.py does)."""
    from tests import rolling_ball as rb
    from tactilesimulation_b200.sim import BatchedSim
    g = np.load(os.path.join(GOLDEN, "rollingball_bdf2_s0.npz"))
    sim = BatchedSim((g["ibuf"], g["dbuf"]), device="cuda:0", lanes=lanes)
    assert sim.integrator == 1
    dev = sim.device
    T, B = g["u"].shape[0], 3
    rows = rb.tactile_rows(T, int(g["tactile_every"]))
    q = torch.tensor(np.tile(g["q0"], (B, 1)), device=dev)
    qd = torch.tensor(np.tile(g["qd0"], (B, 1)), device=dev)
    u = torch.tensor(np.tile(g["u"][:, None, :], (1, B, 1)), device=dev).contiguous()
    out = sim.forward(q, qd, u, T, tac_rows=rows, want_status=True, want_contacts=True)
    torch.cuda.synchronize()
    qt, qdt, st = out["q_traj"].cpu().numpy(), out["qd_traj"].cpu().numpy(), out["status"].cpu().numpy()
    cm, tac, mb = out["contact_masks"].cpu().numpy(), out["tactile"].cpu().numpy(), out["marker_body"].cpu().numpy()
    for e in range(B):
        rb.check_trajectory(qt[:, e], qdt[:, e], st[:, e], cm[:, e], tac[:, e], mb[:, e], g)
    with pytest.raises(Exception):
        sim.forward(q, qd, u, 2, grad=True)          # no adjoint for BDF2, as in the reference
    # chunked through the multistep state: bit-identical to the single call
    q2 = torch.tensor(np.tile(g["q0"], (B, 1)), device=dev)
    qd2 = torch.zeros_like(q2)
    qp, qdp = torch.zeros_like(q2), torch.zeros_like(q2)
    done = 0
    for chunk in (1, 4, 45):
        o = sim.forward(q2, qd2, u[done:done + chunk].contiguous(), chunk, want_tactile=False, q_prev=qp, qd_prev=qdp,
                        steps_done=done)
        assert torch.equal(o["q_traj"], out["q_traj"][done:done + chunk])
        done += chunk


def test_rolling_ball_full_resolution_through_compat_simulation():
    """The real 200x200 sensor (120 000 tactile values per frame) driven like examples/RollingBallExp/test_sim_speed.py:
    set_u + forward(1) per step on the drop-in Simulation, tactile read at four steps."""
    from tests import rolling_ball as rb
    from tactilesimulation_b200.layout import scene_from_blob
    from tactilesimulation_b200.redmax import Simulation
    g = np.load(os.path.join(GOLDEN, "rollingball_bdf2_s0.npz"))
    ib, db = rb.full_resolution_blob(g["ibuf"], g["dbuf"])
    sim = Simulation(scene_from_blob(ib, db))
    assert (sim.ndof_r, sim.ndof_u, sim.ndof_tactile, sim.options.integrator) == (9, 3, 120000, "BDF2")
    with pytest.raises(Exception):
        sim.reset(backward_flag=True)
    sim.reset(backward_flag=False)
    frames = [int(f) for f in g["frames200"]]
    got = []
    for t in range(frames[-1] + 1):
        sim.set_u(g["u"][t])
        sim.forward(1, verbose=False, test_derivatives=False)
        if t in frames:
            got.append(sim.get_tactile_force_vector().copy())
            assert rel_err(sim.get_q(), g["q"][t]) <= rb.TOL_STATE
    rb.check_full_resolution_frames(np.array(got), g)


@pytest.mark.parametrize("name,code", [("BDF2", 1), ("SDIRK2", 2)])
def test_other_integrators_match_reference(name, code):
    """options.integrator = BDF2 / SDIRK2 on the TactilePush scene (kernel variant 17), forward only."""
    from tactilesimulation_b200.sim import BatchedSim
    g = np.load(os.path.join(GOLDEN, "pusher13x10_integrators_s0.npz"))
    ib = g["ibuf"].copy()
    ib[14] = code
    sim = BatchedSim((ib, g["dbuf"]), device="cuda:0")
    dev = sim.device
    T, B = g["u"].shape[0], 2
    q = torch.tensor(np.tile(g["q0"], (B, 1)), device=dev)
    qd = torch.zeros_like(q)
    u = torch.tensor(np.tile(g["u"][:, None, :], (1, B, 1)), device=dev).contiguous()
    out = sim.forward(q, qd, u, T, want_status=True, want_contacts=True)
    qt, qdt, tac = out["q_traj"].cpu().numpy(), out["qd_traj"].cpu().numpy(), out["tactile"].cpu().numpy()
    cm = out["contact_masks"].cpu().numpy()
    assert int((out["status"] >> 16).max().item()) == 0
    for e in range(B):
        for t in range(T):
            assert rel_err(qt[t, e], g["q_" + name][t]) <= 1e-9, (t, e)
            assert rel_err(qdt[t, e], g["qd_" + name][t]) <= 1e-9, (t, e)
            assert rel_err(tac[t, e], g["tactile_" + name][t]) <= 1e-8, (t, e)
            assert _ids(cm[t, e, 1:4]) == [int(x) for x in g["gp_ids_" + name][t] if x >= 0], (t, e)


def test_spherical_exp_chain_matches_reference():
    """spherical-euler shoulder + spherical-exp elbow under BDF2 (kernel variant 17), forward only."""
    from tactilesimulation_b200.sim import BatchedSim
    g = np.load(os.path.join(GOLDEN, "spherical_exp_bdf2_s0.npz"))
    sim = BatchedSim((g["ibuf"], g["dbuf"]), device="cuda:0")
    dev = sim.device
    T, B = g["u"].shape[0], 2
    q = torch.tensor(np.tile(g["q0"], (B, 1)), device=dev)
    qd = torch.zeros_like(q)
    u = torch.tensor(np.tile(g["u"][:, None, :], (1, B, 1)), device=dev).contiguous()
    out = sim.forward(q, qd, u, T, want_status=True, want_contacts=True)
    qt, qdt, tac = out["q_traj"].cpu().numpy(), out["qd_traj"].cpu().numpy(), out["tactile"].cpu().numpy()
    var, cm = out["var"].cpu().numpy(), out["contact_masks"].cpu().numpy()
    assert int((out["status"] >> 16).max().item()) == 0
    for e in range(B):
        for t in range(T):
            assert rel_err(qt[t, e], g["q"][t]) <= 1e-9, (t, e)
            assert rel_err(qdt[t, e], g["qd"][t]) <= 1e-9, (t, e)
            assert rel_err(var[t, e], g["var"][t]) <= 1e-9, (t, e)
            assert rel_err(tac[t, e], g["tactile"][t]) <= 1e-8, (t, e)
            assert _ids(cm[t, e]) == [int(x) for x in g["ground_ids"][t] if x >= 0], (t, e)


def test_trajectory_only_passes_equal_the_in_loop_evaluation():
    """TSIM_OPT_TAC_PASS / TAPE_PASS / VJP_PASS (tactile read-out, G0 / G1 tape blocks, cotangent pull-back in passes of
    their own over the recorded trajectory, the default) against the evaluation inside the step loop / reverse sweep:
    the same code on the same inputs.  The trajectory is bit-identical; fields, tape and gradients agree to rounding
    (different kernels, so fused multiply-adds may be placed differently)."""
    g = np.load(os.path.join(GOLDEN, "pusher32x13_episodic_s0.npz"))
    T, B = g["u"].shape[0], 40
    rng = np.random.default_rng(5)
    q0 = np.tile(g["q0"], (B, 1))
    q0[:, 4] += rng.uniform(-0.01, 0.01, B)
    u = np.tile(g["u"][:, None, :], (1, B, 1))
    u[:, :, :3] += 0.05 * rng.normal(size=(T, B, 3))
    res = []
    for on in (1, 0):
        sim = _sim(g)
        for key in (2, 3, 4):
            sim.set_option(key, on)
        dev = sim.device
        q, qd = torch.tensor(q0, device=dev), torch.zeros((B, 7), dtype=torch.float64, device=dev)
        ut = torch.tensor(u, device=dev).contiguous()
        out = sim.forward(q, qd, ut, T, grad=True, want_status=True)
        dq = torch.ones_like(out["q_traj"])
        dv = torch.ones_like(out["var"])
        dt = torch.full_like(out["tactile"], 1e-3)
        bw = sim.backward(out, ut, T, dq, dv, dt, want_q0=True)
        kt = sim.kernel_times()
        assert (kt["tac_kernel"] is not None) == bool(on) and (kt["tape_kernel"] is not None) == bool(on)
        assert (kt["vjp_kernel"] is not None) == bool(on)
        res.append({k: v.cpu().numpy() for k, v in dict(q=out["q_traj"], tac=out["tactile"], tape=out["tape"], var=out["var"],
                                                         du=bw["df_du"], dq0=bw["df_dq0"], dqd0=bw["df_dqdot0"]).items()})
    a, b = res
    assert float(np.abs(a["tac"]).max()) > 0
    assert np.array_equal(a["q"], b["q"]) and np.array_equal(a["var"], b["var"])
    assert rel_err(a["tac"], b["tac"]) <= 1e-13
    assert rel_err(a["tape"], b["tape"]) <= 1e-13
    for k in ("du", "dq0", "dqd0"):
        assert rel_err(a[k], b[k]) <= 1e-10, k


@pytest.mark.parametrize("name", ["rollingball_bdf1_adjoint_s0", "capsule_press_bdf1_s0"])
def test_rolling_ball_adjoint_matches_reference(name):
    """The rolling-ball scene under BDF1 with Simulation::backward() (kernel variant 17): adjoint through the free3d-exp
    joint, the sphere SDF (ground point, pad contact, tactile field) and the 2168-point pad.  And a scene of our own with
    a capsule (contact primitive, tactile candidate between and beyond its caps, ground contact through sampled points)."""
    from tactilesimulation_b200.sim import BatchedSim
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    gw = (g["ground_ids"].shape[1] + 31) // 32 if "ground_ids" in g.files else 1
    sim = BatchedSim((g["ibuf"], g["dbuf"]), device="cuda:0")
    assert sim.integrator == 0
    dev = sim.device
    T, B, n = g["u"].shape[0], 2, len(g["q0"])
    q = torch.tensor(np.tile(g["q0"], (B, 1)), device=dev)
    qd = torch.zeros_like(q)
    u = torch.tensor(np.tile(g["u"][:, None, :], (1, B, 1)), device=dev).contiguous()
    out = sim.forward(q, qd, u, T, grad=True, want_status=True, want_contacts=True)
    qt, tac = out["q_traj"].cpu().numpy(), out["tactile"].cpu().numpy()
    cm, mb = out["contact_masks"].cpu().numpy(), out["marker_body"].cpu().numpy()
    assert int((out["status"] >> 16).max().item()) == 0
    for e in range(B):
        for t in range(T):
            assert rel_err(qt[t, e], g["q"][t]) <= 1e-9, (t, e)
            assert rel_err(tac[t, e], g["tactile"][t]) <= 1e-8, (t, e)
            assert _ids(cm[t, e, gw:]) == [int(x) for x in g["gp_ids"][t] if x >= 0], (t, e)
            if "ground_ids" in g.files:
                assert _ids(cm[t, e, :gw]) == [int(x) for x in g["ground_ids"][t] if x >= 0], (t, e)
            assert np.array_equal(mb[t, e], g["marker_body"][t]), (t, e)
    rng = np.random.default_rng(int(g["cot_seed"]))
    df_dq = rng.normal(size=(T, n))
    df_dtac = 1e-3 * rng.normal(size=(T, g["tactile"].shape[1]))
    dq = torch.tensor(np.tile(df_dq[:, None, :], (1, B, 1)), device=dev).contiguous()
    dt = torch.tensor(np.tile(df_dtac[:, None, :], (1, B, 1)), device=dev).contiguous()
    bw = sim.backward(out, u, T, dq, None, dt, want_q0=True)
    for e in range(B):
        assert rel_err(bw["df_du"][:, e].cpu().numpy(), g["df_du"]) <= 1e-6
        assert rel_err(bw["df_dq0"][e].cpu().numpy(), g["df_dq0"]) <= 1e-6
        assert rel_err(bw["df_dqdot0"][e].cpu().numpy(), g["df_dqdot0"]) <= 1e-6
