"""The reference-facing API on the GPU: the redmax_py-compatible Simulation (numpy face, batch 1)
driven exactly as R/envs/redmax_torch_functions.py drives the reference, and the batched
autograd Functions.  Checked against the reference's golden vectors."""
import os

import numpy as np
import pytest
import torch

from tests.blob_scene import scene_from_blob
from tests.conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu


def _scene(g):
    return scene_from_blob(g["ibuf"], g["dbuf"])


def test_compat_simulation_stepsim_call_sequence():
    """Verbatim call sequence of the reference StepSimFunction (forward :129-138, backward :141-174)."""
    from tactilesimulation_b200.redmax import Simulation
    g = np.load(os.path.join(GOLDEN, "pusher13x10_stepsim_s0.npz"))
    sim = Simulation(_scene(g))
    assert (sim.ndof_r, sim.ndof_m, sim.ndof_u, sim.ndof_var, sim.ndof_tactile) == (7, 42, 6, 6, 390)
    assert sim.options.h == pytest.approx(0.004999999888241291, abs=0)
    fs, ns = int(g["frame_skip"]), g["u"].shape[0]
    sim.set_state_init(g["q0"], g["qd0"])
    sim.reset(backward_flag=True)
    for t in range(ns):
        sim.set_u(g["u"][t])
        sim.forward(fs, verbose=False, test_derivatives=False, save_last_frame_var_only=True)
        assert rel_err(sim.get_q(), g["q"][t]) <= 1e-9
        assert rel_err(sim.get_variables(), g["var"][t]) <= 1e-9
        assert rel_err(sim.get_tactile_force_vector(), g["tactile"][t]) <= 1e-8
    for t in range(ns - 1, -1, -1):
        sim.backward_info.set_flags(flag_q0=False, flag_qdot0=False, flag_p=False, flag_u=True)
        a = np.zeros(sim.ndof_r * fs); a[-sim.ndof_r:] = g["df_dq"][t]
        b = np.zeros(sim.ndof_var * fs); b[-sim.ndof_var:] = g["df_dvar"][t]
        c = np.zeros(sim.ndof_tactile * fs); c[-sim.ndof_tactile:] = g["df_dtactile"][t]
        sim.backward_info.df_dq, sim.backward_info.df_dvar, sim.backward_info.df_dtactile = a, b, c
        sim.backward_info.df_du = np.zeros(sim.ndof_u * fs)
        sim.backward_steps(fs)
        assert rel_err(sim.backward_results.df_du.reshape(fs, sim.ndof_u), g["df_du"][t]) <= 1e-6


def test_compat_simulation_episodic_call_sequence_and_errors():
    from tactilesimulation_b200 import TactileSimError
    from tactilesimulation_b200.redmax import Simulation
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s1.npz"))
    sim = Simulation(_scene(g))
    with pytest.raises(TactileSimError):
        sim.forward(1)                      # reset() must come first
    with pytest.raises(TactileSimError):
        sim.set_u(np.zeros(3))              # wrong size
    T = g["u"].shape[0]
    sim.set_state_init(g["q0"], g["qd0"])
    sim.reset(backward_flag=True)
    with pytest.raises(TactileSimError):
        sim.backward()                      # forward() must come first
    for t in range(T):
        sim.set_u(g["u"][t])
        sim.forward(1)
        if t % 7 == 0:
            assert rel_err(sim.get_q(), g["q"][t]) <= 1e-9
            assert rel_err(sim.get_tactile_force_vector(), g["tactile"][t]) <= 1e-8
    sim.saveBackwardCache()
    sim.popBackwardCache()
    bi = sim.backward_info
    bi.set_flags(True, True, False, True)
    bi.df_dq, bi.df_dvar, bi.df_dtactile = g["df_dq"].reshape(-1), g["df_dvar"].reshape(-1), g["df_dtactile"].reshape(-1)
    bi.df_dq0, bi.df_dqdot0, bi.df_du = np.zeros(7), np.zeros(7), np.zeros(6 * T)
    sim.backward()
    br = sim.backward_results
    assert rel_err(br.df_du.reshape(T, 6), g["df_du"]) <= 1e-6
    assert rel_err(br.df_dq0, g["df_dq0"]) <= 1e-6
    assert rel_err(br.df_dqdot0, g["df_dqdot0"]) <= 1e-6
    bi.df_dq = np.zeros(3)
    with pytest.raises(TactileSimError):
        sim.backward()                      # size validation, Simulation.cpp:1598-1600


def test_batched_episodic_function_autograd():
    from tactilesimulation_b200.redmax import Simulation
    from tactilesimulation_b200.torch_functions import EpisodicSimFunction
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz"))
    B, T = 4, g["u"].shape[0]
    sim = Simulation(_scene(g), batch=B)
    dev = sim.device
    q0 = torch.tensor(np.tile(g["q0"], (B, 1)), device=dev, requires_grad=True)
    qd0 = torch.zeros((B, 7), dtype=torch.float64, device=dev, requires_grad=True)
    acts = torch.tensor(np.tile(g["u"][:, None, :], (1, B, 1)), device=dev, requires_grad=True)
    masks = torch.ones(T, dtype=torch.bool)
    qs, vs, tacs = EpisodicSimFunction.apply(q0, qd0, acts, masks, sim, True)
    assert rel_err(qs[:, 2].detach().cpu().numpy(), g["q"]) <= 1e-9
    assert rel_err(tacs[:, 1].detach().cpu().numpy(), g["tactile"]) <= 1e-8
    w = [torch.tensor(np.tile(g[k][:, None, :], (1, B, 1)), device=dev) for k in ("df_dq", "df_dvar", "df_dtactile")]
    loss = (qs * w[0]).sum() + (vs * w[1]).sum() + (tacs * w[2]).sum()
    loss.backward()
    for e in range(B):
        assert rel_err(acts.grad[:, e].cpu().numpy(), g["df_du"]) <= 1e-6
        assert rel_err(q0.grad[e].cpu().numpy(), g["df_dq0"]) <= 1e-6
        assert rel_err(qd0.grad[e].cpu().numpy(), g["df_dqdot0"]) <= 1e-6


def test_batched_stepsim_function_autograd_chain():
    """gd.py-style rollout: obs -> action -> StepSimFunction chained through q; gradient w.r.t. the
    per-step actions must equal the reference's backward_steps chain (sum over sub-steps)."""
    from tactilesimulation_b200.redmax import Simulation
    from tactilesimulation_b200.torch_functions import StepSimFunction
    g = np.load(os.path.join(GOLDEN, "pusher13x10_stepsim_s0.npz"))
    B, fs, ns = 2, int(g["frame_skip"]), g["u"].shape[0]
    sim = Simulation(_scene(g), batch=B)
    dev = sim.device
    sim.set_state_init(np.tile(g["q0"], (B, 1)), np.zeros((B, 7)))
    sim.reset(backward_flag=True)
    us = [torch.tensor(np.tile(g["u"][t], (B, 1)), device=dev, requires_grad=True) for t in range(ns)]
    loss = 0.0
    for t in range(ns):
        q, var, tac = StepSimFunction.apply(us[t], fs, sim, True)
        assert rel_err(q[0].detach().cpu().numpy(), g["q"][t]) <= 1e-9
        loss = loss + (q * torch.tensor(g["df_dq"][t], device=dev)).sum() + (var * torch.tensor(g["df_dvar"][t], device=dev)).sum() \
            + (tac * torch.tensor(g["df_dtactile"][t], device=dev)).sum()
    loss.backward()
    for t in range(ns):
        assert rel_err(us[t].grad[1].cpu().numpy(), g["df_du"][t].sum(axis=0)) <= 1e-6, t


def test_compat_simulation_dclaw_surface():
    """The calls R/envs/dclaw_rotate_env.py and tactile_insertion_env.py make beyond the TactilePush set:
    get_tactile_flow_images (DH/Robot.cpp:372-387), get_tactile_image_pos, update_tactile_parameters /
    update_contact_parameters / update_joint_damping (domain randomisation), export_replay."""
    from tactilesimulation_b200 import TactileSimError
    from tactilesimulation_b200.redmax import Simulation
    g = np.load(os.path.join(GOLDEN, "dclaw_episodic_s0.npz"))
    sc = _scene(g)
    for i, s in enumerate(sc.sensors):
        s.name = f"finger{i}"
    sim = Simulation(sc)
    assert (sim.ndof_r, sim.ndof_u, sim.ndof_var, sim.ndof_tactile) == (10, 9, 12, 2718)
    t = 12
    sim.set_state_init(g["q"][t], g["qd"][t])
    sim.reset(backward_flag=False)
    vec = sim.get_tactile_force_vector()
    assert rel_err(vec, g["tactile"][t]) <= 1e-8 and np.abs(vec).max() > 0
    imgs = sim.get_tactile_flow_images()
    assert len(imgs) == 3 and all(im.shape[2] == 3 for im in imgs)
    for i, sen in enumerate(sc.sensors):                     # reference semantics: a later marker overwrites
        ref = np.zeros_like(imgs[i])
        for m, (r, c) in enumerate(sen.image_pos):
            ref[r, c] = vec.reshape(3, 302, 3)[i, m]
        assert np.array_equal(imgs[i], ref) and imgs[i].shape[:2] == (20, 20)
    nf = np.array(sim.get_tactile_normal_force("finger1"))
    assert nf.shape == (302,) and np.allclose(nf, vec.reshape(3, 302, 3)[1, :, 2])
    s1 = sc.sensors[1]
    sim.update_tactile_parameters("finger1", s1.kn * 2.0, s1.kt * 2.0, s1.mu, s1.damping * 2.0)
    vec2 = sim.get_tactile_force_vector().reshape(3, 302, 3)
    assert np.allclose(vec2[0], vec.reshape(3, 302, 3)[0]) and not np.allclose(vec2[1], vec.reshape(3, 302, 3)[1])
    with pytest.raises(TactileSimError):
        sim.update_tactile_parameters("nope", 1, 1, 1, 1)
    with pytest.raises(TactileSimError):
        sim.update_contact_parameters("a", "b", 1, 1, 1, 1)
    sim.update_joint_damping(sc.joint_names[3], 0.5)
    sim.reset(backward_flag=True)
    sim.set_u(g["u"][t])
    sim.forward(3)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        n_mesh = sim.export_replay(os.path.join(d, "replay"))          # the reference's folder format (tests/test_replay_export.py)
        assert sorted(f for f in os.listdir(os.path.join(d, "replay")) if f.endswith(".txt")) == ["0.txt", "1.txt", "2.txt", "3.txt"]
        assert open(os.path.join(d, "replay", "3.txt")).readline().strip() == str(n_mesh)
