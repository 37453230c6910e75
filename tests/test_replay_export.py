"""Replay export: tactilesimulation_b200.replay writes the reference's export_replay format
(DH/Simulation.cpp:2037-2120).  The frame files -- world transforms of every body, render-only object, sensor and
end-effector, "%.6lf" -- are compared with the files the UNMODIFIED reference writes for the same q history, scene by
scene (all joint families: revolute / planar / translational / fixed, free3d-euler, free3d-exp, free2d, spherical)."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import GOLDEN, REF_DIR

A = os.path.join(REF_DIR, "assets")
SCENES = {"pusher": (os.path.join(A, "pusher", "pusher.xml"), "pusher13x10_episodic_s0"),
          "insertion": (os.path.join(A, "tactile_insertion", "tactile_insertion.xml"), "insertion_episodic_s0"),
          "dclaw": (os.path.join(A, "dclaw_rotate", "dclaw_torque_control.xml"), "dclaw_episodic_s0"),
          "rolling_ball": (os.path.join(A, "tactile_pad", "tactile_pad_40x40.xml"), "rollingball_bdf2_s0"),
          "free2d": (os.path.join(A, "synthetic", "free2d_plate.xml"), "free2d_plate_bdf1_s0"),
          "spherical": (os.path.join(A, "synthetic", "spherical_exp_bdf2.xml"), "spherical_exp_bdf2_s0")}

needs_ref = pytest.mark.skipif(not os.path.exists(SCENES["pusher"][0]) or not os.path.isdir(REF_DIR) or
                               not any(f.startswith("redmax_py") and f.endswith(".so") for f in os.listdir(REF_DIR)),
                               reason="oracle/_ref not built")


def _frames(folder):
    out = []
    i = 0
    while os.path.exists(os.path.join(folder, "%d.txt" % i)):
        toks = open(os.path.join(folder, "%d.txt" % i)).read().split()
        out.append((int(toks[0]), np.array([float(x) for x in toks[1:]])))
        i += 1
    return out


@needs_ref
@pytest.mark.parametrize("name", sorted(SCENES))
def test_frame_files_equal_the_reference_export(name, tmp_path):
    xml, case = SCENES[name]
    if not os.path.exists(xml):
        pytest.skip(xml + " not generated (tests/golden/make_golden.py writes the synthetic scenes)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import redmax_py
    from tactilesimulation_b200 import replay
    from tactilesimulation_b200.scene import compile_scene
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    sim = redmax_py.Simulation(xml)
    sim.set_state_init(g["q0"], g["qd0"])
    sim.reset(False)
    T = min(12, g["u"].shape[0])
    qh = [np.array(g["q0"])]
    for t in range(T):
        sim.set_u(g["u"][t])
        sim.forward(1)
        qh.append(np.array(sim.get_q()))
    ref_dir, new_dir = tmp_path / "ref", tmp_path / "new"
    os.makedirs(ref_dir / "meshes")
    sim.export_replay(str(ref_dir))
    sc = compile_scene(xml)
    n_mesh = replay.export_replay(sc, np.array(qh), str(new_dir))
    fr, fn = _frames(str(ref_dir)), _frames(str(new_dir))
    assert len(fr) == len(fn) == T + 1
    # Bodies built from a mesh sit in the principal-axes frame of their inertia, whose axes are defined up to sign: the
    # reference takes the signs of Eigen's eigenvectors (DH/Body/BodyMeshObj.cpp), the scene compiler numpy's; the
    # dynamics do not see the difference (diagonal inertia), the exported frame may differ by a flip of two axes.
    meshy = [b for b in range(sc.nj) if sc.shape[b] == 0]
    for (cr, vr), (cn, vn) in zip(fr, fn):
        assert cr == cn == n_mesh
        assert vr.shape == vn.shape == (16 * n_mesh,)
        Er, En = vr.reshape(-1, 4, 4), vn.reshape(-1, 4, 4)
        for k in range(n_mesh):
            if k in meshy:
                assert np.abs(Er[k][:, 3] - En[k][:, 3]).max() <= 2e-6, k
                assert np.abs(np.abs(Er[k][:3, :3]) - np.abs(En[k][:3, :3])).max() <= 2e-6, k
                assert np.linalg.det(En[k][:3, :3]) > 0.99
            else:
                assert np.abs(Er[k] - En[k]).max() <= 2e-6, k             # "%.6lf" text on both sides
    assert len([f for f in os.listdir(new_dir / "meshes") if f.endswith(".obj")]) == len(os.listdir(ref_dir / "meshes")) == n_mesh


@pytest.mark.gpu
def test_simulation_export_replay_writes_the_rollout(tmp_path):
    """The drop-in Simulation keeps the q history of its rollout (like Simulation::_q_his) and exports it."""
    from tactilesimulation_b200 import replay
    from tactilesimulation_b200.layout import scene_from_blob
    from tactilesimulation_b200.redmax import Simulation
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz"))
    sim = Simulation(scene_from_blob(g["ibuf"], g["dbuf"]))
    sim.set_state_init(g["q0"], g["qd0"])
    sim.reset(False)
    for t in range(5):
        sim.set_u(g["u"][t])
        sim.forward(2)
    n_mesh = sim.export_replay(str(tmp_path))
    frames = _frames(str(tmp_path))
    assert len(frames) == 11 and frames[0][0] == n_mesh
    E_last = frames[-1][1][:16 * sim.scene.nj].reshape(-1, 4, 4)
    _, E_0i = replay.frames(sim.scene, sim.get_q())
    assert np.abs(E_last - np.array(E_0i)).max() <= 2e-6
