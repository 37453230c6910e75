"""The domain-randomisation calls of the reference's envs -- update_joint_damping / update_body_size /
update_endeffector_position / update_joint_location (R/envs/dclaw_rotate_env.py:173-178), update_body_density
(R/envs/stable_grasp_env.py:122); DH/Robot.cpp:571-650 -- as host-side scene edits (tactilesimulation_b200.scene.update_*,
which the drop-in Simulation calls before uploading a new handle).  The fixture holds the ORIGINAL scenes and rollouts of the
reference AFTER its own update calls; the kernel math is run on the edited scene (CPU harness)."""
import os

import numpy as np

from tactilesimulation_b200 import scene as S
from tactilesimulation_b200.layout import scene_from_blob
from tests import emu_lib
from tests.conftest import GOLDEN, rel_err


def test_dclaw_randomisation_matches_reference():
    g = np.load(os.path.join(GOLDEN, "randomized_updates_s0.npz"))
    sc = scene_from_blob(g["dclaw_ibuf"], g["dclaw_dbuf"])
    sc.joint_names = [str(x) for x in g["dclaw_joint_names"]]
    sc.body_names = [str(x) for x in g["dclaw_body_names"]]
    for e, name in zip(sc.end_effectors, g["dclaw_ee_names"]):
        e["name"] = str(name)
    sc.damping[sc.joint_names.index("cap")] = float(g["dclaw_damping"])          # update_joint_damping
    S.update_body_size(sc, "cap", g["dclaw_size"])
    [e for e in sc.end_effectors if e["name"] == "cap"][0]["pos"] = g["dclaw_ee"].copy()   # update_endeffector_position
    S.update_joint_location(sc, "cap", g["dclaw_loc"])
    ib, db = sc.pack()
    T = g["dclaw_u"].shape[0]
    out = emu_lib.forward(ib, db, g["dclaw_q0"], np.zeros(10), g["dclaw_u"][:, None, :], want_masks=False)
    assert int((out["status"] >> 16).max()) == 0
    assert float(np.abs(g["dclaw_tactile"]).max()) > 0
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["dclaw_q"][t]) <= 1e-9, t
        assert rel_err(out["var"][t, 0], g["dclaw_var"][t]) <= 1e-9, t
        if t % 5 == 4:
            assert rel_err(out["tactile"][t, 0], g["dclaw_tactile"][t // 5]) <= 1e-8, t
    # and the edit did change the rollout: the untouched scene differs
    ref = emu_lib.forward(g["dclaw_ibuf"], g["dclaw_dbuf"], g["dclaw_q0"], np.zeros(10), g["dclaw_u"][:, None, :], want_masks=False)
    assert rel_err(ref["q"][-1, 0], g["dclaw_q"][-1]) > 1e-6


def test_stable_grasp_densities_match_reference():
    g = np.load(os.path.join(GOLDEN, "randomized_updates_s0.npz"))
    sc = scene_from_blob(g["sg_ibuf"], g["sg_dbuf"])
    sc.body_names = [str(x) for x in g["sg_body_names"]]
    for i in range(11):          # (the scene's boxes are box_1 .. box_11: "box_0" matches nothing and is ignored, as in the reference)
        S.update_body_density(sc, "box_%d" % i, float(g["sg_dens"][i]))
    ib, db = sc.pack()
    T = g["sg_u"].shape[0]
    out = emu_lib.forward(ib, db, g["sg_q0"], np.zeros(12), g["sg_u"][:, None, :], want_masks=False)
    assert int((out["status"] >> 16).max()) == 0
    assert float(np.abs(g["sg_tactile"]).max()) > 0
    for t in range(T):
        assert rel_err(out["q"][t, 0], g["sg_q"][t]) <= 1e-9, t
        if t % 5 == 4:
            assert rel_err(out["tactile"][t, 0], g["sg_tactile"][t // 5]) <= 1e-8, t
    ref = emu_lib.forward(g["sg_ibuf"], g["sg_dbuf"], g["sg_q0"], np.zeros(12), g["sg_u"][:, None, :], want_masks=False)
    assert rel_err(ref["q"][-1, 0], g["sg_q"][-1]) > 1e-6


def test_updates_agree_with_recompiling_the_scene(tmp_path):
    """update_* on a compiled scene = compiling the scene with the new attributes (values exact in fp32, which the XML
    loader rounds through).  (LOCAL-frame joints: the children of a moved joint keep their transform relative to it, as in
    the reference, which a WORLD-frame child of a recompiled scene would not.)"""
    base = '''<redmax model="u"><option integrator="BDF1" timestep="5e-3" unit="m-kg" gravity="0 0 -9.8"/>
<ground pos="0 0 0" normal="0 0 1"/>
<robot><link name="a"><joint name="ja" type="translational" pos="{ja}" quat="1 0 0 0" damping="0.5"/>
<body name="ba" type="cuboid" size="{sa}" pos="0 0 0" quat="1 0 0 0" density="{da}" general_contact_resolution="3 2 2"/>
<link name="c"><joint name="jc" type="revolute" axis="0 0 1" pos="0.125 0 0.25" quat="1 0 0 0"/>
<body name="bc" type="cylinder" radius="{rc}" length="{lc}" pos="0 0 0" quat="1 0 0 0" density="250"/></link></link></robot>
<contact><ground_contact body="ba" kn="100" kt="1" mu="0.5" damping="1"/>
<general_primitive_contact general_body="ba" primitive_body="bc" kn="10" kt="1" mu="0.5" damping="1"/></contact></redmax>'''
    a = dict(ja="0 0 0.5", sa="0.25 0.125 0.0625", da="500", rc="0.03125", lc="0.5")
    b = dict(ja="0.25 -0.125 0.75", sa="0.5 0.125 0.25", da="1024", rc="0.0625", lc="0.25")
    pa, pb = tmp_path / "a.xml", tmp_path / "b.xml"
    pa.write_text(base.format(**a))
    pb.write_text(base.format(**b))
    sc = S.compile_scene(str(pa))
    S.update_joint_location(sc, "ja", [0.25, -0.125, 0.75])
    S.update_body_size(sc, "ba", [0.5, 0.125, 0.25])
    S.update_body_density(sc, "ba", 1024.0)
    S.update_body_size(sc, "bc", [0.25, 0.0625])
    want = S.compile_scene(str(pb))
    ia, da_ = sc.pack()
    ib, db = want.pack()
    assert np.array_equal(ia, ib)
    assert np.allclose(da_, db, rtol=1e-15, atol=0)
