"""BASELINE configs[0] -- RollingBallExp test_sim_speed.py, 1 env, CPU DiffRedMax -- is the reference's own
CPU-runnable case: it is run on the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh) as a plumbing
check of the reference arm.  The B200 path runs the same scene on kernel variant 17 (BDF2 + SDIRK2 start-up, sphere SDF,
free3d-exp joint: tests/test_emu_vs_golden.py, tests/test_gpu_parity.py); what it still does not implement it must
reject by name instead of falling back."""
import os

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "oracle", "_ref")
XML = os.path.join(REF, "assets", "tactile_pad", "tactile_pad.xml")

needs_ref = pytest.mark.skipif(
    not (os.path.exists(XML) and any(f.startswith("redmax_py") and f.endswith(".so") for f in os.listdir(REF))),
    reason="oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")


@needs_ref
def test_rolling_ball_speed_test_runs_on_the_reference():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_config0
    r = ref_config0.run()
    assert (r["ndof_r"], r["ndof_u"], r["ndof_tactile"]) == (9, 3, 120000)
    assert r["steps"] == 350 and r["fps"] > 1.0
    assert r["peak_tactile_force"] > 0.0 and np.isfinite(r["q_final"]).all()


@needs_ref
def test_b200_scene_compiler_reads_the_rolling_ball_scene():
    """The native XML compiler on the reference's own file; and the fixture helper that rebuilds the 200x200 scene from
    the 40x40 blob (used by the GPU tests, where the XML does not exist) gives exactly the compiled scene."""
    from tactilesimulation_b200.scene import compile_scene
    from tests import rolling_ball as rb
    sc = compile_scene(XML)
    assert (sc.ndof_r, sc.ndof_u, sc.ndof_tactile, sc.integrator) == (9, 3, 120000, "BDF2")
    ib, db = sc.pack()
    g = np.load(os.path.join(ROOT, "tests", "golden", "rollingball_bdf2_s0.npz"))
    ib2, db2 = rb.full_resolution_blob(g["ibuf"], g["dbuf"])
    assert np.array_equal(ib, ib2) and np.allclose(db, db2, rtol=0, atol=1e-15)


def test_b200_path_rejects_unsupported_features_by_name(tmp_path):
    """What is still not built (cuboid-cuboid contact, DH/Force/ForceCuboidCuboidContact.cpp) is refused by name."""
    from tactilesimulation_b200.scene import SceneError, compile_scene
    xml = tmp_path / "boxes.xml"
    xml.write_text('''<redmax model="x"><option integrator="BDF1" timestep="5e-3" gravity="0 0 -9.8"/>
<robot><link name="a"><joint name="j" type="free2d" pos="0 0 0" quat="1 0 0 0"/>
<body name="b" type="capsule" pos="0 0 0" quat="1 0 0 0" radius="0.1" length="0.2" density="1"/></link></robot>
<robot><link name="c"><joint name="k" type="free3d" pos="0 0 1" quat="1 0 0 0"/>
<body name="d" type="cuboid" pos="0 0 0" quat="1 0 0 0" size="0.1 0.1 0.1" density="1"/></link></robot>
<contact><cuboid_cuboid_contact body1="d" body2="d"/></contact></redmax>''')
    with pytest.raises(SceneError) as e:
        compile_scene(str(xml))
    assert "not supported" in str(e.value)
