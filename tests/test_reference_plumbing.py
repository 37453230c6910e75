"""BASELINE configs[0] -- RollingBallExp test_sim_speed.py, 1 env, CPU DiffRedMax -- is the reference's own
CPU-runnable case: it is run on the UNMODIFIED reference (oracle/_ref, built by oracle/build_ref.sh) as a plumbing
check of the reference arm.  The B200 path does not implement its scene (BDF2, sphere SDF, free3d-exp joint) and
must say so by name instead of falling back."""
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
def test_b200_path_rejects_the_rolling_ball_scene_by_name():
    from tactilesimulation_b200.scene import SceneError, compile_scene
    with pytest.raises(SceneError) as e:
        compile_scene(XML).pack()
    assert "not supported" in str(e.value)
