"""BASELINE configs[0]: RollingBallExp test_sim_speed.py logic (R/examples/RollingBallExp/test_sim_speed.py:36-104) on
the UNMODIFIED reference built by oracle/build_ref.sh -- 1 env, CPU, no GPU: a plumbing check of the reference
arm (the B200 path runs the same scene on kernel variant 17: tools/rolling_ball_probe.py is the GPU counterpart).
Usage: python tools/ref_config0.py   -> prints one JSON line {"fps": ..., "steps": 350, ...}"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "oracle", "_ref")


def run():
    sys.path.insert(0, REF)
    import redmax_py as redmax
    sim = redmax.Simulation(os.path.join(REF, "assets", "tactile_pad", "tactile_pad.xml"))
    action_array = [np.array([0., 0., 0.2]), np.array([0.1, 0., 0.2]), np.array([-0.2, 0., 0.2]),
                    np.array([0., 0.1, 0.2]), np.array([0., -0.2, 0.2])]
    steps_array = [0, 100, 150, 200, 250, 350]
    actions = [action_array[i] for i in range(5) for _ in range(steps_array[i], steps_array[i + 1])]
    sim.reset(backward_flag=False)
    image_pos = sim.get_tactile_image_pos(name="pad")
    rows = max(p[0] for p in image_pos) + 1
    cols = max(p[1] for p in image_pos) + 1
    t0 = time.time()
    peak = 0.0
    for i, a in enumerate(actions):
        sim.set_u(a)
        sim.forward(1, verbose=False, test_derivatives=False)
        if i % 5 == 0:
            f = sim.get_tactile_force_vector().copy()
            assert rows * cols == f.shape[0] // 3
            peak = max(peak, float(np.abs(f).max()))
    dt = time.time() - t0
    return {"config": "RollingBallExp test_sim_speed.py, 1 env, CPU DiffRedMax", "steps": steps_array[-1], "fps": steps_array[-1] / dt,
            "ndof_r": sim.ndof_r, "ndof_u": sim.ndof_u, "ndof_tactile": sim.ndof_tactile, "peak_tactile_force": peak,
            "q_final": [float(x) for x in sim.get_q()]}


if __name__ == "__main__":
    print(json.dumps(run()))
