"""bench.py --impl reference: the reference arm of every workload prints the contract's JSON line (CPU only; tiny samples)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "oracle", "_ref")
needs_ref = pytest.mark.skipif(not os.path.isdir(REF) or not any(f.startswith("redmax_py") and f.endswith(".so") for f in os.listdir(REF)),
                               reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("workload,extra", [("push", ["--horizon", "12"]), ("push_fwd", ["--horizon", "12"]), ("dclaw", ["--horizon", "8"]),
                                            ("insertion", ["--horizon", "10"]), ("stepsim", ["--gym-steps", "2"])])
def test_reference_arm_line(workload, extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1",
                          "--warmup", "0", "--ref-envs-per-core", "1"] + extra, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.strip().splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["unit"] == ("gym-steps/s" if workload == "stepsim" else "env-steps/s")
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "-O3 -DNDEBUG" in line["cpu_baseline"]["build"]


def test_workload_inputs_are_seeded_and_shaped():
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    for name, wl in bench.WORKLOADS.items():
        g = np.load(os.path.join(ROOT, "tests", "golden", wl["case"] + ".npz"))
        a = bench.workload_inputs(name, g, 3, 7, seed=5)
        b = bench.workload_inputs(name, g, 3, 7, seed=5)
        assert a[0].shape == (3, len(g["q0"])) and a[2].shape == (7, 3, g["u"].shape[1])
        assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3]))
        c = bench.workload_inputs(name, g, 3, 7, seed=6)
        assert not np.array_equal(a[2], c[2])
