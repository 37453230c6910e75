"""CPU check of the restatement that tests/test_gpu_lu.py holds the kernels' solve against: the row-exchanging elimination
in the reference's order (Eigen partialPivLu, DH/Simulation.cpp:1178) solves the systems of that test's batch."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(__file__)


def _gpu_lu_module():
    spec = importlib.util.spec_from_file_location("_gpu_lu", os.path.join(HERE, "test_gpu_lu.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_partial_pivoting_restatement_solves_its_batch():
    m = _gpu_lu_module()
    for n in (8, 16):
        A, b = m.systems(n, 301, seed=n)
        exchanges = 0
        for s in range(len(b)):
            x, e = m.partial_piv_solve(A[s], b[s])
            exchanges += e
            ref = np.linalg.solve(A[s], b[s])
            assert np.abs(x - ref).max() <= 1e-11 * (np.abs(ref).max() + 1e-300) * max(1.0, np.linalg.cond(A[s])), (n, s)
        assert exchanges > len(b)


def test_first_maximum_wins_on_ties():
    m = _gpu_lu_module()
    A = np.array([[1.0, 2.0, 0.0], [-1.0, -1.5, 1.0], [1.0, 4.0, 1.0]])
    b = np.array([1.0, 2.0, 3.0])
    x, e = m.partial_piv_solve(A, b)
    assert e == 1                       # column 0 ties everywhere: no exchange there; one exchange in column 1
    assert np.allclose(A @ x, b)
