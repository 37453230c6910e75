"""The kernels' n x n solve (csrc/sim_core.cuh lu_rows_solve_pivot: rows stay in their lanes, partial pivoting through the
tile's scratch) against a plain restatement of Eigen's partialPivLu order (DH/Simulation.cpp:1178: first row of maximal
|a| in the column, row exchange, elimination, back substitution), through the C ABI (tsim_debug_lu_solve)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def partial_piv_solve(A, b):
    """Row-exchanging elimination in the reference's order; also returns the number of exchanges."""
    A, b = A.copy(), b.copy()
    n = len(b)
    exchanges = 0
    for j in range(n):
        p = j + int(np.argmax(np.abs(A[j:, j])))          # first index of the maximum
        if p != j:
            A[[j, p]] = A[[p, j]]
            b[[j, p]] = b[[p, j]]
            exchanges += 1
        for i in range(j + 1, n):
            l = A[i, j] / A[j, j]
            A[i, j + 1:] -= l * A[j, j + 1:]
            b[i] -= l * b[j]
    for k in range(n - 1, -1, -1):
        b[k] = b[k] / A[k, k]
        b[:k] -= A[:k, k] * b[k]
    return b, exchanges


def systems(n, nsys, seed):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(nsys, n, n))
    b = rng.normal(size=(nsys, n))
    # a third of them mass dominated (no exchange), a third with exact ties in the pivot column, a third padded with the
    # identity the way the kernels pad a system smaller than the capacity
    for s in range(0, nsys, 3):
        A[s] += 8.0 * np.eye(n)
    for s in range(1, nsys, 3):
        A[s, :, 0] = np.where(rng.uniform(size=n) < 0.5, 1.5, -1.5)
        A[s, 2:, 1] = A[s, 1, 1]
    for s in range(2, nsys, 3):
        m = n - 1 - (s % 3)
        A[s, m:, :] = 0.0
        A[s, :, m:] = 0.0
        A[s, np.arange(m, n), np.arange(m, n)] = 1.0
        b[s, m:] = 0.0
    return A, b


@pytest.mark.parametrize("n", [8, 16])
def test_row_owner_pivoting_solve_matches_partial_pivoting(n):
    from tactilesimulation_b200 import _lib
    A, b = systems(n, 301, seed=n)
    x = _lib.lu_solve(A, b)
    n_exch = 0
    for s in range(len(b)):
        ref, e = partial_piv_solve(A[s], b[s])
        n_exch += e
        scale = np.abs(ref).max() + 1e-300
        # same pivots, same operations: the difference is the fused multiply-adds of the GPU (an ulp per operation,
        # amplified by the conditioning of the random matrices: <= 6e3 here)
        assert np.abs(x[s] - ref).max() <= 1e-10 * scale, s
        assert np.abs(A[s] @ x[s] - b[s]).max() <= 1e-10 * (np.abs(A[s]).sum(axis=1).max() * scale + 1.0), s
    assert n_exch > len(b)          # the batch does exercise the exchanges


def test_lu_solve_rejects_other_sizes():
    from tactilesimulation_b200 import _lib
    with pytest.raises(_lib.TactileSimError):
        _lib.lu_solve(np.eye(5)[None], np.ones((1, 5)))
