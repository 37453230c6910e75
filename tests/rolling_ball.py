"""Shared checks of the rolling-ball scene (examples/RollingBallExp/test_sim_speed.py, BASELINE configs[0]) against
tests/golden/rollingball_bdf2_s0.npz, used by the CPU harness test and by the GPU parity test."""
import numpy as np

from tests.conftest import rel_err

# stated tolerances: BDF2 / SDIRK2 stages are solved by the reference's own Newton iteration (same iterates up to
# fp64 reordering of the residual), so the parity bar of the BDF1 scenes applies
TOL_STATE, TOL_TACTILE = 1e-9, 1e-8


def tactile_rows(T, every):
    rows = np.full(T, -1, dtype=np.int32)
    rows[::every] = np.arange(len(rows[::every]), dtype=np.int32)
    return rows


def full_resolution_blob(ibuf, dbuf, res=200):
    """The real 200x200 scene rebuilt from the 40x40 fixture blob: a rect_array sensor is a regular grid between its
    corner markers (DH/Sensor/TactileSensorRectArray.cpp:43-68), nothing else differs."""
    from tactilesimulation_b200.layout import scene_from_blob
    sc = scene_from_blob(ibuf, dbuf)
    s = sc.sensors[0]
    r0 = int(round(np.sqrt(len(s.pos))))
    p0, p1 = s.pos[0], s.pos[-1]
    ax0, ax1 = s.axis0[0], s.axis1[0]
    l0, l1 = float((p1 - p0) @ ax0), float((p1 - p0) @ ax1)
    s0, s1 = l0 / (res - 1) * ax0, l1 / (res - 1) * ax1
    assert r0 * r0 == len(s.pos)
    pos = np.array([p0 + s0 * i + s1 * j for i in range(res) for j in range(res)])
    M = len(pos)
    s.pos = pos
    s.axis0, s.axis1, s.normal = np.tile(ax0, (M, 1)), np.tile(ax1, (M, 1)), np.tile(s.normal[0], (M, 1))
    s.image_pos = np.array([[i, j] for i in range(res) for j in range(res)], dtype=np.int64)
    return sc.pack()


def ids_of(words):
    out = []
    for w, word in enumerate(words):
        word = int(word) & 0xffffffff
        while word:
            b = (word & -word).bit_length() - 1
            out.append(32 * w + b)
            word &= word - 1
    return out


def check_trajectory(q, qd, status, cmask, tactile, marker_body, g, steps=None):
    """q, qd [T,n]; status [T]; cmask [T,W]; tactile [frames, 3M]; marker_body [frames, M] of ONE environment."""
    T = q.shape[0] if steps is None else steps
    every = int(g["tactile_every"])
    assert int((np.asarray(status[:T]) >> 16).max()) == 0
    for t in range(T):
        assert rel_err(q[t], g["q"][t]) <= TOL_STATE, t
        assert rel_err(qd[t], g["qd"][t]) <= TOL_STATE, t
        if cmask is not None:
            assert ids_of(cmask[t, 0:1]) == [int(x) for x in g["ground_ids"][t] if x >= 0], t
            assert ids_of(cmask[t, 1:]) == [int(x) for x in g["gp_ids"][t] if x >= 0], t
    for f in range((T + every - 1) // every):
        assert rel_err(tactile[f], g["tactile"][f]) <= TOL_TACTILE, f
        if marker_body is not None:
            assert np.array_equal(marker_body[f], g["marker_body"][f]), f
    assert float(np.abs(g["tactile"]).max()) > 0 and int((g["gp_ids"] >= 0).sum()) > 0


def check_full_resolution_frames(tactile_frames, g):
    """tactile_frames [4, 120000]: the 200x200 field at the steps g['frames200']."""
    for k in range(len(g["frames200"])):
        ref = np.zeros(tactile_frames.shape[1])
        idx = g["tactile200_idx"][k]
        ref[idx[idx >= 0]] = g["tactile200_val"][k][idx >= 0]
        assert rel_err(tactile_frames[k], ref) <= TOL_TACTILE, k
        assert np.array_equal(np.nonzero(tactile_frames[k])[0], np.nonzero(ref)[0]), k
