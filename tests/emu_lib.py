"""ctypes wrapper of the TEST-ONLY host harness (tests/emu/emu.cpp): the kernel's per-lane math
compiled by g++ and run with a one-lane tile.  Builds the library on first use."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu.cpp")
LIB = os.path.join(HERE, "emu", "libtsim_emu_v%d.so")
CORE = os.path.join(HERE, "..", "tactilesimulation_b200", "csrc")

_libs = {}


def lib(variant=8):
    """The harness compiled with the capacities of kernel variant 8, 16 or 17 (csrc/kernel_layout.h)."""
    if variant not in _libs:
        path = LIB % variant
        deps = [SRC] + [os.path.join(CORE, f) for f in ("sim_core.cuh", "dual.cuh", "scene_layout.h", "kernel_layout.h", "scene_lower.h")]
        if not os.path.exists(path) or any(os.path.getmtime(d) > os.path.getmtime(path) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++14", "-shared", "-fPIC", f"-DTS_VARIANT={variant}", "-o", path, SRC])
        _libs[variant] = ctypes.CDLL(path)
    return _libs[variant]


def lib_for(ibuf, dbuf):
    """(library, contact bitmask words) of the smallest variant that accepts the scene."""
    for variant in (8, 16, 17):
        l = lib(variant)
        w = l.emu_cmask_words(_p(ibuf, ctypes.c_int32), _p(dbuf))
        if w >= 0:
            return l, w
    raise RuntimeError("scene rejected by every kernel variant")


def _p(a, t=ctypes.c_double):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(t))


def _rows(x):
    return None if x is None else np.ascontiguousarray(x, dtype=np.int32)


def forward(ibuf, dbuf, q0, qd0, u, grad=False, var_row=None, tac_row=None, want_masks=True, hist=None, steps_done=0):
    """u: [T,B,nu].  Returns dict of trajectories.  hist = (q_prev, qd_prev) [B,n] arrays (in/out) and steps_done:
    the multistep state of tsim_forward_multistep."""
    ibuf = np.ascontiguousarray(ibuf, dtype=np.int32)
    dbuf = np.ascontiguousarray(dbuf, dtype=np.float64)
    n, nu, nee, M = int(ibuf[3]), int(ibuf[4]), int(ibuf[5]), int(ibuf[6])
    u = np.ascontiguousarray(u, dtype=np.float64)
    T, B = u.shape[0], u.shape[1]
    q = np.ascontiguousarray(np.broadcast_to(q0, (B, n)), dtype=np.float64).copy()
    qd = np.ascontiguousarray(np.broadcast_to(qd0, (B, n)), dtype=np.float64).copy()
    vr, tr = _rows(var_row), _rows(tac_row)
    nv = T if vr is None else int(vr.max()) + 1
    nt = T if tr is None else int(tr.max()) + 1
    l, cmw = lib_for(ibuf, dbuf)
    out = dict(q=np.zeros((T, B, n)), qd=np.zeros((T, B, n)), var=np.zeros((nv, B, 3 * nee)),
               tactile=np.zeros((nt, B, 3 * M)), status=np.zeros((T, B), dtype=np.int32),
               tape=np.zeros((T, B, 3 * n * n + nu)) if grad else None,
               cmask=np.zeros((T, B, cmw), dtype=np.uint32) if want_masks else None,
               marker_body=np.zeros((nt, B, M), dtype=np.int32))
    l.emu_forward(_p(ibuf, ctypes.c_int32), _p(dbuf), B, T, _p(q), _p(qd), _p(u), ctypes.c_int64(B * nu),
                      _p(out["q"]), _p(out["qd"]), _p(out["var"]), _p(vr, ctypes.c_int32), _p(out["tactile"]),
                      _p(tr, ctypes.c_int32), _p(out["tape"]), _p(out["status"], ctypes.c_int32),
                      _p(out["cmask"], ctypes.c_uint32), _p(out["marker_body"], ctypes.c_int32),
                      _p(None if hist is None else hist[0]), _p(None if hist is None else hist[1]), ctypes.c_int32(steps_done))
    out["q_final"], out["qd_final"] = q, qd
    return out


def backward(ibuf, dbuf, fwd, u, df_dq=None, df_dvar=None, df_dtac=None, dq_row=None, dvar_row=None, dtac_row=None,
             carry=None, want_q0=True):
    ibuf = np.ascontiguousarray(ibuf, dtype=np.int32)
    dbuf = np.ascontiguousarray(dbuf, dtype=np.float64)
    n, nu = int(ibuf[3]), int(ibuf[4])
    u = np.ascontiguousarray(u, dtype=np.float64)
    T, B = u.shape[0], u.shape[1]
    if carry is None:
        carry = np.zeros((B, 2, n))
    cots = [None if c is None else np.ascontiguousarray(c, dtype=np.float64) for c in (df_dq, df_dvar, df_dtac)]
    rows = [_rows(r) for r in (dq_row, dvar_row, dtac_row)]
    df_du = np.zeros((T, B, nu))
    dq0 = np.zeros((B, n)) if want_q0 else None
    dqd0 = np.zeros((B, n)) if want_q0 else None
    lib_for(ibuf, dbuf)[0].emu_backward(_p(ibuf, ctypes.c_int32), _p(dbuf), B, T, _p(fwd["q"]), _p(fwd["qd"]), _p(u),
                       ctypes.c_int64(B * nu), _p(fwd["tape"]), _p(cots[0]), _p(rows[0], ctypes.c_int32), _p(cots[1]),
                       _p(rows[1], ctypes.c_int32), _p(cots[2]), _p(rows[2], ctypes.c_int32), _p(carry), _p(df_du),
                       _p(dq0), _p(dqd0))
    return dict(df_du=df_du, df_dq0=dq0, df_dqdot0=dqd0, carry=carry)


def mask_to_ids(words):
    ids = []
    for w, word in enumerate(words):
        for b in range(32):
            if (int(word) >> b) & 1:
                ids.append(32 * w + b)
    return ids
