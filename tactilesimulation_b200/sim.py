"""Batched, device-resident simulator handle: thin torch plumbing over the C ABI.

``BatchedSim`` owns a scene handle on one GPU and exposes the three kernels (forward / readout /
backward) on torch CUDA tensors.  All tensors are fp64, C-contiguous, laid out
``[step][env][component]``.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .scene import Scene, compile_scene


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class BatchedSim:
    def __init__(self, scene, device="cuda:0", lanes: Optional[int] = None):
        """lanes: lanes of a warp cooperating on one environment (None = the default of the kernel variant
        the scene runs on: one lane per reduced coordinate, 8 or 16)."""
        if not torch.cuda.is_available():
            raise _lib.TactileSimError("no CUDA device visible: tactilesimulation_b200 has no CPU fallback")
        self.lib = _lib.load()
        if isinstance(scene, str):
            scene = compile_scene(scene)
        if isinstance(scene, Scene):
            self.scene = scene
            ibuf, dbuf = scene.pack()
        else:
            self.scene = None
            ibuf, dbuf = scene
        self.ibuf = np.ascontiguousarray(ibuf, dtype=np.int32)
        self.dbuf = np.ascontiguousarray(dbuf, dtype=np.float64)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TactileSimError("device must be a CUDA device")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", idx)
        h = ctypes.c_void_p()
        _lib.check(self.lib.tsim_scene_create(self.ibuf.ctypes.data, self.ibuf.size, self.dbuf.ctypes.data,
                                              self.dbuf.size, idx, ctypes.byref(h)), self.lib)
        self.handle = h
        sizes = np.zeros(_lib.N_SIZES, dtype=np.int32)
        _lib.check(self.lib.tsim_scene_sizes(self.handle, sizes.ctypes.data), self.lib)
        self.nj, self.ndof_r, self.ndof_m, self.ndof_u, self.ndof_var, self.ndof_tactile, self.n_markers, \
            self.tape_doubles = (int(x) for x in sizes[:8])
        self.cmask_words = int(sizes[_lib.CMASK_WORDS])
        self.integrator = int(sizes[_lib.INTEGRATOR])      # _lib.INT_BDF1 / INT_BDF2 / INT_SDIRK2
        self.h = float(self.dbuf[0])
        self.lanes = None
        self.options = {}
        self.env_batch = 0
        self._row_cache = {}
        if lanes is not None:
            self.set_lanes(lanes)

    def set_option(self, key: int, value: int):
        """tsim_scene_set_option (include/tactilesim_b200.h): 0 = TSIM_OPT_LS_BATCH."""
        _lib.check(self.lib.tsim_scene_set_option(self.handle, key, value), self.lib)
        self.options[int(key)] = int(value)

    def set_env_scenes(self, ibufs, dbufs):
        """Per-environment parameters (tsim_scene_set_env_scenes): ibufs [B, n_int] int32, dbufs [B, n_dbl] float64 -- one
        packed scene of this handle's topology per environment; None / empty returns to one parameter set."""
        if ibufs is None or len(ibufs) == 0:
            _lib.check(self.lib.tsim_scene_set_env_scenes(self.handle, 0, None, 0, None, 0), self.lib)
            self.env_batch = 0
            return
        ib = np.ascontiguousarray(ibufs, dtype=np.int32)
        db = np.ascontiguousarray(dbufs, dtype=np.float64)
        if ib.ndim != 2 or db.ndim != 2 or ib.shape[0] != db.shape[0]:
            raise _lib.TactileSimError("set_env_scenes: expected ibufs [B, n_int] and dbufs [B, n_dbl]")
        _lib.check(self.lib.tsim_scene_set_env_scenes(self.handle, ib.shape[0], ib.ctypes.data, ib.shape[1],
                                                      db.ctypes.data, db.shape[1]), self.lib)
        self.env_batch = ib.shape[0]

    def kernel_times(self):
        """Device time (ms) of each kernel of the last forward / backward call on this handle (None: did not run);
        synchronises the device first.  tsim_scene_kernel_times."""
        torch.cuda.synchronize(self.device)
        ms = np.zeros(len(_lib.KERNELS), dtype=np.float64)
        _lib.check(self.lib.tsim_scene_kernel_times(self.handle, ms.ctypes.data), self.lib)
        return {k: (None if v < 0 else float(v)) for k, v in zip(_lib.KERNELS, ms)}

    def set_lanes(self, lanes: int):
        _lib.check(self.lib.tsim_scene_set_lanes(self.handle, lanes), self.lib)
        self.lanes = lanes

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                self.lib.tsim_scene_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _chk(self, t: Optional[torch.Tensor], shape, name, dtype=torch.float64):
        if t is None:
            return
        if t.device != self.device or t.dtype != dtype or not t.is_contiguous():
            raise _lib.TactileSimError(f"{name}: expected a contiguous {dtype} tensor on {self.device}")
        if tuple(t.shape) != tuple(shape):
            raise _lib.TactileSimError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")

    def _rows(self, rows, T: int):
        """row map [T] (int32 device tensor) and number of rows; None = identity."""
        if rows is None:
            return None, T
        if isinstance(rows, torch.Tensor):
            r = rows.to(torch.int32).cpu()
        else:
            r = torch.as_tensor(np.asarray(rows, dtype=np.int32))
        if r.numel() != T:
            raise _lib.TactileSimError("row map must have one entry per step")
        # (the same few maps come back every gym step of a StepSimFunction loop: keep their device copies)
        key = tuple(r.tolist())
        hit = self._row_cache.get(key)
        if hit is None:
            if len(self._row_cache) > 64:
                self._row_cache.clear()
            hit = (r.to(self.device), max((max(key) if key else -1) + 1, 0))
            self._row_cache[key] = hit
        return hit

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ kernels
    def forward(self, q: torch.Tensor, qd: torch.Tensor, u: torch.Tensor, T: int, grad: bool = False,
                var_rows=None, tac_rows=None, want_var=True, want_tactile=True, want_status=False,
                want_contacts=False, want_traj=True, q_prev: Optional[torch.Tensor] = None,
                qd_prev: Optional[torch.Tensor] = None, steps_done: int = 0):
        """Advance (q, qd) [B,n] in place by T steps.  u: [T,B,nu], or [B,nu] held for all T steps.
        BDF2 scenes (tsim_forward_multistep): q_prev, qd_prev [B,n] hold the state one step back (in/out) and
        steps_done the steps taken since reset; omit them for a whole trajectory from reset in one call.
        Returns a dict with q_traj, qd_traj [T,B,n], var [rows,B,nvar], tactile [rows,B,ntac] and, when
        ``grad``, the adjoint tape [T,B,3 n^2 + nu] (H, G0, G1 row-major, then d f_r/d u per control)."""
        B, n, nu = q.shape[0], self.ndof_r, self.ndof_u
        self._chk(q, (B, n), "q")
        self._chk(qd, (B, n), "qd")
        if u.dim() == 2:
            self._chk(u, (B, nu), "u")
            ustride = 0
        else:
            self._chk(u, (T, B, nu), "u")
            ustride = B * nu
        self._chk(q_prev, (B, n), "q_prev")
        self._chk(qd_prev, (B, n), "qd_prev")
        if grad and self.integrator != _lib.INT_BDF1:
            raise _lib.TactileSimError("the adjoint exists for BDF1 scenes only (as Simulation::backward of the reference)")
        dev = self.device
        out = {}
        need_traj = want_traj or grad
        out["q_traj"] = torch.empty((T, B, n), dtype=torch.float64, device=dev) if need_traj else None
        out["qd_traj"] = torch.empty((T, B, n), dtype=torch.float64, device=dev) if need_traj else None
        vr, nvr = self._rows(var_rows, T)
        tr, ntr = self._rows(tac_rows, T)
        out["var"] = (torch.zeros((nvr, B, self.ndof_var), dtype=torch.float64, device=dev)
                      if (want_var and self.ndof_var) else None)
        # every row of an identity row map is written by the kernel: no 8 GB memset in front of it
        alloc = torch.empty if tac_rows is None else torch.zeros
        out["tactile"] = (alloc((ntr, B, self.ndof_tactile), dtype=torch.float64, device=dev)
                          if (want_tactile and self.ndof_tactile) else None)
        out["tape"] = torch.empty((T, B, self.tape_doubles), dtype=torch.float64, device=dev) if grad else None
        out["status"] = torch.zeros((T, B), dtype=torch.int32, device=dev) if want_status else None
        out["contact_masks"] = torch.zeros((T, B, self.cmask_words), dtype=torch.int32, device=dev) if want_contacts else None
        out["marker_body"] = (torch.full((ntr, B, self.n_markers), -1, dtype=torch.int32, device=dev)
                              if (want_contacts and out["tactile"] is not None) else None)
        with torch.cuda.device(dev):
            _lib.check(self.lib.tsim_forward_multistep(
                self.handle, B, T, _ptr(q), _ptr(qd), _ptr(q_prev), _ptr(qd_prev), int(steps_done), _ptr(u), ustride,
                _ptr(out["q_traj"]), _ptr(out["qd_traj"]), _ptr(out["var"]), _ptr(vr), _ptr(out["tactile"]), _ptr(tr),
                _ptr(out["tape"]), _ptr(out["status"]), _ptr(out["contact_masks"]), _ptr(out["marker_body"]),
                self._stream()), self.lib)
        out["_keep"] = (vr, tr)
        return out

    def readout(self, q: torch.Tensor, qd: torch.Tensor, want_contacts=False):
        B = q.shape[0]
        self._chk(q, (B, self.ndof_r), "q")
        self._chk(qd, (B, self.ndof_r), "qd")
        dev = self.device
        var = torch.zeros((B, self.ndof_var), dtype=torch.float64, device=dev) if self.ndof_var else None
        tac = torch.zeros((B, self.ndof_tactile), dtype=torch.float64, device=dev) if self.ndof_tactile else None
        mb = torch.full((B, self.n_markers), -1, dtype=torch.int32, device=dev) if want_contacts else None
        cm = torch.zeros((B, self.cmask_words), dtype=torch.int32, device=dev) if want_contacts else None
        with torch.cuda.device(dev):
            _lib.check(self.lib.tsim_readout(self.handle, B, _ptr(q), _ptr(qd), _ptr(var), _ptr(tac), _ptr(mb),
                                             _ptr(cm), self._stream()), self.lib)
        return dict(var=var, tactile=tac, marker_body=mb, contact_masks=cm)

    def backward(self, fwd: dict, u: torch.Tensor, T: int, df_dq=None, df_dvar=None, df_dtactile=None,
                 dq_rows=None, dvar_rows=None, dtac_rows=None, carry: Optional[torch.Tensor] = None,
                 want_q0: bool = False):
        """Reverse sweep over the T steps of ``fwd`` (a dict returned by forward(grad=True)).
        Returns df_du [T,B,nu], the updated carry [B,2,n] and (optionally) df_dq0, df_dqdot0 [B,n]."""
        if fwd.get("tape") is None:
            raise _lib.TactileSimError("backward needs the tape of a forward(grad=True) call")
        B, n, nu = fwd["q_traj"].shape[1], self.ndof_r, self.ndof_u
        dev = self.device
        ustride = 0 if u.dim() == 2 else B * nu
        r0, _ = self._rows(dq_rows, T)
        r1, _ = self._rows(dvar_rows, T)
        r2, _ = self._rows(dtac_rows, T)
        for name, t in (("df_dq", df_dq), ("df_dvar", df_dvar), ("df_dtactile", df_dtactile)):
            if t is not None and (t.device != dev or t.dtype != torch.float64 or not t.is_contiguous()):
                raise _lib.TactileSimError(f"{name}: expected a contiguous float64 tensor on {dev}")
        if carry is None:
            carry = torch.zeros((B, 2, n), dtype=torch.float64, device=dev)
        self._chk(carry, (B, 2, n), "carry")
        df_du = torch.zeros((T, B, nu), dtype=torch.float64, device=dev)
        dq0 = torch.zeros((B, n), dtype=torch.float64, device=dev) if want_q0 else None
        dqd0 = torch.zeros((B, n), dtype=torch.float64, device=dev) if want_q0 else None
        with torch.cuda.device(dev):
            _lib.check(self.lib.tsim_backward(self.handle, B, T, _ptr(fwd["q_traj"]), _ptr(fwd["qd_traj"]), _ptr(u),
                                              ustride, _ptr(fwd["tape"]), _ptr(df_dq), _ptr(r0), _ptr(df_dvar),
                                              _ptr(r1), _ptr(df_dtactile), _ptr(r2), _ptr(carry), _ptr(df_du),
                                              _ptr(dq0), _ptr(dqd0), self._stream()), self.lib)
        return dict(df_du=df_du, carry=carry, df_dq0=dq0, df_dqdot0=dqd0, _keep=(r0, r1, r2))
