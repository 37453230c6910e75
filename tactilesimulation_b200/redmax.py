"""Host-side mirror of the reference's ``redmax_py.Simulation`` for the accelerated path.

Same method names, argument meaning and error behaviour as the pybind11 class
(``DH/python_interface.cpp:33-249``, DH = externals/DiffHand/core/projects/redmax) so that
``R/envs/redmax_torch_functions.py`` / ``redmax_torch_env.py`` / ``tactile_push_env.py`` can be
pointed at it, but every call runs on the GPU through the C ABI and the
object may hold a whole batch of environments:

* numpy face (batch == 1): ``get_q() -> float64[n]`` etc., exactly the reference shapes;
* tensor face (any batch): the ``*_t`` methods take/return CUDA tensors with a leading env dim and
  never leave the device; ``tactilesimulation_b200.torch_functions`` builds the batched
  ``StepSimFunction`` / ``EpisodicSimFunction`` on them.

State machine as in the reference: ``reset()`` must precede ``forward()``
(``DH/Simulation.cpp:1061-1064``), ``forward`` precedes ``backward``/``backward_steps``
(``:1570-1573``, ``:1877-1880``), ``backward_steps`` forbids ``flag_q0``/``flag_qdot0``
(``:1884-1889``); size mismatches raise (``:1578-1606``).  Newton non-convergence is not an error.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from ._lib import TactileSimError
from .scene import Scene, compile_scene
from .sim import BatchedSim


class Options:
    """``Simulation::Options`` (read-only view, python_interface.cpp:36-44)."""

    def __init__(self, gravity, h, integrator):
        self.gravity = np.array(gravity, dtype=np.float64)
        self.h = float(h)
        self.integrator = integrator


class ViewerOptions:
    """Accepted and ignored: there is no display on the GPU box (python_interface.cpp:47-61)."""

    def __init__(self):
        self.fps = 30
        self.speed = 1.0
        self.camera_pos = np.zeros(3)
        self.camera_lookat = np.zeros(3)
        self.camera_up = np.array([0.0, 0.0, 1.0])
        self.ground = True
        self.E_g = np.eye(4)
        self.record = False
        self.record_folder = ""
        self.loop = False
        self.infinite = False


class BackwardInfo:
    """Cotangent inputs of the adjoint (``DH/BackwardData.h:32-46``, python_interface.cpp:64-76)."""

    def __init__(self):
        self.df_dq0 = np.zeros(0)
        self.df_dqdot0 = np.zeros(0)
        self.df_dp = np.zeros(0)
        self.df_du = np.zeros(0)
        self.df_dq = np.zeros(0)
        self.df_dvar = np.zeros(0)
        self.df_dtactile = np.zeros(0)
        self.flag_q0 = self.flag_qdot0 = self.flag_p = self.flag_u = False

    def set_flags(self, flag_q0=False, flag_qdot0=False, flag_p=False, flag_u=False):
        self.flag_q0, self.flag_qdot0, self.flag_p, self.flag_u = bool(flag_q0), bool(flag_qdot0), bool(flag_p), bool(flag_u)


class BackwardResults:
    def __init__(self):
        self.df_dq0 = np.zeros(0)
        self.df_dqdot0 = np.zeros(0)
        self.df_dp = np.zeros(0)
        self.df_du = np.zeros(0)


class _Chunk:
    """One forward() call: T steps recorded on the device."""
    __slots__ = ("T", "u", "fwd", "rows", "start")

    def __init__(self, T, u, fwd, rows, start):
        self.T, self.u, self.fwd, self.rows, self.start = T, u, fwd, rows, start


class Simulation:
    def __init__(self, xml_file_path, verbose: bool = False, batch: int = 1, device="cuda:0", lanes: Optional[int] = None):
        if isinstance(xml_file_path, Scene):
            self.scene = xml_file_path
        else:
            self.scene = compile_scene(str(xml_file_path))
        self._lanes = lanes
        self.core = BatchedSim(self.scene, device=device, lanes=lanes)
        self.device = self.core.device
        self.batch = int(batch)
        c = self.core
        self.ndof_r, self.ndof_m, self.ndof_u = c.ndof_r, c.ndof_m, c.ndof_u
        self.ndof_var, self.ndof_tactile, self.ndof_p = c.ndof_var, c.ndof_tactile, 0
        self.options = Options(self.scene.gravity, self.scene.h, self.scene.integrator)
        self.viewer_options = ViewerOptions()
        self.backward_info = BackwardInfo()
        self.backward_results = BackwardResults()
        z = lambda: torch.zeros((self.batch, self.ndof_r), dtype=torch.float64, device=self.device)
        self._q_init, self._qd_init = z(), z()
        self._q, self._qd = z(), z()
        self._q_prev, self._qd_prev = z(), z()       # state one step back (BDF2: Simulation::_q_his of the reference)
        self._u = torch.zeros((self.batch, self.ndof_u), dtype=torch.float64, device=self.device)
        self._reset_done = False
        self._grad = False
        self._chunks: List[_Chunk] = []
        self._nsteps = 0
        self._current_backward_step = 0
        self._carry: Optional[torch.Tensor] = None
        self._cache = []
        self._env_updates = []                         # per-environment parameter updates: (fn, names, values [B, ...])
        self._dirty = False                            # host scene edited since the device handle was built
        self._q_hist: List[torch.Tensor] = []          # q trajectories since reset() (export_replay; Simulation::_q_his)
        self._virtual = {}                             # poses of the render-only objects

    # ------------------------------------------------------------------ helpers
    def _t(self, x, width, name):
        """numpy/tensor -> [batch,width] fp64 device tensor; size mismatch raises like the reference."""
        if isinstance(x, torch.Tensor):
            t = x.detach().to(device=self.device, dtype=torch.float64)
        else:
            t = torch.as_tensor(np.asarray(x, dtype=np.float64), device=self.device)
        if t.dim() == 1:
            if t.numel() != width:
                raise TactileSimError(f"[Error] {name}: size {t.numel()} != {width}.")
            t = t.unsqueeze(0).expand(self.batch, width)
        if tuple(t.shape) != (self.batch, width):
            raise TactileSimError(f"[Error] {name}: shape {tuple(t.shape)} != {(self.batch, width)}.")
        return t.contiguous().clone()

    def _np(self, t):
        if self.batch != 1:
            raise TactileSimError("the numpy face needs batch == 1; use the *_t methods for batches")
        return t[0].detach().cpu().numpy().copy()

    # ------------------------------------------------------------------ state (python_interface.cpp:102-135)
    def set_state_init(self, q_init, qdot_init):
        self.set_q_init(q_init)
        self.set_qdot_init(qdot_init)

    def set_q_init(self, q_init):
        self._q_init = self._t(q_init, self.ndof_r, "set_q_init")

    def set_qdot_init(self, qdot_init):
        self._qd_init = self._t(qdot_init, self.ndof_r, "set_qdot_init")

    def get_q_init(self):
        return self._np(self._q_init)

    def get_qdot_init(self):
        return self._np(self._qd_init)

    def get_q(self):
        return self._np(self._q)

    def get_qdot(self):
        return self._np(self._qd)

    def get_q_t(self):
        return self._q.clone()

    def get_qdot_t(self):
        return self._qd.clone()

    def set_u(self, u):
        self._u = self._t(u, self.ndof_u, "set_u")

    def get_variables(self):
        return self._np(self.get_variables_t())

    def get_variables_t(self):
        self._ensure_core()
        return self.core.readout(self._q, self._qd)["var"]

    def get_tactile_force_vector(self):
        return self._np(self.get_tactile_force_vector_t())

    def get_tactile_force_vector_t(self):
        self._ensure_core()
        return self.core.readout(self._q, self._qd)["tactile"]

    def get_tactile_sensor_pos(self, name):
        for s in self.scene.sensors:
            if s.name == name:
                return [p.copy() for p in s.pos]
        raise TactileSimError(f"tactile sensor {name} not found")

    def get_tactile_image_pos(self, name):
        for s in self.scene.sensors:
            if s.name == name:
                return [p.copy() for p in s.image_pos]
        raise TactileSimError(f"tactile sensor {name} not found")

    def get_tactile_force(self, name):
        off = 0
        vec = self.get_tactile_force_vector()
        for s in self.scene.sensors:
            M = len(s.pos)
            if s.name == name:
                return [vec[off + 3 * i: off + 3 * i + 3].copy() for i in range(M)]
            off += 3 * M
        raise TactileSimError(f"tactile sensor {name} not found")

    def _sensor(self, name):
        for i, s in enumerate(self.scene.sensors):
            if s.name == name:
                return i, s
        raise TactileSimError(f"tactile sensor {name} not found")

    def _sensor_forces(self, name):
        i, s = self._sensor(name)
        off = 3 * sum(len(t.pos) for t in self.scene.sensors[:i])
        return self.get_tactile_force_vector()[off:off + 3 * len(s.pos)].reshape(-1, 3)

    def get_tactile_normal_force(self, name):
        """python_interface.cpp:150-152: the normal component per marker."""
        return [float(f[2]) for f in self._sensor_forces(name)]

    def get_tactile_shear_force(self, name):
        """python_interface.cpp:153-155: (shear . axis0, shear . axis1) per marker."""
        return [f[:2].copy() for f in self._sensor_forces(name)]

    def get_tactile_flow_images(self):
        """DH/Robot.cpp:372-387: per sensor, the force vectors scattered to their image positions
        (rows x cols x 3, zero where the sensor has no marker)."""
        vec = self.get_tactile_force_vector()
        out, off = [], 0
        for s in self.scene.sensors:
            M = len(s.pos)
            ip = np.asarray(s.image_pos, dtype=np.int64).reshape(M, 2)
            img = np.zeros((int(ip[:, 0].max()) + 1, int(ip[:, 1].max()) + 1, 3))
            img[ip[:, 0], ip[:, 1]] = vec[off:off + 3 * M].reshape(M, 3)
            out.append(img)
            off += 3 * M
        return out

    # ------------------------------------------------------------------ parameter updates (domain randomisation)
    # python_interface.cpp:181-211; DH/Robot.cpp update_* walk the pointer graph of ONE environment.  Here a value may be
    # a scalar / vector (all environments, as in the reference) or carry a leading batch dimension [B, ...]: then every
    # environment gets its own value -- the reference's per-environment randomisation (dclaw_rotate_env.py:173-178,
    # stable_grasp_env.py:122-128, tactile_insertion_env.py:254-275) in one batch.  Updates edit the host scene(s); the
    # device handle is rebuilt lazily, once, at the next reset() / forward() / read-out (options are carried over).
    def _per_env(self, value, tail_ndim):
        a = value.detach().cpu().numpy() if isinstance(value, torch.Tensor) else np.asarray(value, dtype=np.float64)
        return self.batch > 1 and a.ndim == tail_ndim + 1 and a.shape[0] == self.batch, a

    def _update(self, fn, names, values, tail_ndims):
        """fn(scene, *names, *values) edits a Scene; values with a leading [B] dimension are logged per environment."""
        conv = [self._per_env(v, nd) for v, nd in zip(values, tail_ndims)]
        if any(pe for pe, _ in conv):
            self._env_updates.append((fn, names, [a if pe else np.broadcast_to(a, (self.batch,) + a.shape) for pe, a in conv]))
        else:
            fn(self.scene, *names, *[a for _, a in conv])
            # a broadcast update after per-environment ones applies to every environment as well
            if self._env_updates:
                self._env_updates.append((fn, names, [np.broadcast_to(a, (self.batch,) + a.shape) for _, a in conv]))
        self._dirty = True

    def clear_env_parameters(self):
        """Back to one parameter set for all environments (drops the per-environment updates)."""
        self._env_updates = []
        self._dirty = True

    def _rebuild(self):
        import copy
        old = getattr(self, "core", None)
        self.core = BatchedSim(self.scene, device=self.device, lanes=self._lanes)
        if old is not None:
            for k, v in getattr(old, "options", {}).items():
                self.core.set_option(k, v)
        if self._env_updates:
            ibs, dbs = [], []
            for e in range(self.batch):
                sc = copy.deepcopy(self.scene)
                for fn, names, vals in self._env_updates:
                    fn(sc, *names, *[v[e] for v in vals])
                ib, db = sc.pack()
                ibs.append(ib)
                dbs.append(db)
            self.core.set_env_scenes(np.stack(ibs), np.stack(dbs))
        self._dirty = False

    def _ensure_core(self):
        if self._dirty:
            self._rebuild()

    @staticmethod
    def _set_contact(sc, body1, body2, kn, kt, mu, damping):
        names = sc.body_names
        for f in sc.gp_contacts:
            if names[f["body1"]] == body1 and names[f["body2"]] == body2:
                f.update(kn=float(kn), kt=float(kt), mu=float(mu), damping=float(damping))

    @staticmethod
    def _set_tactile(sc, name, kn, kt, mu, damping):
        for s_ in sc.sensors:
            if s_.name == name:
                s_.kn, s_.kt, s_.mu, s_.damping = float(kn), float(kt), float(mu), float(damping)

    @staticmethod
    def _set_damping(sc, joint_name, damping):
        if joint_name in sc.joint_names:
            sc.damping[sc.joint_names.index(joint_name)] = float(damping)

    @staticmethod
    def _set_ee(sc, name, position):
        for e in sc.end_effectors:
            if e["name"] == name:
                e["pos"] = np.asarray(position, dtype=np.float64).copy()

    def update_contact_parameters(self, body1, body2, kn, kt, mu, damping):
        names = self.scene.body_names
        if not any(names[f["body1"]] == body1 and names[f["body2"]] == body2 for f in self.scene.gp_contacts):
            raise TactileSimError(f"contact {body1} - {body2} not found")
        self._update(self._set_contact, (body1, body2), (kn, kt, mu, damping), (0, 0, 0, 0))

    def update_tactile_parameters(self, name, kn, kt, mu, damping):
        self._sensor(name)
        self._update(self._set_tactile, (name,), (kn, kt, mu, damping), (0, 0, 0, 0))

    def update_joint_damping(self, joint_name, damping):
        if joint_name not in self.scene.joint_names:
            raise TactileSimError(f"joint {joint_name} not found")
        self._update(self._set_damping, (joint_name,), (damping,), (0,))

    def update_endeffector_position(self, endeffector_name, position):
        if not any(e["name"] == endeffector_name for e in self.scene.end_effectors):
            raise TactileSimError(f"endeffector {endeffector_name} not found")
        self._update(self._set_ee, (endeffector_name,), (position,), (1,))

    def _scene_update(self, fn, name, value, tail_ndim):
        from . import scene as _scene
        try:
            self._update(fn, (name,), (value,), (tail_ndim,))
        except _scene.SceneError as e:
            raise TactileSimError(str(e))

    def update_body_density(self, body_name, density):
        """DH/python_interface.cpp:181-211, DH/Robot.cpp:596-610 (R/envs/stable_grasp_env.py:122)."""
        from . import scene as _scene
        self._scene_update(_scene.update_body_density, body_name, density, 0)

    def update_body_size(self, body_name, body_size):
        """DH/Robot.cpp:612-626 (R/envs/dclaw_rotate_env.py:175: the cap's (length, radius))."""
        from . import scene as _scene
        self._scene_update(_scene.update_body_size, body_name, body_size, 1)

    def update_joint_location(self, joint_name, joint_location):
        """DH/Robot.cpp:636-650, DH/Joint/Joint.cpp:98-117 (R/envs/dclaw_rotate_env.py:178)."""
        from . import scene as _scene
        self._scene_update(_scene.update_joint_location, joint_name, joint_location, 1)

    def update_body_color(self, body_name, color):
        """Render-only."""
        return None

    def update_virtual_object(self, name, data):
        """Render-only objects (goal marker) do not enter the dynamics; the pose (pos, quat wxyz) is kept for
        export_replay (DH/VirtualObject/VirtualObjectCuboid.cpp:13-30)."""
        d = np.asarray(data, dtype=np.float64).reshape(-1)
        if d.size >= 7:
            self._virtual[name] = d[:7].copy()
        elif d.size >= 3:
            self._virtual[name] = np.concatenate([d[:3], [1.0, 0.0, 0.0, 0.0]])
        return None

    # ------------------------------------------------------------------ reset / caches (Simulation.cpp:999-1055)
    def reset(self, backward_flag: bool = False, backward_design_params_flag: bool = False):
        if backward_design_params_flag:
            raise TactileSimError("design-parameter gradients are out of scope of the B200 path")
        self._ensure_core()
        if backward_flag and self.core.integrator != 0:
            raise TactileSimError("gradients exist for integrator BDF1 only on the B200 path (options.integrator is "
                                  + str(self.options.integrator) + ")")
        self._q = self._q_init.clone()
        self._qd = self._qd_init.clone()
        self._grad = bool(backward_flag)
        self._chunks = []
        self._nsteps = 0
        self._current_backward_step = 0
        self._carry = None
        self._q_hist = []
        self._reset_done = True

    def clearBackwardCache(self):
        self._cache = []

    def saveBackwardCache(self):
        self._cache.append((list(self._chunks), self._nsteps, self._current_backward_step,
                            None if self._carry is None else self._carry.clone(), self.backward_info, self.backward_results))
        self.backward_info = BackwardInfo()
        self.backward_results = BackwardResults()

    def popBackwardCache(self):
        self._chunks, self._nsteps, self._current_backward_step, self._carry, self.backward_info, self.backward_results = self._cache.pop()

    def backwardCacheSize(self):
        return len(self._cache)

    # ------------------------------------------------------------------ forward (Simulation.cpp:1057-1148)
    def forward(self, num_steps, verbose=False, test_derivatives=False, save_last_frame_var_only=False):
        self.forward_t(int(num_steps), self._u, save_last_frame_var_only)

    def forward_t(self, num_steps: int, u: torch.Tensor, save_last_frame_var_only: bool = False, tac_rows=None,
                  want_outputs: bool = False):
        """u: [batch,nu] held for all steps, or [T,batch,nu].  Returns the kernel outputs when
        ``want_outputs`` (q_traj, var, tactile ...)."""
        if not self._reset_done:
            raise TactileSimError("[Error] Please call simulation.reset() before simulation.forward().")
        T = int(num_steps)
        if T <= 0:
            return None
        self._ensure_core()
        rows = ([-1] * (T - 1) + [0]) if save_last_frame_var_only else None
        trows = rows if tac_rows is None else tac_rows
        u = u.contiguous()
        fwd = self.core.forward(self._q, self._qd, u, T, grad=self._grad, var_rows=rows, tac_rows=trows,
                                want_var=want_outputs, want_tactile=want_outputs,
                                want_traj=want_outputs or self._grad or self.batch == 1,
                                q_prev=self._q_prev, qd_prev=self._qd_prev, steps_done=self._nsteps)
        if self._grad:
            self._chunks.append(_Chunk(T, u, fwd, rows, self._nsteps))
            self._current_backward_step += T
        if fwd.get("q_traj") is not None and (self.batch == 1 or self._grad):
            self._q_hist.append(fwd["q_traj"])          # the compat face keeps the history like Simulation::_q_his
        self._nsteps += T
        return fwd if want_outputs else None

    # ------------------------------------------------------------------ adjoint
    def _sweep(self, lo: int, hi: int, df_dq, df_dvar, df_dtac, want_q0: bool):
        """Reverse sweep over global steps [lo, hi); cotangent tensors are [hi-lo, batch, *] or None.
        Returns df_du [hi-lo,batch,nu] and, if want_q0, (df_dq0, df_dqdot0) of step lo."""
        nu = self.ndof_u
        df_du = torch.zeros((hi - lo, self.batch, nu), dtype=torch.float64, device=self.device)
        if self._carry is None:
            self._carry = torch.zeros((self.batch, 2, self.ndof_r), dtype=torch.float64, device=self.device)
        dq0 = dqd0 = None
        for ch in reversed(self._chunks):
            a, b = max(lo, ch.start), min(hi, ch.start + ch.T)
            if a >= b:
                continue
            s, e = a - ch.start, b - ch.start          # slice inside the chunk
            f = ch.fwd
            sub = dict(q_traj=f["q_traj"][s:e], qd_traj=f["qd_traj"][s:e], tape=f["tape"][s:e])
            u = ch.u if ch.u.dim() == 2 else ch.u[s:e]
            T = e - s
            # cotangent rows: the reference zeroes dvar_dq / dtactile on non-final sub-steps of a
            # save_last_frame_var_only chunk (Simulation.cpp:1135-1139)
            vrows = None
            if ch.rows is not None:
                vrows = [(-1 if ch.rows[s + i] < 0 else i) for i in range(T)]
            sl = slice(a - lo, b - lo)
            res = self.core.backward(sub, u, T,
                                     None if df_dq is None else df_dq[sl].contiguous(),
                                     None if df_dvar is None else df_dvar[sl].contiguous(),
                                     None if df_dtac is None else df_dtac[sl].contiguous(),
                                     dq_rows=None, dvar_rows=vrows, dtac_rows=vrows, carry=self._carry,
                                     want_q0=want_q0 and a == lo)
            df_du[sl] = res["df_du"]
            if want_q0 and a == lo:
                dq0, dqd0 = res["df_dq0"], res["df_dqdot0"]
        return df_du, dq0, dqd0

    def backward_t(self, df_dq, df_dvar, df_dtactile, want_q0=True):
        """Full-trajectory adjoint on device tensors [T,batch,*] (Simulation::backward)."""
        if self._nsteps < 1 or not self._chunks:
            raise TactileSimError("[Error] Please call simulation.forward() before simulation.backward().")
        T = self._nsteps
        self._carry = None
        out = self._sweep(0, T, df_dq, df_dvar, df_dtactile, want_q0)
        self._carry = None
        return out

    def backward_last_frame_t(self, num_backward_steps: int, df_dq, df_dvar, df_dtactile):
        """``backward_steps`` for the StepSimFunction shape (R/envs/redmax_torch_functions.py:147-174): cotangents [batch, *]
        on the LAST of the ``num_backward_steps`` sub-steps only.  The reference pads them with zeros to
        ``num_steps`` frames; here the kernel's row maps say which step has a cotangent, so no [T, B, 3M] zero tensor is
        built per gym step.  Same result as ``backward_steps_t`` with the padded tensors."""
        ns = int(num_backward_steps)
        hi = self._current_backward_step
        lo = hi - ns
        ch = self._chunks[-1] if self._chunks else None
        for c in reversed(self._chunks):
            if c.start <= lo and hi <= c.start + c.T:
                ch = c
                break
        if hi <= 0 or lo < 0 or ch is None or not (ch.start == lo and ch.start + ch.T == hi):
            # general case (the sweep crosses forward() calls): pad like the reference and take the general path
            def pad(c, width):
                if c is None or width == 0:
                    return None
                full = torch.zeros((ns, self.batch, width), dtype=torch.float64, device=self.device)
                full[-1] = c.reshape(self.batch, width)
                return full
            return self.backward_steps_t(ns, pad(df_dq, self.ndof_r), pad(df_dvar, self.ndof_var), pad(df_dtactile, self.ndof_tactile))
        if self._carry is None:
            self._carry = torch.zeros((self.batch, 2, self.ndof_r), dtype=torch.float64, device=self.device)
        last = [-1] * (ns - 1) + [0]
        one = lambda c, width: None if (c is None or width == 0) else c.reshape(1, self.batch, width).contiguous()
        f = ch.fwd
        res = self.core.backward(dict(q_traj=f["q_traj"], qd_traj=f["qd_traj"], tape=f["tape"]), ch.u, ns,
                                 one(df_dq, self.ndof_r), one(df_dvar, self.ndof_var), one(df_dtactile, self.ndof_tactile),
                                 dq_rows=last, dvar_rows=last, dtac_rows=last, carry=self._carry, want_q0=False)
        self._current_backward_step = lo
        return res["df_du"]

    def backward_steps_t(self, num_backward_steps: int, df_dq, df_dvar, df_dtactile):
        """The last ``num_backward_steps`` not yet swept steps (Simulation::backward_steps)."""
        if self._current_backward_step <= 0:
            raise TactileSimError("[Error] Please call simulation.forward() before simulation.backward().")
        hi = self._current_backward_step
        lo = hi - int(num_backward_steps)
        if lo < 0:
            raise TactileSimError("backward_steps: more steps than recorded")
        df_du, _, _ = self._sweep(lo, hi, df_dq, df_dvar, df_dtactile, False)
        self._current_backward_step = lo
        return df_du

    def _cot(self, flat, T, width, name):
        a = np.asarray(flat, dtype=np.float64).reshape(-1)
        if a.size != width * T:
            raise TactileSimError(f"_backward_info.{name}.size != {width} * T")
        if width == 0:
            return None
        return torch.as_tensor(a.reshape(T, 1, width), device=self.device)

    def backward(self):
        if self._nsteps < 1 or not self._grad:
            raise TactileSimError("[Error] Please call simulation.forward() before simulation.backward().")
        bi, T = self.backward_info, self._nsteps
        n, nu = self.ndof_r, self.ndof_u
        if bi.flag_q0 and np.asarray(bi.df_dq0).size != n:
            raise TactileSimError("_backward_info._df_dq0.size != _ndof_r")
        if bi.flag_qdot0 and np.asarray(bi.df_dqdot0).size != n:
            raise TactileSimError("_backward_info._df_dqdot0.size != _ndof_r")
        if bi.flag_p:
            raise TactileSimError("design-parameter gradients are out of scope of the B200 path")
        if bi.flag_u and np.asarray(bi.df_du).size != nu * T:
            raise TactileSimError("_backward_info._df_du.size != _ndof_u * T")
        dq = self._cot(bi.df_dq, T, n, "_df_dq")
        dv = self._cot(bi.df_dvar, T, self.ndof_var, "_df_dvar")
        dt = self._cot(bi.df_dtactile, T, self.ndof_tactile, "_df_dtactile")
        df_du, dq0, dqd0 = self.backward_t(dq, dv, dt, want_q0=True)
        br = self.backward_results
        if bi.flag_q0:
            br.df_dq0 = np.asarray(bi.df_dq0, dtype=np.float64) + self._np(dq0)
        if bi.flag_qdot0:
            br.df_dqdot0 = np.asarray(bi.df_dqdot0, dtype=np.float64) + self._np(dqd0)
        if bi.flag_u:
            br.df_du = np.asarray(bi.df_du, dtype=np.float64).reshape(-1) + df_du[:, 0].reshape(-1).cpu().numpy()

    def backward_steps(self, num_backward_steps):
        bi, ns = self.backward_info, int(num_backward_steps)
        if self._current_backward_step <= 0:
            raise TactileSimError("[Error] Please call simulation.forward() before simulation.backward().")
        if bi.flag_q0:
            raise TactileSimError("_backward_info._flag_q0 should be false for backward_steps")
        if bi.flag_qdot0:
            raise TactileSimError("_backward_info._flag_qdot0 should be false for backward_steps")
        if bi.flag_p:
            raise TactileSimError("design-parameter gradients are out of scope of the B200 path")
        n, nu = self.ndof_r, self.ndof_u
        if bi.flag_u and np.asarray(bi.df_du).size != nu * ns:
            raise TactileSimError("_backward_info._df_du.size != _ndof_u * num_backward_steps")
        dq = self._cot(bi.df_dq, ns, n, "_df_dq")
        dv = self._cot(bi.df_dvar, ns, self.ndof_var, "_df_dvar")
        dt = self._cot(bi.df_dtactile, ns, self.ndof_tactile, "_df_dtactile")
        df_du = self.backward_steps_t(ns, dq, dv, dt)
        if bi.flag_u:
            self.backward_results.df_du = (np.asarray(bi.df_du, dtype=np.float64).reshape(-1)
                                           + df_du[:, 0].reshape(-1).cpu().numpy())

    # ------------------------------------------------------------------ viewer / reports: no display here
    def replay(self):
        return None

    def export_replay(self, folder):
        """``Simulation::export_replay`` (DH/Simulation.cpp:2037-2120): <folder>/meshes/<k>.obj and one <folder>/<i>.txt per
        state of the q history since reset() (frame 0 = the initial state) holding the 4 x 4 world transforms of the
        bodies, render-only objects, sensors and end-effectors -- the reference's text format, written by
        ``tactilesimulation_b200.replay`` (host-side kinematics).  Environment 0 of a batch is exported."""
        from . import replay
        qs = [self._q_init[0:1].detach().cpu().numpy()] if self._reset_done else []
        qs += [h[:, 0].detach().cpu().numpy() for h in self._q_hist]
        qh = np.concatenate(qs, axis=0) if qs else np.zeros((0, self.ndof_r))
        vp = [self._virtual.get(nm, self.scene.virtual_pose[i] if i < len(self.scene.virtual_pose) else np.array([0, 0, 0, 1.0, 0, 0, 0]))
              for i, nm in enumerate(self.scene.virtual_names)]
        return replay.export_replay(self.scene, qh, str(folder), virtual_pose=vp)

    def print_time_report(self):
        print("[tactilesimulation_b200] timing lives in CUDA events / ncu; see bench.py")

    def print_ctrl_info(self):
        for a in self.scene.actuators:
            print("motor on joint", self.scene.joint_names[a["joint"]], "ctrl_range", a["cmin"], a["cmax"])


def make_sim(env_name, integrator="BDF2"):
    raise TactileSimError("programmatic test scenes (SimEnvGenerator) are out of scope of the B200 path")
