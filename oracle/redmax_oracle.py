"""CPU oracle: a numpy restatement of the reference DiffRedMax hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``tactilesimulation_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may.  It is the
checker, never the thing measured or shipped.

Parity status: PINNED.  ``tests/test_oracle_vs_golden.py`` checks this restatement against
golden vectors produced by the unmodified reference C++ (built headless by
``oracle/build_ref.sh`` into ``oracle/_ref``; generator ``tests/golden/make_golden.py``), and,
when ``oracle/_ref`` is importable, against the reference run live.

The restatement deliberately keeps the reference's *dense maximal-coordinate* structure
(J is ndof_m x ndof_r, Km/Dm are ndof_m x ndof_m, dJ_dq is a rank-3 tensor), one scene per
object, fp64 -- unlike the CUDA path, which never materialises those objects.  So a match
between the two is a match between two independent derivations.

Path shorthand: DH = /root/reference/externals/DiffHand/core/projects/redmax
"""
from __future__ import annotations

import math
import numpy as np

JT_FIXED, JT_REVOLUTE, JT_PRISMATIC, JT_PLANAR, JT_TRANSLATIONAL = 0, 1, 2, 3, 4
SH_NONE, SH_CUBOID, SH_CYLINDER, SH_SPHERE = 0, 1, 2, 3
ACT_FORCE, ACT_POS = 0, 1
EPS = 1e-8  # constants::eps, DH/Common.h


# ------------------------------------------------------------------ math helpers (DH/Utils.h:25-140)
def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def SE(R, p):
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = p
    return E


def Einv(E):
    Rt = E[:3, :3].T
    return SE(Rt, -Rt @ E[:3, 3])


def Ad(E):
    R = E[:3, :3]
    A = np.zeros((6, 6))
    A[:3, :3] = R
    A[3:, :3] = skew(E[:3, 3]) @ R
    A[3:, 3:] = R
    return A


def ad(phi):
    a = np.zeros((6, 6))
    a[:3, :3] = skew(phi[:3])
    a[3:, :3] = skew(phi[3:])
    a[3:, 3:] = skew(phi[:3])
    return a


def gamma(x):
    G = np.zeros((3, 6))
    G[:, :3] = skew(x).T
    G[:, 3:] = np.eye(3)
    return G


def angle_axis(angle, axis):
    """Eigen::AngleAxis::toRotationMatrix (Rodrigues)."""
    c, s = math.cos(angle), math.sin(angle)
    return c * np.eye(3) + s * skew(axis) + (1.0 - c) * np.outer(axis, axis)


E3 = [np.array([1.0, 0, 0]), np.array([0, 1.0, 0]), np.array([0, 0, 1.0])]


class OracleSim:
    """One scene, one environment.  API mirrors ``redmax_py.Simulation`` where it matters."""

    def __init__(self, scene):
        self.sc = sc = scene
        if sc.integrator != "BDF1":
            raise RuntimeError("oracle restates the BDF1 path only")
        self.n = sc.ndof_r
        self.m = sc.ndof_m
        self.nu = sc.ndof_u
        self.nvar = sc.ndof_var
        self.ntac = sc.ndof_tactile
        self.nj = sc.nj
        self.h = sc.h
        self.q = np.zeros(self.n)
        self.qdot = np.zeros(self.n)
        self.u = np.zeros(self.nu)
        self.q_init = np.zeros(self.n)
        self.qdot_init = np.zeros(self.n)
        self.act_q = [np.zeros(a["ndof"]) for a in sc.actuators]
        self.act_qd = [np.zeros(a["ndof"]) for a in sc.actuators]
        self.q_his = []
        self.qdot_his = []
        self.backward_flag = False
        self.newton_iters = []
        self.ls_evals = []
        self.anc = []  # ancestor joint chain (parent first) per joint
        for j in range(self.nj):
            c, p = [], sc.parent[j]
            while p >= 0:
                c.append(p)
                p = sc.parent[p]
            self.anc.append(c)
        self._update()

    # ------------------------------------------------------------------ state plumbing
    def set_state_init(self, q, qd):
        self.q_init = np.array(q, dtype=np.float64)
        self.qdot_init = np.array(qd, dtype=np.float64)

    def set_q_init(self, q):
        self.q_init = np.array(q, dtype=np.float64)

    def get_q_init(self):
        return self.q_init.copy()

    def set_u(self, u):
        u = np.asarray(u, dtype=np.float64)
        if u.size != self.nu:
            raise RuntimeError("[Error] set_u: u.size() != _ndof_u.")
        self.u = u.copy()

    def get_q(self):
        return self.q.copy()

    def get_qdot(self):
        return self.qdot.copy()

    def _set_state(self, q, qdot):
        self.q = np.array(q, dtype=np.float64)
        self.qdot = np.array(qdot, dtype=np.float64)

    def reset(self, backward_flag=False):
        """DH/Simulation.cpp:999-1033"""
        self._set_state(self.q_init, self.qdot_init)
        self._update()
        self._update_actuator_states(self.q_init, self.qdot_init)
        self.q_his = [self.q_init.copy()]
        self.qdot_his = [self.qdot_init.copy()]
        self.backward_flag = backward_flag
        self.tape = dict(H=[], M=[], D=[], dfr_dqprev=[], dfr_dqdotprev=[], dg_du=[], dvar_dq=[],
                         dtactile_dq=[], dtactile_dqdot=[])
        self.z_his = []
        self.df_dtactile_his = []
        self.current_backward_step = 0
        self.newton_iters = []
        self.ls_evals = []

    def _update_actuator_states(self, q, qd):
        sc = self.sc
        for i, a in enumerate(sc.actuators):
            o = sc.qoff[a["joint"]]
            self.act_q[i] = q[o:o + a["ndof"]].copy()
            self.act_qd[i] = qd[o:o + a["ndof"]].copy()

    # ------------------------------------------------------------------ kinematics
    def _joint_local(self, j):
        """Q, A, Adot, S, dA_dq, dAdot_dq, dQ_dq of joint j (DH/Joint/Joint*.cpp update())."""
        sc = self.sc
        jt, nd, o = sc.jtype[j], sc.ndof[j], sc.qoff[j]
        q, qd = self.q[o:o + nd], self.qdot[o:o + nd]
        Q = np.eye(4)
        Adot = np.zeros((6, 6))
        S = np.zeros((6, nd))
        dA = [np.zeros((6, 6)) for _ in range(nd)]
        dAdot = [np.zeros((6, 6)) for _ in range(nd)]
        dQ = [np.zeros((4, 4)) for _ in range(nd)]
        if jt == JT_REVOLUTE:                      # JointRevolute.cpp:38-69
            a = sc.axis0[j]
            R = angle_axis(q[0], a)
            Q = SE(R, np.zeros(3))
            ab = skew(a)
            dR = R @ ab
            dA[0][:3, :3] = dR
            dA[0][3:, 3:] = dR
            Rdot = dR * qd[0]
            Adot[:3, :3] = Rdot
            Adot[3:, 3:] = Rdot
            dRdot = dR @ ab * qd[0]
            dAdot[0][:3, :3] = dRdot
            dAdot[0][3:, 3:] = dRdot
            S[:3, 0] = a
            dQ[0][:3, :3] = dR
        elif jt == JT_PRISMATIC:                   # JointPrismatic.cpp:22-45
            a = sc.axis0[j]
            Q[:3, 3] = a * q[0]
            dA[0][3:, :3] = skew(a)
            Adot[3:, :3] = skew(a) * qd[0]
            S[3:, 0] = a
            dQ[0][:3, 3] = a
        elif jt == JT_PLANAR:                      # JointPlanar.cpp:7-33
            a0, a1 = sc.axis0[j], sc.axis1[j]
            Q[:3, 3] = a0 * q[0] + a1 * q[1]
            dA[0][3:, :3] = skew(a0)
            dA[1][3:, :3] = skew(a1)
            Adot[3:, :3] = skew(a0) * qd[0] + skew(a1) * qd[1]
            S[3:, 0] = a0
            S[3:, 1] = a1
            dQ[0][:3, 3] = a0
            dQ[1][:3, 3] = a1
        elif jt == JT_TRANSLATIONAL:               # JointTranslational.cpp:9-40
            Q[:3, 3] = q
            S[3:, :] = np.eye(3)
            Adot[3:, :3] = skew(qd)
            for i in range(3):
                dA[i][3:, :3] = skew(E3[i])
                dQ[i][i, 3] = 1.0
        A = Ad(Q)
        return Q, A, Adot, S, dA, dAdot, dQ

    def _update(self):
        """Robot::update -> Joint::update -> Body::update (Joint.cpp:119-165, Body.cpp:122-165)."""
        sc = self.sc
        nj = self.nj
        self.J_ = [None] * nj
        for j in range(nj):
            Q, A, Adot, S, dA, dAdot, dQ = self._joint_local(j)
            nd, o, p = sc.ndof[j], sc.qoff[j], sc.parent[j]
            Q_inv = Einv(Q)
            A_inv = Ad(Q_inv)
            E_pj = sc.E_pj0[j] @ Q
            E_jp = Einv(E_pj)
            dEpj = [sc.E_pj0[j] @ dQ[i] for i in range(nd)]
            A_jp = Ad(E_jp)
            E_0j = E_pj if p < 0 else self.J_[p]["E_0j"] @ E_pj
            phi = S @ self.qdot[o:o + nd] if nd > 0 else np.zeros(6)
            if p >= 0:
                phi = phi + A_jp @ self.J_[p]["phi"]
            # body
            E_ji = sc.E_ji[j]
            E_ij = Einv(E_ji)
            A_ij = Ad(E_ij)
            E_0i = E_0j @ E_ji
            E_i0 = Einv(E_0i)
            bphi = A_ij @ phi
            d = dict(Q=Q, S=S, E_pj=E_pj, dEpj=dEpj, E_0j=E_0j, phi=phi, E_0i=E_0i, E_i0=E_i0,
                     bphi=bphi, A_ij=A_ij, E_ji=E_ji)
            if p >= 0:
                E_ip = E_i0 @ self.J_[p]["E_0i"]
                d["A_ip"] = Ad(E_ip)
                Aleft = -Ad(E_ij @ Q_inv)
                E_jp_0 = Einv(sc.E_pj0[j]) @ self.J_[p]["E_ji"]
                Aright = Ad(Q_inv @ E_jp_0)
                d["A_ip_dot"] = Aleft @ Adot @ Aright
                d["dAip_dq"] = [Aleft @ dA[k] @ Aright for k in range(nd)]
                d["dAipdot_dq"] = [Aleft @ (dAdot[k] - dA[k] @ A_inv @ Adot - Adot @ A_inv @ dA[k]) @ Aright
                                   for k in range(nd)]
            self.J_[j] = d

    # ------------------------------------------------------------------ Jacobians (Robot.cpp:803-885)
    def _jacobian(self, deriv):
        sc, n, m = self.sc, self.n, self.m
        J = np.zeros((m, n))
        Jd = np.zeros((m, n))
        dJ = np.zeros((n, m, n)) if deriv else None
        dJd = np.zeros((n, m, n)) if deriv else None
        for j in range(self.nj):
            d = self.J_[j]
            nd, o, p = sc.ndof[j], sc.qoff[j], sc.parent[j]
            r = slice(6 * j, 6 * j + 6)
            if nd > 0:
                J[r, o:o + nd] = d["A_ij"] @ d["S"]
                # S_j_dot and dS*/dq vanish for every joint type on this path
            if p < 0:
                continue
            rp = slice(6 * p, 6 * p + 6)
            for now in self.anc[j]:
                ndn, on = sc.ndof[now], sc.qoff[now]
                if ndn == 0:
                    continue
                c = slice(on, on + ndn)
                J[r, c] = d["A_ip"] @ J[rp, c]
                Jd[r, c] = d["A_ip_dot"] @ J[rp, c] + d["A_ip"] @ Jd[rp, c]
                if deriv:
                    for k in range(nd):
                        dJ[o + k][r, c] = d["dAip_dq"][k] @ J[rp, c]
                        dJd[o + k][r, c] = d["dAipdot_dq"][k] @ J[rp, c] + d["dAip_dq"][k] @ Jd[rp, c]
                    for nex in self.anc[j]:
                        for k in range(sc.ndof[nex]):
                            kk = sc.qoff[nex] + k
                            dJ[kk][r, c] = d["A_ip"] @ dJ[kk][rp, c]
                            dJd[kk][r, c] = d["A_ip_dot"] @ dJ[kk][rp, c] + d["A_ip"] @ dJd[kk][rp, c]
                        if nex == now:
                            break
        return J, Jd, dJ, dJd

    # ------------------------------------------------------------------ cuboid SDF (BodyCuboid.cpp:135-282)
    def _cuboid_distance(self, b, xw):
        E = self.J_[b]["E_0i"]
        x = E[:3, :3].T @ (xw - E[:3, 3])
        s = self.sc.size[b] / 2.0
        d = -99999999.0
        for i in range(3):
            d = max(d, max(x[i] - s[i], -x[i] - s[i]))
        return d

    def _cuboid_collision(self, b, xw, xw_dot, deriv):
        E = self.J_[b]["E_0i"]
        phi = self.J_[b]["bphi"]
        I = np.eye(3)
        R2, p2 = E[:3, :3], E[:3, 3]
        w2, v2 = phi[:3], phi[3:]
        s = self.sc.size[b] / 2.0
        x = R2.T @ (xw - p2)
        d = -9999999.0
        e = np.zeros(3)
        for i in range(3):
            if x[i] - s[i] > d:
                d = x[i] - s[i]
                e = E3[i].copy()
            if -x[i] - s[i] > d:
                d = -x[i] - s[i]
                e = -E3[i]
        n = R2 @ e
        xi2 = x - d * e
        ddot = e @ (skew(w2).T @ x + R2.T @ xw_dot - v2)
        xw2_dot = R2 @ (np.cross(w2, xi2) + v2)
        vw = xw_dot - xw2_dot
        P = I - np.outer(n, n)
        tdot = P @ vw
        out = dict(d=d, n=n, ddot=ddot, tdot=tdot, xi2=xi2)
        if not deriv:
            return out
        dx_dxw = R2.T
        dx_dw2 = skew(x)
        dd_dxw = e @ dx_dxw
        dd_dq2 = np.concatenate([e @ dx_dw2, -e])
        dn_dxw = np.zeros((3, 3))
        dn_dq2 = np.zeros((3, 6))
        dn_dq2[:, :3] = -R2 @ skew(e)
        dxi2_dxw = dx_dxw - np.outer(e, dd_dxw)
        dxi2_dq2 = np.zeros((3, 6))
        dxi2_dq2[:, :3] = dx_dw2 - np.outer(e, dd_dq2[:3])
        dxi2_dq2[:, 3:] = -I + np.outer(e, e)
        dddot_dxw = e @ skew(w2).T @ R2.T
        dddot_dxwdot = e @ R2.T
        dddot_dq2 = np.concatenate([e @ (skew(w2).T @ skew(x) + skew(R2.T @ xw_dot)), -e @ skew(w2).T])
        dddot_dphi2 = np.concatenate([[e @ skew(E3[0]).T @ x, e @ skew(E3[1]).T @ x, e @ skew(E3[2]).T @ x], -e])
        dxw2dot_dxw = R2 @ skew(w2) @ dxi2_dxw
        dxw2dot_dq2 = np.zeros((3, 6))
        dxw2dot_dq2[:, :3] = -R2 @ skew(np.cross(w2, xi2) + v2) + R2 @ skew(w2) @ dxi2_dq2[:, :3]
        dxw2dot_dq2[:, 3:] = R2 @ skew(w2) @ dxi2_dq2[:, 3:]
        dxw2dot_dphi2 = np.zeros((3, 6))
        for i in range(3):
            dxw2dot_dphi2[:, i] = R2 @ skew(E3[i]) @ xi2
        dxw2dot_dphi2[:, 3:] = R2
        dtdot_dxw = -P @ dxw2dot_dxw
        dtdot_dxwdot = P.copy()
        dtdot_dq2 = np.zeros((3, 6))
        dtdot_dq2[:, :3] = (-dxw2dot_dq2[:, :3] - (n @ vw) * dn_dq2[:, :3]
                            - np.outer(n, vw @ dn_dq2[:, :3] - n @ dxw2dot_dq2[:, :3]))
        dtdot_dq2[:, 3:] = P @ (-dxw2dot_dq2[:, 3:])
        dtdot_dphi2 = P @ (-dxw2dot_dphi2)
        out.update(dd_dxw=dd_dxw, dd_dq2=dd_dq2, dn_dxw=dn_dxw, dn_dq2=dn_dq2, dddot_dxw=dddot_dxw,
                   dddot_dxwdot=dddot_dxwdot, dddot_dq2=dddot_dq2, dddot_dphi2=dddot_dphi2,
                   dtdot_dxw=dtdot_dxw, dtdot_dxwdot=dtdot_dxwdot, dtdot_dq2=dtdot_dq2,
                   dtdot_dphi2=dtdot_dphi2, dxi2_dxw=dxi2_dxw, dxi2_dq2=dxi2_dq2)
        return out

    # ------------------------------------------------------------------ contact detection (CollisionDetection.cpp)
    def ground_contact_ids(self, g):
        """collision_detection_ground_object :13-42 -- d <= 0"""
        E = self.J_[g["body"]]["E_0i"]
        xg, ng = self.sc.E_g[:3, 3], self.sc.E_g[:3, 2]
        ids = []
        for i, xi in enumerate(self.sc.contact_points[g["body"]]):
            xw = E[:3, :3] @ xi + E[:3, 3]
            if ng @ (xw - xg) <= 0.0:
                ids.append(i)
        return ids

    def gp_contact_ids(self, f):
        """collision_detection_general_primitive :66-83 -- d < 0"""
        E = self.J_[f["body1"]]["E_0i"]
        ids = []
        for i, xi in enumerate(self.sc.contact_points[f["body1"]]):
            xw = E[:3, :3] @ xi + E[:3, 3]
            if self._cuboid_distance(f["body2"], xw) < 0.0:
                ids.append(i)
        return ids

    def contact_sets(self):
        out = dict(ground=[self.ground_contact_ids(g) for g in self.sc.ground_contacts],
                   gp=[self.gp_contact_ids(f) for f in self.sc.gp_contacts], marker_body=[])
        for s in self.sc.sensors:
            _, body = self._tactile_values(s)
            out["marker_body"].append(body)
        return out

    # ------------------------------------------------------------------ forces (Robot.cpp:660-698)
    def _forces(self, deriv):
        sc, n, m = self.sc, self.n, self.m
        fm, fr = np.zeros(m), np.zeros(n)
        Km = np.zeros((m, m)) if deriv else None
        Dm = np.zeros((m, m)) if deriv else None
        Kr = np.zeros((n, n)) if deriv else None
        Dr = np.zeros((n, n)) if deriv else None
        # bodies: Coriolis + gravity (Body.cpp:234-290)
        for b in range(self.nj):
            d = self.J_[b]
            R = d["E_0i"][:3, :3]
            phi = d["bphi"]
            I6 = sc.inertia[b]
            adT = ad(phi).T
            fcor = adT @ (I6 * phi)
            fg = np.zeros(6)
            fg[3:] = R.T @ (I6[3] * sc.gravity)
            r = slice(6 * b, 6 * b + 6)
            fm[r] += fcor + fg
            if deriv:
                Km[6 * b + 3:6 * b + 6, 6 * b:6 * b + 3] += skew(fg[3:])
                Iw = I6[:3] * phi[:3]
                mv = I6[3:] * phi[3:]
                for k in range(6):
                    Dm[r, 6 * b + k] += adT[:, k] * I6[k]
                for k in range(3):
                    ek = skew(E3[k])
                    Dm[6 * b:6 * b + 3, 6 * b + k] -= ek @ Iw
                    Dm[6 * b:6 * b + 3, 6 * b + 3 + k] -= ek @ mv
                    Dm[6 * b + 3:6 * b + 6, 6 * b + k] -= ek @ mv
        # joints: damping + limits (Joint.cpp:251-289); joint stiffness is never parsed
        for j in range(self.nj):
            nd, o = sc.ndof[j], sc.qoff[j]
            for i in range(nd):
                fr[o + i] += -sc.damping[j] * self.qdot[o + i]
                qi = self.q[o + i]
                if qi < sc.lim_lo[j]:
                    fr[o + i] += sc.lim_k[j] * (sc.lim_lo[j] - qi)
                if qi > sc.lim_hi[j]:
                    fr[o + i] += sc.lim_k[j] * (sc.lim_hi[j] - qi)
                if deriv:
                    Dr[o + i, o + i] -= sc.damping[j]
                    if qi < sc.lim_lo[j] or qi > sc.lim_hi[j]:
                        Kr[o + i, o + i] += -sc.lim_k[j]
        # forces, in XML order: here ground contacts then general-primitive contacts are kept in
        # their own lists; addition order does not change the sums beyond rounding.
        for g in sc.ground_contacts:
            self._ground_force(g, fm, Km, Dm, deriv)
        for f in sc.gp_contacts:
            self._gp_force(f, fm, Km, Dm, deriv)
        # actuators (ActuatorMotor.cpp:31-42)
        for i, a in enumerate(sc.actuators):
            o = sc.qoff[a["joint"]]
            for k in range(a["ndof"]):
                fr[o + k] += self._act_force(i, a, k)
        return fm, fr, Km, Dm, Kr, Dr

    def _act_force(self, i, a, k):
        u = self.u[a["uoff"] + k]
        if a["mode"] == ACT_FORCE:
            uc = max(min(u, 1.0), -1.0)
            f = (uc - (-1.0)) * (1.0 / (1.0 - (-1.0))) * (a["cmax"][k] - a["cmin"][k]) + a["cmin"][k]
            return max(min(f, a["cmax"][k]), a["cmin"][k])
        f = a["P"][k] * (u - self.act_q[i][k]) + a["D"][k] * (-self.act_qd[i][k])
        return max(min(f, a["cmax"][k]), a["cmin"][k])

    def _dfr_du(self):
        """compute_dfdu (ActuatorMotor.cpp:48-62)"""
        sc = self.sc
        out = np.zeros((self.n, self.nu))
        for i, a in enumerate(sc.actuators):
            o = sc.qoff[a["joint"]]
            for k in range(a["ndof"]):
                u = self.u[a["uoff"] + k]
                if a["mode"] == ACT_FORCE:
                    if -1.0 <= u <= 1.0:
                        out[o + k, a["uoff"] + k] += (a["cmax"][k] - a["cmin"][k]) / 2.0
                else:
                    f = a["P"][k] * (u - self.act_q[i][k]) + a["D"][k] * (-self.act_qd[i][k])
                    if a["cmin"][k] <= f <= a["cmax"][k]:
                        out[o + k, a["uoff"] + k] += a["P"][k]
        return out

    def _extra_derivatives(self):
        """compute_extra_derivatives (ActuatorMotor.cpp:64-75); note the column index is the
        actuator's u index, as in the reference."""
        sc = self.sc
        dqp = np.zeros((self.n, self.n))
        dqdp = np.zeros((self.n, self.n))
        for i, a in enumerate(sc.actuators):
            if a["mode"] != ACT_POS:
                continue
            o = sc.qoff[a["joint"]]
            for k in range(a["ndof"]):
                u = self.u[a["uoff"] + k]
                f = a["P"][k] * (u - self.act_q[i][k]) + a["D"][k] * (-self.act_qd[i][k])
                if a["cmin"][k] <= f <= a["cmax"][k]:
                    dqp[o + k, a["uoff"] + k] -= a["P"][k]
                    dqdp[o + k, a["uoff"] + k] -= a["D"][k]
        return dqp, dqdp

    def _ground_force(self, g, fm, Km, Dm, deriv):
        """ForceGroundContact.cpp:105-147 / :149-242"""
        sc = self.sc
        b = g["body"]
        kn, kt, mu, damping = g["kn"], g["kt"], g["mu"], g["damping"]
        xg, ng = sc.E_g[:3, 3], sc.E_g[:3, 2]
        N = np.outer(ng, ng)
        I = np.eye(3)
        Z = np.zeros((3, 3))
        T = I - N
        E = self.J_[b]["E_0i"]
        R, p = E[:3, :3], E[:3, 3]
        phi = self.J_[b]["bphi"]
        RNR = R.T @ N @ R
        pxgtmp = skew(R.T @ N @ (p - xg))
        r = slice(6 * b, 6 * b + 6)
        for i in self.ground_contact_ids(g):
            xi = sc.contact_points[b][i]
            xw = R @ xi + p
            d = ng @ (xw - xg)
            G = gamma(xi)
            Jc = R @ G
            Gphi = G @ phi
            vwi = R @ Gphi
            fc = -kn * ng * d - damping * N @ vwi
            fm[r] += Jc.T @ fc
            if deriv:
                RNRxi = RNR @ xi
                RNRGphi = RNR @ Gphi
                tmp1 = np.stack([-skew(E3[k]) @ RNRxi for k in range(3)], axis=1)
                tmp1 = tmp1 + (-RNR @ skew(xi) + pxgtmp)
                tmp2 = np.stack([-skew(E3[k]) @ RNRGphi for k in range(3)], axis=1)
                tmp2 = tmp2 + (-RNR @ skew(Gphi))
                Km[r, r] -= kn * G.T @ np.hstack([tmp1, RNR]) + damping * G.T @ np.hstack([tmp2, Z])
                Dm[r, r] -= damping * G.T @ RNR @ G
            if mu < EPS:
                continue
            a = T @ vwi
            anorm = np.linalg.norm(a)
            if mu * abs(kn * d) >= kt * anorm - EPS:
                fm[r] += Jc.T @ (-kt * a)
                if deriv:
                    B = R.T @ T @ R
                    tmp3 = np.hstack([np.stack([(B @ skew(E3[k]) - skew(E3[k]) @ B) @ Gphi for k in range(3)], axis=1), Z])
                    Dm[r, r] += -kt * Jc.T @ T @ Jc
                    Km[r, r] += -kt * G.T @ tmp3
            else:
                mukn = mu * kn
                t = a / anorm
                fm[r] += Jc.T @ (mukn * d * t)
                if deriv:
                    A = (I - np.outer(t, t)) / anorm
                    Rt = R.T @ t
                    K1 = np.hstack([np.stack([skew(E3[k]) @ Rt for k in range(3)], axis=1), Z]) * (-d)
                    K2 = np.outer(R.T @ t, ng) @ Jc
                    K3 = -d * R.T @ A @ T @ R @ np.hstack([skew(Gphi), Z])
                    Dm[r, r] += mukn * Jc.T @ (d * A) @ T @ Jc
                    Km[r, r] += mukn * G.T @ (K1 + K2 + K3)

    def _gp_force(self, f, fm, Km, Dm, deriv):
        """ForceGeneralPrimitiveContact.cpp:154-229 / :231-456"""
        sc = self.sc
        b1, b2 = f["body1"], f["body2"]
        kn, kt, mu, damping = f["kn"], f["kt"], f["mu"], f["damping"]   # _scale = _contact_scale = 1
        E1, E2 = self.J_[b1]["E_0i"], self.J_[b2]["E_0i"]
        R1, p1, R2 = E1[:3, :3], E1[:3, 3], E2[:3, :3]
        phi1 = self.J_[b1]["bphi"]
        fsub = np.zeros(12)
        K = np.zeros((12, 12))
        D = np.zeros((12, 12))
        for i in self.gp_contact_ids(f):
            xi1 = sc.contact_points[b1][i]
            xw1 = R1 @ xi1 + p1
            v1 = skew(xi1).T @ phi1[:3] + phi1[3:]
            xw1_dot = R1 @ v1
            c = self._cuboid_collision(b2, xw1, xw1_dot, deriv)
            d, n, ddot, tdot, xi2 = c["d"], c["n"], c["ddot"], c["tdot"], c["xi2"]
            G1, G2 = gamma(xi1), gamma(xi2)
            GTRT1 = G1.T @ R1.T
            GTRT2 = G2.T @ R2.T
            ni1, ni2 = GTRT1 @ n, GTRT2 @ n
            s = kn * d - damping * ddot * d
            fc1 = -s * ni1
            fsub[:6] += fc1
            fsub[6:] += s * ni2
            fc_norm = np.linalg.norm(fc1)
            tdot_norm = np.linalg.norm(tdot)
            static = mu * fc_norm >= kt * tdot_norm - EPS
            if deriv:
                dxw1_dq1 = np.hstack([-R1 @ skew(xi1), R1])
                dxw1dot_dq1 = np.hstack([-R1 @ skew(v1), np.zeros((3, 3))])
                dxw1dot_dphi1 = np.hstack([R1 @ skew(xi1).T, R1])
                dxi2_dq1 = c["dxi2_dxw"] @ dxw1_dq1
                dG = np.zeros((3, 3, 6))  # dGamma2_dxi2(k) : (3,6)
                dG[0][1, 2] = 1
                dG[0][2, 1] = -1
                dG[1][0, 2] = -1
                dG[1][2, 0] = 1
                dG[2][0, 1] = 1
                dG[2][1, 0] = -1
                dG2_dq1 = [sum(dG[k] * dxi2_dq1[k, j] for k in range(3)) for j in range(6)]
                dG2_dq2 = [sum(dG[k] * c["dxi2_dq2"][k, j] for k in range(3)) for j in range(6)]
                dd_dq1 = c["dd_dxw"] @ dxw1_dq1
                dddot_dq1 = c["dddot_dxw"] @ dxw1_dq1 + c["dddot_dxwdot"] @ dxw1dot_dq1
                dn_dq1 = c["dn_dxw"] @ dxw1_dq1
                dddot_dphi1 = c["dddot_dxwdot"] @ dxw1dot_dphi1
                dtdot_dq1 = c["dtdot_dxw"] @ dxw1_dq1 + c["dtdot_dxwdot"] @ dxw1dot_dq1
                dtdot_dphi1 = c["dtdot_dxwdot"] @ dxw1dot_dphi1
                dd_dq2, dddot_dq2, dddot_dphi2 = c["dd_dq2"], c["dddot_dq2"], c["dddot_dphi2"]
                dn_dq2, dtdot_dq2, dtdot_dphi2 = c["dn_dq2"], c["dtdot_dq2"], c["dtdot_dphi2"]
                ds_dq1 = kn * dd_dq1 - damping * dddot_dq1 * d - damping * ddot * dd_dq1
                ds_dq2 = kn * dd_dq2 - damping * dddot_dq2 * d - damping * ddot * dd_dq2
                dfc1_dq1 = -np.outer(ni1, ds_dq1) - s * GTRT1 @ dn_dq1
                dfc1_dq1[:, :3] -= s * G1.T @ skew(R1.T @ n)
                dfc1_dq2 = -np.outer(ni1, ds_dq2) - s * GTRT1 @ dn_dq2
                dfc1_dphi1 = np.outer(ni1, damping * dddot_dphi1 * d)
                dfc1_dphi2 = np.outer(ni1, damping * dddot_dphi2 * d)
                K[:6, :6] += dfc1_dq1
                K[:6, 6:] += dfc1_dq2
                D[:6, :6] += dfc1_dphi1
                D[:6, 6:] += dfc1_dphi2
                R2Tn = R2.T @ n
                tmp1 = np.stack([dG2_dq1[j].T @ R2Tn for j in range(6)], axis=1)
                tmp2 = np.stack([dG2_dq2[j].T @ R2Tn for j in range(6)], axis=1)
                K[6:, :6] += np.outer(ni2, ds_dq1) + s * (tmp1 + GTRT2 @ dn_dq1)
                K[6:, 6:] += np.outer(ni2, ds_dq2) + s * (tmp2 + GTRT2 @ dn_dq2)
                K[6:, 6:9] += s * G2.T @ skew(R2Tn)
                D[6:, :6] -= np.outer(ni2, damping * dddot_dphi1 * d)
                D[6:, 6:] -= np.outer(ni2, damping * dddot_dphi2 * d)
            if mu > EPS:
                if deriv:
                    with np.errstate(divide="ignore", invalid="ignore"):
                        dfcn_dq1 = (1.0 / fc_norm) * (fc1 @ dfc1_dq1)
                        dfcn_dq2 = (1.0 / fc_norm) * (fc1 @ dfc1_dq2)
                        dfcn_dphi1 = (1.0 / fc_norm) * (fc1 @ dfc1_dphi1)
                        dfcn_dphi2 = (1.0 / fc_norm) * (fc1 @ dfc1_dphi2)
                        dtn_dq1 = (1.0 / tdot_norm) * (tdot @ dtdot_dq1)
                        dtn_dq2 = (1.0 / tdot_norm) * (tdot @ dtdot_dq2)
                        dtn_dphi1 = (1.0 / tdot_norm) * (tdot @ dtdot_dphi1)
                        dtn_dphi2 = (1.0 / tdot_norm) * (tdot @ dtdot_dphi2)
                    R2Tt = R2.T @ tdot
                    tmp1t = np.stack([dG2_dq1[j].T @ R2Tt for j in range(6)], axis=1)
                    tmp2t = np.stack([dG2_dq2[j].T @ R2Tt for j in range(6)], axis=1)
                if static:
                    fsub[:6] += -kt * GTRT1 @ tdot
                    fsub[6:] += kt * GTRT2 @ tdot
                    if deriv:
                        K[:6, :6] -= kt * GTRT1 @ dtdot_dq1
                        K[:6, :3] -= kt * G1.T @ skew(R1.T @ tdot)
                        K[:6, 6:] -= kt * GTRT1 @ dtdot_dq2
                        D[:6, :6] -= kt * GTRT1 @ dtdot_dphi1
                        D[:6, 6:] -= kt * GTRT1 @ dtdot_dphi2
                        K[6:, :6] += kt * GTRT2 @ dtdot_dq1 + kt * tmp1t
                        K[6:, 6:] += kt * GTRT2 @ dtdot_dq2 + kt * tmp2t
                        K[6:, 6:9] += kt * G2.T @ skew(R2Tt)
                        D[6:, :6] += kt * GTRT2 @ dtdot_dphi1
                        D[6:, 6:] += kt * GTRT2 @ dtdot_dphi2
                else:
                    fsub[:6] += -mu * fc_norm / tdot_norm * GTRT1 @ tdot
                    fsub[6:] += mu * fc_norm / tdot_norm * GTRT2 @ tdot
                    if deriv:
                        def dyn(dfcn, dtn, dtd):
                            return np.outer(tdot, dfcn) - fc_norm / tdot_norm * np.outer(tdot, dtn) + fc_norm * dtd
                        K[:6, :6] -= mu / tdot_norm * GTRT1 @ dyn(dfcn_dq1, dtn_dq1, dtdot_dq1)
                        K[:6, :3] -= mu * G1.T @ skew(fc_norm / tdot_norm * R1.T @ tdot)
                        K[:6, 6:] -= mu / tdot_norm * GTRT1 @ dyn(dfcn_dq2, dtn_dq2, dtdot_dq2)
                        D[:6, :6] -= mu / tdot_norm * GTRT1 @ dyn(dfcn_dphi1, dtn_dphi1, dtdot_dphi1)
                        D[:6, 6:] -= mu / tdot_norm * GTRT1 @ dyn(dfcn_dphi2, dtn_dphi2, dtdot_dphi2)
                        K[6:, :6] += mu / tdot_norm * GTRT2 @ dyn(dfcn_dq1, dtn_dq1, dtdot_dq1) + mu * fc_norm / tdot_norm * tmp1t
                        K[6:, 6:] += mu / tdot_norm * GTRT2 @ dyn(dfcn_dq2, dtn_dq2, dtdot_dq2) + mu * fc_norm / tdot_norm * tmp2t
                        K[6:, 6:9] += mu * G2.T @ skew(fc_norm / tdot_norm * R2Tt)
                        D[6:, :6] += mu / tdot_norm * GTRT2 @ dyn(dfcn_dphi1, dtn_dphi1, dtdot_dphi1)
                        D[6:, 6:] += mu / tdot_norm * GTRT2 @ dyn(dfcn_dphi2, dtn_dphi2, dtdot_dphi2)
        r1, r2 = slice(6 * b1, 6 * b1 + 6), slice(6 * b2, 6 * b2 + 6)
        fm[r1] += fsub[:6]
        fm[r2] += fsub[6:]
        if deriv:
            Km[r1, r1] += K[:6, :6]
            Km[r1, r2] += K[:6, 6:]
            Km[r2, r1] += K[6:, :6]
            Km[r2, r2] += K[6:, 6:]
            Dm[r1, r1] += D[:6, :6]
            Dm[r1, r2] += D[:6, 6:]
            Dm[r2, r1] += D[6:, :6]
            Dm[r2, r2] += D[6:, 6:]

    # ------------------------------------------------------------------ computeMatrices (Simulation.cpp:256-454)
    def compute_matrices(self, deriv):
        self._update()
        n = self.n
        qdot = self.qdot
        J, Jd, dJ, dJd = self._jacobian(deriv)
        Mm = np.concatenate([self.sc.inertia[b] for b in range(self.nj)])
        fm, fr_j, Km, Dm, Kr, Dr = self._forces(deriv)
        MmJ = J * Mm[:, None]
        MmJd = Jd * Mm[:, None]
        M = J.T @ MmJ
        out = dict(J=J, Jdot=Jd, fm=fm)
        if not deriv:
            out.update(M=M, fr=J.T @ (fm - MmJd @ qdot) + fr_j)
            return out
        JTMm = MmJ.T
        JTMmJd = JTMm @ Jd
        fr = J.T @ fm + (-JTMmJd @ qdot) + fr_j
        dM_dq = []
        for k in range(n):
            tmp = dJ[k].T @ MmJ
            dM_dq.append(tmp + tmp.T)
        dphi_dq = np.stack([dJ[k] @ qdot for k in range(n)], axis=1)
        Dqvv = -JTMmJd - JTMm @ dphi_dq
        MmJdqd = MmJd @ qdot
        Kqvv = np.stack([-dJ[k].T @ MmJdqd - JTMm @ (dJd[k] @ qdot) for k in range(n)], axis=1)
        JTDm = J.T @ Dm
        K = Kqvv + Kr + J.T @ Km @ J + JTDm @ dphi_dq
        for k in range(n):
            K[:, k] += dJ[k].T @ fm
        D = Dqvv + Dr + JTDm @ J
        out.update(M=M, fr=fr, dM_dq=dM_dq, K=K, D=D, dphi_dq=dphi_dq, Km=Km, Dm=Dm)
        return out

    # ------------------------------------------------------------------ BDF1 + Newton (Simulation.cpp:1150-1351)
    def _eval_g(self, q1):
        h = self.h
        self._set_state(q1, (q1 - self._q0) / h)
        c = self.compute_matrices(False)
        return c["M"] @ (q1 - self._q0 - h * self._qdot0) - h * h * c["fr"]

    def _eval_g_deriv(self, q1, save):
        h = self.h
        self._set_state(q1, (q1 - self._q0) / h)
        c = self.compute_matrices(True)
        dq = q1 - self._q0 - h * self._qdot0
        g = c["M"] @ dq - h * h * c["fr"]
        H = c["M"] - h * h * c["K"] - h * c["D"]
        for k in range(self.n):
            H[:, k] += c["dM_dq"][k] @ dq
        if save:
            dqp, dqdp = self._extra_derivatives()
            self.tape["M"].append(c["M"])
            self.tape["D"].append(c["D"])
            self.tape["dfr_dqprev"].append(dqp)
            self.tape["dfr_dqdotprev"].append(dqdp)
            self.tape["dg_du"].append(-h * h * self._dfr_du())
        self._last = c
        return g, H

    def _newton(self, x):
        tol = self.sc.tol
        max_newton = max(20 * self.n, self.sc.max_iter)
        fail_strike = 0
        iters = ls = 0
        for _ in range(max_newton):
            iters += 1
            g, H = self._eval_g_deriv(x, False)
            dx = np.linalg.solve(H, -g)          # partialPivLu().solve
            gnorm = np.linalg.norm(g)
            alpha = 1.0
            success = False
            g_new = g
            for _trial in range(self.sc.max_ls):
                ls += 1
                g_new = self._eval_g(x + alpha * dx)
                if np.linalg.norm(g_new) < gnorm:
                    success = True
                    break
                alpha *= 0.5
            if success:
                fail_strike = 0
            else:
                fail_strike += 1
                if fail_strike >= 10:
                    break
            x = x + alpha * dx
            if np.linalg.norm(g_new) < tol:
                break
        self.newton_iters.append(iters)
        self.ls_evals.append(ls)
        return x

    def forward(self, num_steps, save_last_frame_var_only=False):
        """Simulation::forward (Simulation.cpp:1057-1148) with integration_BDF1 (:1325-1351)."""
        if len(self.q_his) == 0:
            raise RuntimeError("[Error] Please call simulation.reset() before simulation.forward().")
        h = self.h
        for i in range(num_steps):
            q0, qd0 = self.q.copy(), self.qdot.copy()
            self._q0, self._qdot0 = q0, qd0
            q1 = self._newton(q0 + h * qd0)
            qd1 = (q1 - q0) / h
            if self.backward_flag:
                _, H = self._eval_g_deriv(q1, True)
                self.tape["H"].append(H)
            self._set_state(q1, qd1)
            self._update()
            self._update_actuator_states(q1, qd1)
            self.q_his.append(q1.copy())
            self.qdot_his.append(qd1.copy())
            if self.backward_flag:
                self.current_backward_step += 1
                if (not save_last_frame_var_only) or i == num_steps - 1:
                    _, dvar = self.variables(True)
                    self.tape["dvar_dq"].append(dvar)
                    if self.ntac > 0:
                        _, dtq, dtqd = self.tactile_with_derivatives()
                        self.tape["dtactile_dq"].append(dtq)
                        self.tape["dtactile_dqdot"].append(dtqd)
                else:
                    self.tape["dvar_dq"].append(np.zeros((self.nvar, self.n)))
                    self.tape["dtactile_dq"].append(np.zeros((self.ntac, self.n)))
                    self.tape["dtactile_dqdot"].append(np.zeros((self.ntac, self.n)))

    # ------------------------------------------------------------------ variables (EndEffector.cpp:31-59)
    def variables(self, deriv=False):
        sc = self.sc
        var = np.zeros(self.nvar)
        dvar = np.zeros((self.nvar, self.n)) if deriv else None
        for e, ee in enumerate(sc.end_effectors):
            j = ee["joint"]
            E = self.J_[j]["E_0j"]
            var[3 * e:3 * e + 3] = E[:3, :3] @ ee["pos"] + E[:3, 3]
            if deriv:
                tmp = np.concatenate([ee["pos"], [1.0]])
                now = j
                while now >= 0:
                    p = sc.parent[now]
                    for i in range(sc.ndof[now]):
                        v = self.J_[now]["dEpj"][i] @ tmp
                        if p >= 0:
                            v = self.J_[p]["E_0j"] @ v
                        dvar[3 * e:3 * e + 3, sc.qoff[now] + i] += v[:3]
                    tmp = self.J_[now]["E_pj"] @ tmp
                    now = p
        return var, dvar

    def get_variables(self):
        return self.variables(False)[0]

    # ------------------------------------------------------------------ tactile (TactileSensor.cpp:29-235)
    def _tactile_values(self, s):
        b1 = s.body
        E1 = self.J_[b1]["E_0i"]
        R1, p1 = E1[:3, :3], E1[:3, 3]
        phi1 = self.J_[b1]["bphi"]
        M = len(s.pos)
        out = np.zeros((M, 3))
        body = [-1] * M
        for i in range(M):
            xi1 = s.pos[i]
            xw1 = R1 @ xi1 + p1
            xw1_dot = R1 @ (skew(xi1).T @ phi1[:3] + phi1[3:])
            for b2 in s.candidates:
                if self._cuboid_distance(b2, xw1) < 0.0:
                    c = self._cuboid_collision(b2, xw1, xw1_dot, False)
                    fc = (-s.kn * c["d"] + s.damping * c["ddot"] * c["d"]) * (R1.T @ c["n"])
                    fc_norm = np.linalg.norm(fc)
                    tn = np.linalg.norm(c["tdot"])
                    ft = np.zeros(3)
                    if s.mu > EPS:
                        if s.mu * fc_norm >= s.kt * tn - EPS:
                            ft = -s.kt * R1.T @ c["tdot"]
                        else:
                            ft = -s.mu * fc_norm / tn * R1.T @ c["tdot"]
                    force = fc + ft
                    out[i] = [force @ s.axis0[i], force @ s.axis1[i], -(force @ s.normal[i])]
                    body[i] = b2
        return out, body

    def get_tactile_force_vector(self):
        if not self.sc.sensors:
            return np.zeros(0)
        return np.concatenate([self._tactile_values(s)[0].reshape(-1) for s in self.sc.sensors])

    def tactile_with_derivatives(self):
        """compute_tactile_values_with_derivatives + Simulation::computeTactileWithDerivatives
        (TactileSensor.cpp:89-235, Simulation.cpp:811-838).  Uses J and dphi_dq of the last
        with-derivative evaluation, as the reference does."""
        n = self.n
        J, dphi_dq = self._last["J"], self._last["dphi_dq"]
        tac, dq_all, dqd_all = [], [], []
        for s in self.sc.sensors:
            b1 = s.body
            E1 = self.J_[b1]["E_0i"]
            R1, p1 = E1[:3, :3], E1[:3, 3]
            R1T = R1.T
            phi1 = self.J_[b1]["bphi"]
            M = len(s.pos)
            val = np.zeros((M, 3))
            dtq = np.zeros((3 * M, n))
            dtqd = np.zeros((3 * M, n))
            for i in range(M):
                xi1 = s.pos[i]
                xw1 = R1 @ xi1 + p1
                v1 = skew(xi1).T @ phi1[:3] + phi1[3:]
                xw1_dot = R1 @ v1
                dxw1_dq1 = np.hstack([-R1 @ skew(xi1), R1])
                dxw1dot_dq1 = np.hstack([-R1 @ skew(v1), np.zeros((3, 3))])
                dxw1dot_dphi1 = np.hstack([R1 @ skew(xi1).T, R1])
                proj = np.stack([s.axis0[i], s.axis1[i], -s.normal[i]], axis=1)
                hit = None
                for b2 in s.candidates:
                    if self._cuboid_distance(b2, xw1) < 0.0:
                        c = self._cuboid_collision(b2, xw1, xw1_dot, True)
                        d, n_, ddot, tdot = c["d"], c["n"], c["ddot"], c["tdot"]
                        n1 = R1T @ n_
                        sm = s.kn * d - s.damping * ddot * d
                        fc = -sm * n1
                        dd_dq1 = c["dd_dxw"] @ dxw1_dq1
                        dddot_dq1 = c["dddot_dxw"] @ dxw1_dq1 + c["dddot_dxwdot"] @ dxw1dot_dq1
                        dn_dq1 = c["dn_dxw"] @ dxw1_dq1
                        dddot_dphi1 = c["dddot_dxwdot"] @ dxw1dot_dphi1
                        dtdot_dq1 = c["dtdot_dxw"] @ dxw1_dq1 + c["dtdot_dxwdot"] @ dxw1dot_dq1
                        dtdot_dphi1 = c["dtdot_dxwdot"] @ dxw1dot_dphi1
                        ds_dq1 = s.kn * dd_dq1 - s.damping * dddot_dq1 * d - s.damping * ddot * dd_dq1
                        ds_dq2 = s.kn * c["dd_dq2"] - s.damping * c["dddot_dq2"] * d - s.damping * ddot * c["dd_dq2"]
                        dfc_dq1 = -np.outer(n1, ds_dq1) - sm * R1T @ dn_dq1
                        dfc_dq1[:, :3] -= sm * skew(R1T @ n_)
                        dfc_dq2 = -np.outer(n1, ds_dq2) - sm * R1T @ c["dn_dq2"]
                        dfc_dphi1 = np.outer(n1, s.damping * dddot_dphi1 * d)
                        dfc_dphi2 = np.outer(n1, s.damping * c["dddot_dphi2"] * d)
                        fc_norm = np.linalg.norm(fc)
                        tn = np.linalg.norm(tdot)
                        ft = np.zeros(3)
                        dft = [np.zeros((3, 6)) for _ in range(4)]
                        if s.mu > EPS:
                            if s.mu * fc_norm >= s.kt * tn - EPS:
                                ft = -s.kt * R1T @ tdot
                                dft[0] = -s.kt * R1T @ dtdot_dq1
                                dft[0][:, :3] -= s.kt * skew(R1T @ tdot)
                                dft[1] = -s.kt * R1T @ c["dtdot_dq2"]
                                dft[2] = -s.kt * R1T @ dtdot_dphi1
                                dft[3] = -s.kt * R1T @ c["dtdot_dphi2"]
                            else:
                                ft = -s.mu * fc_norm / tn * R1T @ tdot

                                def dyn(dfc_, dtd):
                                    dfcn = (1.0 / fc_norm) * (fc @ dfc_)
                                    dtn = (1.0 / tn) * (tdot @ dtd)
                                    return -s.mu / tn * R1T @ (np.outer(tdot, dfcn) - fc_norm / tn * np.outer(tdot, dtn) + fc_norm * dtd)
                                dft[0] = dyn(dfc_dq1, dtdot_dq1)
                                dft[0][:, :3] -= s.mu * skew(fc_norm / tn * R1T @ tdot)
                                dft[1] = dyn(dfc_dq2, c["dtdot_dq2"])
                                dft[2] = dyn(dfc_dphi1, dtdot_dphi1)
                                dft[3] = dyn(dfc_dphi2, c["dtdot_dphi2"])
                        force = fc + ft
                        val[i] = [force @ s.axis0[i], force @ s.axis1[i], -(force @ s.normal[i])]
                        hit = (b2, proj.T @ (dft[0] + dfc_dq1), proj.T @ (dft[1] + dfc_dq2),
                               proj.T @ (dft[2] + dfc_dphi1), proj.T @ (dft[3] + dfc_dphi2))
                if hit is not None:
                    b2, dq1, dq2, dphi1, dphi2 = hit
                    r1, r2 = slice(6 * b1, 6 * b1 + 6), slice(6 * b2, 6 * b2 + 6)
                    rows = slice(3 * i, 3 * i + 3)
                    dtq[rows] += dq1 @ J[r1] + dphi1 @ dphi_dq[r1]
                    dtqd[rows] += dphi1 @ J[r1]
                    dtq[rows] += dq2 @ J[r2] + dphi2 @ dphi_dq[r2]
                    dtqd[rows] += dphi2 @ J[r2]
            tac.append(val.reshape(-1))
            dq_all.append(dtq)
            dqd_all.append(dtqd)
        return np.concatenate(tac), np.vstack(dq_all), np.vstack(dqd_all)

    # ------------------------------------------------------------------ adjoint (Simulation.cpp:1619-1713)
    def backward(self, df_dq, df_dvar, df_dtactile, df_dq0=None, df_dqdot0=None, df_du=None):
        T = len(self.q_his) - 1
        if T < 1:
            raise RuntimeError("[Error] Please call simulation.forward() before simulation.backward().")
        n, nu, nv, nt, h, tp = self.n, self.nu, self.nvar, self.ntac, self.h, self.tape
        df_dq = np.asarray(df_dq, dtype=np.float64).reshape(-1)
        df_dvar = np.asarray(df_dvar, dtype=np.float64).reshape(-1)
        df_dtactile = np.asarray(df_dtactile, dtype=np.float64).reshape(-1)
        if df_dq.size != n * T or df_dvar.size != nv * T or df_dtactile.size != nt * T:
            raise RuntimeError("backward: cotangent size mismatch")
        z = np.zeros((T, n))
        for k in range(T - 1, -1, -1):
            yk = df_dq[k * n:(k + 1) * n] + tp["dvar_dq"][k].T @ df_dvar[k * nv:(k + 1) * nv]
            if nt > 0:
                w = df_dtactile[k * nt:(k + 1) * nt]
                yk = yk + tp["dtactile_dq"][k].T @ w
                yk = yk + 1.0 / h * tp["dtactile_dqdot"][k].T @ w
                if k < T - 1:
                    yk = yk - 1.0 / h * tp["dtactile_dqdot"][k + 1].T @ df_dtactile[(k + 1) * nt:(k + 2) * nt]
            if k < T - 1:
                Hm = -2.0 * tp["M"][k + 1] + h * tp["D"][k + 1] - h * h * (tp["dfr_dqprev"][k + 1] + 1.0 / h * tp["dfr_dqdotprev"][k + 1])
                yk = yk - Hm.T @ z[k + 1]
            if k < T - 2:
                Hm = tp["M"][k + 2] + h * tp["dfr_dqdotprev"][k + 2]
                yk = yk - Hm.T @ z[k + 2]
            z[k] = np.linalg.solve(tp["H"][k].T, yk)
        res = {}
        r = np.zeros(n) if df_dq0 is None else np.array(df_dq0, dtype=np.float64)
        r = r - (-tp["M"][0] + h * tp["D"][0] - h * h * tp["dfr_dqprev"][0]).T @ z[0]
        if T > 1:
            r = r - (tp["M"][1] + h * tp["dfr_dqdotprev"][1]).T @ z[1]
        if nt > 0:
            r = r - 1.0 / h * tp["dtactile_dqdot"][0].T @ df_dtactile[:nt]
        res["df_dq0"] = r
        r = np.zeros(n) if df_dqdot0 is None else np.array(df_dqdot0, dtype=np.float64)
        res["df_dqdot0"] = r - (-h * tp["M"][0] - h * h * tp["dfr_dqdotprev"][0]).T @ z[0]
        r = np.zeros(nu * T) if df_du is None else np.array(df_du, dtype=np.float64).reshape(-1)
        for k in range(T):
            r[k * nu:(k + 1) * nu] -= tp["dg_du"][k].T @ z[k]
        res["df_du"] = r.reshape(T, nu)
        return res

    def backward_steps(self, num, df_dq, df_dvar, df_dtactile, df_du=None):
        """backward_steps_BDF1 (Simulation.cpp:1921-1971)"""
        if self.current_backward_step <= 0:
            raise RuntimeError("[Error] Please call simulation.forward() before simulation.backward().")
        n, nu, nv, nt, h, tp = self.n, self.nu, self.nvar, self.ntac, self.h, self.tape
        T = len(self.q_his) - 1
        df_dq = np.asarray(df_dq, dtype=np.float64).reshape(-1)
        df_dvar = np.asarray(df_dvar, dtype=np.float64).reshape(-1)
        df_dtactile = np.asarray(df_dtactile, dtype=np.float64).reshape(-1)
        r = np.zeros(nu * num) if df_du is None else np.array(df_du, dtype=np.float64).reshape(-1)
        for i in range(num, 0, -1):
            kr = self.current_backward_step - num + i - 1
            k = i - 1
            w = df_dtactile[k * nt:(k + 1) * nt]
            yk = df_dq[k * n:(k + 1) * n] + tp["dvar_dq"][kr].T @ df_dvar[k * nv:(k + 1) * nv]
            yk = yk + tp["dtactile_dq"][kr].T @ w
            yk = yk + 1.0 / h * tp["dtactile_dqdot"][kr].T @ w
            if kr < T - 1:
                yk = yk - 1.0 / h * tp["dtactile_dqdot"][kr + 1].T @ self.df_dtactile_his[-1]
                Hm = -2.0 * tp["M"][kr + 1] + h * tp["D"][kr + 1] - h * h * (tp["dfr_dqprev"][kr + 1] + 1.0 / h * tp["dfr_dqdotprev"][kr + 1])
                yk = yk - Hm.T @ self.z_his[-1]
            if kr < T - 2:
                Hm = tp["M"][kr + 2] + h * tp["dfr_dqdotprev"][kr + 2]
                yk = yk - Hm.T @ self.z_his[-2]
            self.z_his.append(np.linalg.solve(tp["H"][kr].T, yk))
            r[k * nu:(k + 1) * nu] -= tp["dg_du"][kr].T @ self.z_his[-1]
            self.df_dtactile_his.append(w.copy())
        self.current_backward_step -= num
        return {"df_du": r.reshape(num, nu)}
