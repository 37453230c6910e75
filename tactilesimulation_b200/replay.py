"""Trajectory export in the reference's ``export_replay`` format (``DH/Simulation.cpp:2037-2120``): the way to look at a
rollout of the B200 path without the OpenGL viewer.

    <folder>/meshes/<k>.obj   one mesh per body, render-only object, tactile sensor and end-effector (radius > 0)
    <folder>/<i>.txt          frame i of the q history (frame 0 = the initial state): first line the number of meshes,
                              then one 4 x 4 world transform per mesh, rows as lines, "%.6lf " per entry -- bodies
                              (E_0i), render-only objects, sensors (the transform of their body), end-effectors
                              (joint frame with the end-effector's world position as translation)

Host-side numpy forward kinematics of the scene (no dynamics): joint transforms as ``DH/Joint/Joint*.cpp update``,
recursion ``DH/Joint/Joint.cpp:119-165``, body frames ``DH/Body/Body.cpp:122-165``.  The frame files are checked against
the reference's own export (tests/test_replay_export.py; bodies built from a mesh sit in the principal-axes frame of
their inertia, whose axis signs are Eigen's in the reference and numpy's here -- same frame up to a flip of two axes,
with the exported mesh given in the same frame).  Meshes: primitives are tessellated here (the reference
writes its rendering meshes); mesh / abstract bodies are exported as their sampled contact points (vertices only) --
enough to see the motion, not the reference's triangle soup.
"""
import math
import os

import numpy as np

from .scene import (JT_FIXED, JT_FREE2D, JT_FREE3D_EULER, JT_FREE3D_EXP, JT_PLANAR, JT_PRISMATIC, JT_REVOLUTE,
                    JT_SPHERICAL_EULER, JT_SPHERICAL_EXP, JT_TRANSLATIONAL, SH_CAPSULE, SH_CUBOID, SH_CYLINDER, SH_SPHERE,
                    Scene, quat2mat)


def _axis_angle(a, th):
    c, s = math.cos(th), math.sin(th)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) * c + s * K + (1 - c) * np.outer(a, a)


def _euler_xyz(r):
    """R = Rx(r0) Ry(r1) Rz(r2)  (DH/Joint/JointSphericalEuler.cpp, chart XYZ)."""
    ex, ey, ez = np.eye(3)
    return _axis_angle(ex, r[0]) @ _axis_angle(ey, r[1]) @ _axis_angle(ez, r[2])


def _exp_so3(r):
    """math::exp (DH/Utils.h:82-98)."""
    th = float(np.linalg.norm(r))
    if th < 1e-12:
        return np.eye(3)
    return _axis_angle(np.asarray(r) / th, th)


def joint_transform(jt, a0, a1, q):
    """Q(q) of one joint."""
    Q = np.eye(4)
    if jt == JT_FIXED:
        return Q
    if jt == JT_REVOLUTE:
        Q[:3, :3] = _axis_angle(a0, q[0])
    elif jt == JT_PRISMATIC:
        Q[:3, 3] = a0 * q[0]
    elif jt == JT_PLANAR:
        Q[:3, 3] = a0 * q[0] + a1 * q[1]
    elif jt == JT_TRANSLATIONAL:
        Q[:3, 3] = q[:3]
    elif jt == JT_FREE2D:
        Q[:3, :3] = _axis_angle(np.array([0.0, 0.0, 1.0]), q[2])
        Q[:2, 3] = q[:2]
    elif jt == JT_FREE3D_EULER:
        Q[:3, :3] = _euler_xyz(q[3:6])
        Q[:3, 3] = q[:3]
    elif jt == JT_FREE3D_EXP:
        Q[:3, :3] = _exp_so3(q[3:6])
        Q[:3, 3] = q[:3]
    elif jt == JT_SPHERICAL_EULER:
        Q[:3, :3] = _euler_xyz(q[:3])
    elif jt == JT_SPHERICAL_EXP:
        Q[:3, :3] = _exp_so3(q[:3])
    else:
        raise ValueError("joint type %d" % jt)
    return Q


def frames(sc: Scene, q):
    """World frames of every joint (E_0j) and body (E_0i) at reduced coordinates q."""
    q = np.asarray(q, dtype=np.float64)
    E_0j, E_0i = [], []
    for j in range(sc.nj):
        o, nd, p = sc.qoff[j], sc.ndof[j], sc.parent[j]
        E_pj = sc.E_pj0[j] @ joint_transform(sc.jtype[j], sc.axis0[j], sc.axis1[j], q[o:o + nd])
        E = E_pj if p < 0 else E_0j[p] @ E_pj
        E_0j.append(E)
        E_0i.append(E @ sc.E_ji[j])
    return E_0j, E_0i


def _write_obj(path, V, F=None):
    with open(path, "w") as f:
        for v in V:
            f.write("v %.6f %.6f %.6f\n" % (v[0], v[1], v[2]))
        if F is not None:
            for t in F:
                f.write("f %d %d %d\n" % (t[0] + 1, t[1] + 1, t[2] + 1))


def _box_mesh(size):
    h = np.asarray(size, dtype=np.float64)[:3] / 2.0
    V = np.array([[sx * h[0], sy * h[1], sz * h[2]] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)])
    F = [(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6), (0, 2, 6), (0, 6, 4), (1, 5, 7), (1, 7, 3)]
    return V, F


def _round_mesh(radius, half_len, n=16, caps=False):
    """cylinder about z (or a capsule / sphere as a stack of rings)."""
    rings = []
    if caps:
        for k in range(1, 5):
            a = math.pi / 2 * k / 4
            rings.append((-half_len - radius * math.cos(a), radius * math.sin(a)))
    rings += [(-half_len, radius), (half_len, radius)]
    if caps:
        for k in range(3, -1, -1):
            a = math.pi / 2 * k / 4
            rings.append((half_len + radius * math.cos(a), radius * math.sin(a)))
    V, F = [], []
    for z, r in rings:
        for i in range(n):
            V.append([r * math.cos(2 * math.pi * i / n), r * math.sin(2 * math.pi * i / n), z])
    for k in range(len(rings) - 1):
        for i in range(n):
            a, b, c, d = k * n + i, k * n + (i + 1) % n, (k + 1) * n + i, (k + 1) * n + (i + 1) % n
            F += [(a, b, d), (a, d, c)]
    return np.array(V), F


def body_mesh(sc: Scene, b: int):
    sh, size = sc.shape[b], sc.size[b]
    if sh == SH_CUBOID:
        return _box_mesh(size)
    if sh == SH_CYLINDER:
        return _round_mesh(size[0], size[1] / 2.0)
    if sh == SH_SPHERE:
        return _round_mesh(size[0], 0.0, caps=True)
    if sh == SH_CAPSULE:
        return _round_mesh(size[0], size[1] / 2.0, caps=True)
    pts = sc.contact_points[b] if b < len(sc.contact_points) and len(sc.contact_points[b]) else np.zeros((1, 3))
    return np.asarray(pts, dtype=np.float64).reshape(-1, 3), None


def export_replay(sc: Scene, q_history, folder: str, virtual_pose=None):
    """Writes <folder>/meshes/*.obj and <folder>/<i>.txt for every state of q_history ([frames, ndof_r])."""
    os.makedirs(os.path.join(folder, "meshes"), exist_ok=True)
    idx = 0
    for b in range(sc.nj):
        V, F = body_mesh(sc, b)
        _write_obj(os.path.join(folder, "meshes", "%d.obj" % idx), V, F)
        idx += 1
    vposes = list(virtual_pose if virtual_pose is not None else sc.virtual_pose)
    for _ in vposes:
        V, F = _box_mesh([0.05, 0.05, 0.05])
        _write_obj(os.path.join(folder, "meshes", "%d.obj" % idx), V, F)
        idx += 1
    for s in sc.sensors:
        _write_obj(os.path.join(folder, "meshes", "%d.obj" % idx), np.asarray(s.pos, dtype=np.float64).reshape(-1, 3))
        idx += 1
    ees = [e for e in sc.end_effectors if e.get("radius", 0.1) > 0.0]
    for e in ees:
        V, F = _round_mesh(e.get("radius", 0.1), 0.0, n=8, caps=True)
        _write_obj(os.path.join(folder, "meshes", "%d.obj" % idx), V, F)
        idx += 1

    def mat(f, E):
        for j in range(4):
            f.write("".join("%.6f " % E[j, k] for k in range(4)) + "\n")
    for i, q in enumerate(np.asarray(q_history, dtype=np.float64)):
        E_0j, E_0i = frames(sc, q)
        with open(os.path.join(folder, "%d.txt" % i), "w") as f:
            f.write("%d\n" % idx)
            for E in E_0i:
                mat(f, E)
            for vp in vposes:
                E = np.eye(4)
                E[:3, :3] = quat2mat(vp[3:7])
                E[:3, 3] = vp[:3]
                mat(f, E)
            for s in sc.sensors:
                mat(f, E_0i[s.body])
            for e in ees:
                j = e["joint"]
                E = np.eye(4) if j < 0 else E_0j[j].copy()
                E[:3, 3] = (E_0j[j] @ np.append(e["pos"], 1.0))[:3] if j >= 0 else e["pos"]
                mat(f, E)
    return idx
