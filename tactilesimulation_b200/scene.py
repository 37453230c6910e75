"""Scene compiler: redmax XML -> flat, batch-invariant topology tables.

This is the host-side loader of the B200 simulator.  It reads the same
``<redmax>`` XML scene files the reference reads
(``DH/Simulation_Constructor.cpp:70-705``; DH = externals/DiffHand/core/projects/redmax)
and produces a :class:`Scene` of plain numpy tables which ``Scene.pack()`` flattens into
the two buffers (int32 / float64) that the CUDA kernels index directly.

Semantics reproduced from the reference loader (behaviour, not code):

* attributes parsed with pugixml ``as_float()`` are rounded through fp32
  (timestep, damping, density, radius, length, kn/kt/mu/damping, P, D, ...),
  ``Simulation_Constructor.cpp:108,230-251,510,559,578-579``; vectors go through
  ``str_to_eigen`` in double (``Utils.h:388-402``).
* joints are numbered in DFS order over ``<link>`` nesting, reduced dofs are assigned
  in that order, every link carries one body (``Robot.cpp:38-100``).
* ``<default>`` fallbacks for joint damping / lim_stiffness, contact and tactile
  coefficients, motor ctrl_range/P/D.
* body mass properties: cuboid ``BodyCuboid.cpp:20-26``, cylinder
  ``BodyCylinder.cpp:21-27``, sphere ``BodySphere.cpp:17-21``, mesh (volume integral +
  principal axes) ``BodyMeshObj.cpp:69-173``.
* contact sample points: cuboid surface grid ``BodyCuboid.cpp:28-41``, cylinder faces
  ``BodyCylinder.cpp:30-45``.
* rect_array marker grid, row-major with the axis0 index outermost
  ``TactileSensorRectArray.cpp:43-68``.
* tactile candidate bodies = every primitive-shape body except the pad, in body order
  (``TactileSensor.cpp:44-46``).

Unsupported scene features raise :class:`SceneError` loudly (no silent fallback).
"""
from __future__ import annotations

import math
import os
import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# joint types (values are shared with csrc/scene_layout.h)
JT_FIXED, JT_REVOLUTE, JT_PRISMATIC, JT_PLANAR, JT_TRANSLATIONAL, JT_FREE3D_EULER, JT_FREE3D_EXP = 0, 1, 2, 3, 4, 5, 6
JT_SPHERICAL_EULER, JT_SPHERICAL_EXP, JT_FREE2D = 7, 8, 9
JOINT_NDOF = {JT_FIXED: 0, JT_REVOLUTE: 1, JT_PRISMATIC: 1, JT_PLANAR: 2, JT_TRANSLATIONAL: 3, JT_FREE3D_EULER: 6,
              JT_FREE3D_EXP: 6, JT_SPHERICAL_EULER: 3, JT_SPHERICAL_EXP: 3, JT_FREE2D: 3}
# "free3d" is the XYZ-Euler chart in the reference too (DH/Simulation_Constructor.cpp:483-484)
JOINT_TYPES = {"fixed": JT_FIXED, "revolute": JT_REVOLUTE, "prismatic": JT_PRISMATIC,
               "planar": JT_PLANAR, "translational": JT_TRANSLATIONAL,
               "free3d": JT_FREE3D_EULER, "free3d-euler": JT_FREE3D_EULER, "free3d-exp": JT_FREE3D_EXP,
               "spherical": JT_SPHERICAL_EULER, "spherical-euler": JT_SPHERICAL_EULER, "spherical-exp": JT_SPHERICAL_EXP, "free2d": JT_FREE2D}
# time integrators (DH/Simulation.cpp:1076-1092); codes shared with csrc/scene_layout.h
INTEGRATORS = {"BDF1": 0, "BDF2": 1, "SDIRK2": 2}
# body shapes
SH_NONE, SH_CUBOID, SH_CYLINDER, SH_SPHERE, SH_CAPSULE = 0, 1, 2, 3, 4
# actuator modes
ACT_FORCE, ACT_POS = 0, 1

INT_MIN, INT_MAX = -2147483648, 2147483647


class SceneError(RuntimeError):
    pass


def _f32(s) -> float:
    """pugixml as_float(): strtod then cast to float, widened back to double."""
    return float(np.float32(float(s)))


def _vec(s: str) -> np.ndarray:
    return np.array([float(t) for t in s.split()], dtype=np.float64)


def _ivec(s: str) -> np.ndarray:
    return np.array([int(t) for t in s.split()], dtype=np.int64)


def quat2mat(quat) -> np.ndarray:
    """w x y z quaternion -> rotation (normalised first), ``Utils.h:142-150``."""
    q = np.asarray(quat, dtype=np.float64)
    q = q / np.linalg.norm(q)
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def SE(R, p) -> np.ndarray:
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = p
    return E


def Einv(E) -> np.ndarray:
    Rt = E[:3, :3].T
    return SE(Rt, -Rt @ E[:3, 3])


def _attr(node, root_default, tag, name):
    """attribute with <default><tag name=.../> fallback; None if absent in both."""
    if node.get(name) is not None:
        return node.get(name)
    if root_default is not None:
        d = root_default.find(tag)
        if d is not None and d.get(name) is not None:
            return d.get(name)
    return None


def _load_obj(path: str):
    """Vertices (through fp32, as tinyobj's real_t) and triangulated faces of the first
    shape of a Wavefront OBJ (``BodyMeshObj.cpp:45-67``)."""
    V, F = [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                t = line.split()
                V.append([float(np.float32(float(t[1]))), float(np.float32(float(t[2]))),
                          float(np.float32(float(t[3])))])
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    i = int(tok.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(V) + i)
                for k in range(1, len(idx) - 1):
                    F.append([idx[0], idx[k], idx[k + 1]])
    if not V or not F:
        raise SceneError(f"mesh {path} has no geometry")
    return np.array(V, dtype=np.float64), np.array(F, dtype=np.int64)


def mesh_mass_properties(V: np.ndarray, F: np.ndarray):
    """volume, centre of mass and unit-mass inertia tensor of a closed triangle mesh
    (signed tetrahedra against the origin), following ``BodyMeshObj.cpp:118-173``."""
    A = V[F]                                    # (nf, 3 verts, 3 coords)
    vol_f = np.linalg.det(np.transpose(A, (0, 2, 1)))
    volume = vol_f.sum()
    COM = (vol_f[:, None] * A.sum(axis=1)).sum(axis=0) / (volume * 4.0)
    volume = volume / 6.0
    B = A - COM                                 # rows = vertices, cols = coords
    d = np.linalg.det(B)
    diag = np.zeros(3)
    offd = np.zeros(3)
    for j in range(3):
        j1, j2 = (j + 1) % 3, (j + 2) % 3
        a = B[:, :, j]
        diag[j] = ((a[:, 0] * a[:, 1] + a[:, 1] * a[:, 2] + a[:, 2] * a[:, 0]
                    + a[:, 0] ** 2 + a[:, 1] ** 2 + a[:, 2] ** 2) * d).sum()
        b, c = B[:, :, j1], B[:, :, j2]
        offd[j] = ((b[:, 0] * c[:, 1] + b[:, 1] * c[:, 2] + b[:, 2] * c[:, 0]
                    + b[:, 0] * c[:, 2] + b[:, 1] * c[:, 0] + b[:, 2] * c[:, 1]
                    + 2 * b[:, 0] * c[:, 0] + 2 * b[:, 1] * c[:, 1] + 2 * b[:, 2] * c[:, 2]) * d).sum()
    diag /= volume * 60.0
    offd /= volume * 120.0
    I = np.array([[diag[1] + diag[2], -offd[2], -offd[1]],
                  [-offd[2], diag[0] + diag[2], -offd[0]],
                  [-offd[1], -offd[0], diag[0] + diag[1]]])
    return volume, COM, I


@dataclass
class TactileSensor:
    name: str
    body: int
    kn: float
    kt: float
    mu: float
    damping: float
    pos: np.ndarray          # (M,3) marker positions in the pad body frame
    axis0: np.ndarray        # (M,3)
    axis1: np.ndarray        # (M,3)
    normal: np.ndarray       # (M,3)
    image_pos: np.ndarray    # (M,2) int
    candidates: List[int] = field(default_factory=list)


@dataclass
class Scene:
    name: str = ""
    h: float = 0.01
    gravity: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, -980.0]))
    integrator: str = "BDF2"
    tol: float = 1e-9
    max_iter: int = 10
    max_ls: int = 20
    # joints / bodies, DFS order (one body per joint)
    joint_names: List[str] = field(default_factory=list)
    body_names: List[str] = field(default_factory=list)
    jtype: List[int] = field(default_factory=list)
    parent: List[int] = field(default_factory=list)
    qoff: List[int] = field(default_factory=list)
    ndof: List[int] = field(default_factory=list)
    E_pj0: List[np.ndarray] = field(default_factory=list)
    E_j0_0: List[np.ndarray] = field(default_factory=list)
    axis0: List[np.ndarray] = field(default_factory=list)
    axis1: List[np.ndarray] = field(default_factory=list)
    damping: List[float] = field(default_factory=list)
    lim_lo: List[float] = field(default_factory=list)
    lim_hi: List[float] = field(default_factory=list)
    lim_k: List[float] = field(default_factory=list)
    E_ji: List[np.ndarray] = field(default_factory=list)
    inertia: List[np.ndarray] = field(default_factory=list)   # (6,) Ixx Iyy Izz m m m
    shape: List[int] = field(default_factory=list)
    btype: List[str] = field(default_factory=list)            # XML body type
    size: List[np.ndarray] = field(default_factory=list)      # cuboid: lengths; cylinder: (r, l, 0)
    contact_points: List[np.ndarray] = field(default_factory=list)
    E_g: np.ndarray = field(default_factory=lambda: np.eye(4))
    has_ground: bool = False
    ground_contacts: List[dict] = field(default_factory=list)  # body,kn,kt,mu,damping
    gp_contacts: List[dict] = field(default_factory=list)      # body1,body2,kn,kt,mu,damping
    actuators: List[dict] = field(default_factory=list)        # joint,mode,cmin,cmax,P,D,uoff
    end_effectors: List[dict] = field(default_factory=list)    # joint,pos
    sensors: List[TactileSensor] = field(default_factory=list)
    virtual_names: List[str] = field(default_factory=list)
    virtual_pose: List[np.ndarray] = field(default_factory=list)   # host-only: (pos, quat wxyz) of each render-only object
    # host-only construction parameters (not in the blob): what the update_* calls of the reference need to redo a part
    # of the construction (DH/Robot.cpp:571-650).  Empty for scenes rebuilt from a blob.
    body_density: List[float] = field(default_factory=list)
    body_res: List[Optional[np.ndarray]] = field(default_factory=list)     # contact sampling resolution of the body
    joint_R0: List[np.ndarray] = field(default_factory=list)
    joint_frame: List[str] = field(default_factory=list)
    ndof_r: int = 0
    ndof_u: int = 0

    # ------------------------------------------------------------------ sizes
    @property
    def nj(self) -> int:
        return len(self.jtype)

    @property
    def ndof_m(self) -> int:
        return 6 * self.nj

    @property
    def ndof_var(self) -> int:
        return 3 * len(self.end_effectors)

    @property
    def ndof_tactile(self) -> int:
        return 3 * sum(len(s.pos) for s in self.sensors)

    @property
    def n_markers(self) -> int:
        return sum(len(s.pos) for s in self.sensors)

    # ------------------------------------------------------------ serialisation
    def to_npz_dict(self) -> Dict[str, np.ndarray]:
        ib, db = self.pack()
        return {"ibuf": ib, "dbuf": db,
                "names": np.array([self.name] + self.joint_names + self.body_names
                                  + [s.name for s in self.sensors])}

    # ------------------------------------------------------------------ pack
    def pack(self):
        """Flatten to (int32 ibuf, float64 dbuf) in the layout of csrc/scene_layout.h."""
        from .layout import pack_scene
        return pack_scene(self)


def _cuboid_points(length, res):
    pts = []
    for i in range(res[0]):
        for j in range(res[1]):
            for k in range(res[2]):
                if i in (0, res[0] - 1) or j in (0, res[1] - 1) or k in (0, res[2] - 1):
                    pts.append([i / (res[0] - 1) * length[0] - length[0] / 2.0,
                                j / (res[1] - 1) * length[1] - length[1] / 2.0,
                                k / (res[2] - 1) * length[2] - length[2] / 2.0])
    return np.array(pts, dtype=np.float64)


def _cylinder_points(radius, length, ares, rres):
    pts = []
    for dz in (-1, 1):
        pts.append([0.0, 0.0, dz * length / 2.0])
        for a in range(ares):
            angle = math.pi * 2.0 / ares * a
            for r in range(rres):
                pts.append([0.0 + math.cos(angle) * (r + 1) / rres * radius,
                            0.0 + math.sin(angle) * (r + 1) / rres * radius,
                            dz * length / 2.0 + 0.0])
    return np.array(pts, dtype=np.float64)


def _capsule_points(radius, length, res):
    """DH/Body/BodyCapsule.cpp:34-50: the two poles, then res[0] rings of res[1] points on the cylindrical part."""
    pts = [[0.0, 0.0, length / 2.0 + radius], [0.0, 0.0, -length / 2.0 - radius]]
    for i in range(res[0]):
        z = i / (res[0] - 1) * length - length / 2.0
        for j in range(res[1]):
            pts.append([radius * math.cos(2.0 * math.pi * j / res[1]), radius * math.sin(2.0 * math.pi * j / res[1]), z])
    return np.array(pts, dtype=np.float64)


def compile_scene(xml_path: str) -> Scene:
    """Parse a redmax XML scene file into a :class:`Scene`."""
    if not os.path.isfile(xml_path):
        raise SceneError("Input Model File (.xml) is incorrect.")
    try:
        root = ET.parse(xml_path).getroot()
    except ET.ParseError as e:
        raise SceneError("Input Model File (.xml) is incorrect.") from e
    if root.tag != "redmax":
        raise SceneError("Input Model File (.xml) is incorrect.")
    asset_dir = os.path.dirname(os.path.abspath(xml_path))
    sc = Scene(name=root.get("model", ""))
    default = root.find("default")

    opt = root.find("option")
    if opt is not None:
        if opt.get("gravity") is not None:
            sc.gravity = _vec(opt.get("gravity"))
        if opt.get("timestep") is not None:
            sc.h = _f32(opt.get("timestep"))
        if opt.get("integrator") is not None:
            sc.integrator = opt.get("integrator")
        unit = opt.get("unit", "cm-g")
    else:
        unit = "cm-g"
    so = root.find("solver_option")
    if so is not None:
        if so.get("tol") is not None:
            sc.tol = float(so.get("tol"))
        if so.get("max_iter") is not None:
            sc.max_iter = int(so.get("max_iter"))
        if so.get("max_ls") is not None:
            sc.max_ls = int(so.get("max_ls"))

    g = root.find("ground")
    if g is not None:
        sc.has_ground = True
        pos = _vec(g.get("pos"))
        nz = _vec(g.get("normal"))
        nz = nz / np.linalg.norm(nz)
        # rotation taking +z to nz (Eigen setFromTwoVectors); only the normal and the
        # origin enter the dynamics (ForceGroundContact.cpp:106-107)
        z = np.array([0.0, 0.0, 1.0])
        c = float(z @ nz)
        if c > 1.0 - 1e-12:
            Rg = np.eye(3)
        elif c < -1.0 + 1e-12:
            Rg = np.diag([1.0, -1.0, -1.0])
        else:
            ax = np.cross(z, nz)
            s = np.linalg.norm(ax)
            ax = ax / s
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            Rg = np.eye(3) + s * K + (1 - c) * (K @ K)
        Rg[:, 2] = nz
        # the reference scales the origin by 10 for the viewer and back (cm-g: /10 then *10)
        p = (pos / 10.0) * 10.0 if unit == "cm-g" else (pos * 10.0) / 10.0
        sc.E_g = SE(Rg, p)

    joint_map: Dict[str, int] = {}
    body_map: Dict[str, int] = {}

    def parse_link(node, parent_idx):
        jn = node.find("joint")
        bn = node.find("body")
        if jn is None or bn is None:
            raise SceneError("every <link> needs a <joint> and a <body>")
        tname = jn.get("type")
        if tname not in JOINT_TYPES:
            raise SceneError(f"Joint type not supported by the B200 path yet: {tname}")
        jt = JOINT_TYPES[tname]
        pos = _vec(jn.get("pos"))
        R = quat2mat(_vec(jn.get("quat")))
        frame = jn.get("frame", "LOCAL")
        if frame not in ("LOCAL", "WORLD"):
            raise SceneError("Frame type error: " + frame)
        idx = sc.nj
        if frame == "WORLD" and parent_idx >= 0:
            E_pj0 = sc.E_j0_0[parent_idx] @ SE(R, pos)
        else:
            E_pj0 = SE(R, pos)
        E_jp0 = Einv(E_pj0)
        E_j00 = E_jp0 if parent_idx < 0 else E_jp0 @ sc.E_j0_0[parent_idx]
        a0 = np.zeros(3)
        a1 = np.zeros(3)
        if jt in (JT_REVOLUTE, JT_PRISMATIC):
            a0 = _vec(jn.get("axis"))
            a0 = a0 / np.linalg.norm(a0)
            if frame == "WORLD":
                a0 = E_j00[:3, :3] @ a0
        elif jt == JT_PLANAR:
            a0 = _vec(jn.get("axis0"))
            a1 = _vec(jn.get("axis1"))
            a0 = a0 / np.linalg.norm(a0)
            a1 = a1 / np.linalg.norm(a1)
            if frame == "WORLD":
                a0 = E_j00[:3, :3] @ a0
                a1 = E_j00[:3, :3] @ a1
        sc.jtype.append(jt)
        sc.parent.append(parent_idx)
        sc.ndof.append(JOINT_NDOF[jt])
        sc.qoff.append(-1)
        sc.E_pj0.append(E_pj0)
        sc.E_j0_0.append(E_j00)
        sc.joint_R0.append(R)
        sc.joint_frame.append(frame)
        sc.axis0.append(a0)
        sc.axis1.append(a1)
        name = jn.get("name", f"joint{idx}")
        sc.joint_names.append(name)
        if jn.get("name") is not None:
            joint_map[name] = idx
        d = _attr(jn, default, "joint", "damping")
        sc.damping.append(_f32(d) if d is not None else 0.0)
        if jn.get("lim") is not None:
            lim = _vec(jn.get("lim"))
            lk = _attr(jn, default, "joint", "lim_stiffness")
            sc.lim_lo.append(float(lim[0]))
            sc.lim_hi.append(float(lim[1]))
            sc.lim_k.append(_f32(lk) if lk is not None else 0.0)
        else:
            sc.lim_lo.append(float(INT_MIN))
            sc.lim_hi.append(float(INT_MAX))
            sc.lim_k.append(0.0)

        # ---------------------------------------------------------------- body
        btype = bn.get("type")
        bpos = _vec(bn.get("pos"))
        bR = quat2mat(_vec(bn.get("quat")))
        dens = _attr(bn, default, "body", "density")
        density = _f32(dens) if dens is not None else 1.0
        sc_attr = _attr(bn, default, "body", "scale")
        scale = _vec(sc_attr) if sc_attr is not None else np.ones(3)
        inertia = np.zeros(6)
        size = np.zeros(3)
        pts = np.zeros((0, 3))
        shape = SH_NONE
        E_ji = SE(bR, bpos)
        if btype == "cuboid":
            length = _vec(bn.get("size"))
            res = _ivec(bn.get("general_contact_resolution")) if bn.get("general_contact_resolution") else np.array([6, 6, 6])
            if (res <= 1).any():
                raise SceneError("General contact resolution of cuboid should be at least 2.")
            mass = float(np.prod(length)) * density
            inertia[0] = mass / 12.0 * (length[1] * length[1] + length[2] * length[2])
            inertia[1] = mass / 12.0 * (length[0] * length[0] + length[2] * length[2])
            inertia[2] = mass / 12.0 * (length[0] * length[0] + length[1] * length[1])
            inertia[3:] = mass
            shape, size, pts = SH_CUBOID, length, _cuboid_points(length, res)
        elif btype == "cylinder":
            length = _f32(bn.get("length"))
            radius = _f32(bn.get("radius"))
            ares = int(bn.get("general_contact_angle_resolution", 8))
            rres = int(bn.get("general_contact_radius_resolution", 3))
            mass = math.pi * radius * radius * length * density
            inertia[0] = mass * length * length / 12.0 + mass * radius * radius / 4.0
            inertia[1] = mass * length * length / 12.0 + mass * radius * radius / 4.0
            inertia[2] = mass * radius * radius / 2.0
            inertia[3:] = mass
            shape, size = SH_CYLINDER, np.array([radius, length, 0.0])
            pts = _cylinder_points(radius, length, ares, rres)
        elif btype == "capsule":
            # DH/Simulation_Constructor.cpp:589-596, DH/Body/BodyCapsule.cpp:22-31 (axis z)
            length = _f32(bn.get("length"))
            radius = _f32(bn.get("radius"))
            res = _ivec(bn.get("general_contact_resolution")) if bn.get("general_contact_resolution") else np.array([5, 4])
            m_cy = density * length * math.pi * radius * radius
            m_hs = density * 2.0 / 3.0 * math.pi * radius * radius * radius
            mass = m_cy + 2.0 * m_hs
            inertia[0] = m_cy * (length * length / 12.0 + radius * radius / 4.0) + 2.0 * m_hs * (2.0 * radius * radius / 5.0 + length * length / 2.0 + 3.0 * length * radius / 8.0)
            inertia[1] = inertia[0]
            inertia[2] = m_cy * radius * radius / 2.0 + 2.0 * m_hs * 2.0 * radius * radius / 5.0
            inertia[3:] = mass
            shape, size = SH_CAPSULE, np.array([radius, length, 0.0])
            pts = _capsule_points(radius, length, res)
        elif btype == "sphere":
            # DH/Simulation_Constructor.cpp:597-599, DH/Body/BodySphere.cpp:17-21; no sampled contact points: a sphere
            # is a primitive body and touches the ground at one state-dependent point (CollisionDetection.cpp:17-25)
            radius = _f32(bn.get("radius"))
            mass = 4.0 / 3.0 * math.pi * radius * radius * radius * density
            inertia[:3] = 0.4 * mass * radius * radius
            inertia[3:] = mass
            shape, size = SH_SPHERE, np.array([radius, 0.0, 0.0])
        elif btype == "mesh":
            V, F = _load_obj(os.path.join(asset_dir, bn.get("filename")))
            V = V * scale[None, :]
            volume, COM, I = mesh_mass_properties(V, F)
            mass = volume * density
            I = I * mass
            w, vecs = np.linalg.eigh(I)
            if np.dot(np.cross(vecs[:, 0], vecs[:, 1]), vecs[:, 2]) < 0.0:
                vecs[:, 2] *= -1.0
            inertia[:3] = w
            inertia[3:] = mass
            E_oi = SE(vecs, COM)
            tt = bn.get("transform_type", "BODY_TO_JOINT")
            if tt == "BODY_TO_JOINT":
                E_ji = SE(bR, bpos)
            elif tt == "OBJ_TO_WORLD":
                E_ji = E_j00 @ SE(bR, bpos) @ E_oi
            elif tt == "OBJ_TO_JOINT":
                E_ji = SE(bR @ E_oi[:3, :3], bR @ E_oi[:3, 3] + bpos)
            else:
                raise SceneError("Transform type error: " + tt)
            # mesh vertices sampled as contact points are never used by the supported
            # force types (general_primitive_contact needs them only on the general body)
        elif btype == "abstract":
            # DH/Simulation_Constructor.cpp:619-657, DH/Body/BodyAbstract.cpp:7-75: given mass and principal
            # inertia; contact points from a text file (read as fp32), mapped by the <collision> frame
            mass = _f32(bn.get("mass"))
            I = _vec(bn.get("inertia"))
            if I.shape[0] != 3:
                raise SceneError("abstract bodies with a full inertia tensor are not supported by the B200 path yet")
            inertia[:3] = I
            inertia[3:] = mass
            cfile, cpos, cR = bn.get("contacts"), np.zeros(3), np.eye(3)
            cn = bn.find("collision")
            if cfile is None and cn is not None and cn.get("contacts") is not None:
                cfile = cn.get("contacts")
                if cn.get("pos") is not None:
                    cpos = _vec(cn.get("pos"))
                if cn.get("quat") is not None:
                    cR = quat2mat(_vec(cn.get("quat")))
            if cfile is not None:
                toks = open(os.path.join(asset_dir, cfile)).read().split()
                npts = int(toks[0])
                raw = np.array([np.float32(t) for t in toks[1:1 + 3 * npts]], dtype=np.float64).reshape(npts, 3)
                pts = raw @ cR.T + cpos[None, :]
        else:
            raise SceneError(f"Body type not supported by the B200 path yet: {btype}")
        sc.E_ji.append(E_ji)
        sc.inertia.append(inertia)
        sc.shape.append(shape)
        sc.btype.append(btype)
        sc.size.append(np.asarray(size, dtype=np.float64))
        sc.contact_points.append(pts)
        sc.body_density.append(density)
        sc.body_res.append({"cuboid": lambda: res, "cylinder": lambda: np.array([ares, rres]), "capsule": lambda: res}
                           .get(btype, lambda: None)())
        bname = bn.get("name", f"body{idx}")
        sc.body_names.append(bname)
        if bn.get("name") is not None:
            body_map[bname] = idx
        for child in node:
            if child.tag == "link":
                parse_link(child, idx)
        return idx

    for robot in root:
        if robot.tag == "robot":
            for rn in robot:
                if rn.tag == "link":
                    parse_link(rn, -1)

    # reduced dof offsets in DFS order
    off = 0
    for j in range(sc.nj):
        sc.qoff[j] = off
        off += sc.ndof[j]
    sc.ndof_r = off

    # actuators
    uoff = 0
    for an in root:
        if an.tag != "actuator":
            continue
        for m in an:
            if m.tag != "motor":
                continue
            jname = m.get("joint")
            if jname not in joint_map:
                raise SceneError("Actuator joint name error: " + str(jname))
            j = joint_map[jname]
            cr = _attr(m, default, "motor", "ctrl_range")
            rng = _vec(cr) if cr is not None else np.array([float(INT_MIN), float(INT_MAX)])
            mode = m.get("ctrl")
            nd = sc.ndof[j]
            act = {"joint": j, "uoff": uoff, "ndof": nd,
                   "cmin": np.full(nd, rng[0]), "cmax": np.full(nd, rng[1]),
                   "P": np.zeros(nd), "D": np.zeros(nd)}
            if mode == "force":
                act["mode"] = ACT_FORCE
            elif mode == "position":
                act["mode"] = ACT_POS
                P = _attr(m, default, "motor", "P")
                D = _attr(m, default, "motor", "D")
                act["P"] = np.full(nd, _f32(P) if P is not None else 0.0)
                act["D"] = np.full(nd, _f32(D) if D is not None else 0.0)
            else:
                continue  # the reference silently ignores unknown ctrl types
            sc.actuators.append(act)
            uoff += nd
    sc.ndof_u = uoff

    # tactile sensors
    for sn in root:
        if sn.tag != "sensor":
            continue
        for t in sn:
            if t.tag != "tactile":
                continue
            bname = t.get("body")
            if bname not in body_map:
                raise SceneError("Tactile body name error: " + str(bname))
            b = body_map[bname]
            coef = {}
            for key in ("kn", "kt", "mu", "damping"):
                v = _attr(t, default, "tactile", key)
                coef[key] = _f32(v) if v is not None else 0.0
            ttype = t.get("type")
            cands = [k for k in range(sc.nj) if sc.shape[k] != SH_NONE and k != b]
            if ttype == "abstract":
                # DH/Sensor/TactileSensorAbstract.cpp:8-124.  Markers, normals and axes are all mapped by
                # R_it v + p_it (the translation is applied to the direction vectors too, as the reference does);
                # for mesh / abstract bodies the frame is first composed with E_io (identity for the
                # principal-inertia abstract bodies supported here).
                if sc.btype[b] == "mesh":
                    raise SceneError("abstract tactile sensors on mesh bodies are not supported by the B200 path yet")
                p_it = _vec(t.get("pos"))
                R_it = quat2mat(_vec(t.get("quat")))
                txt = open(os.path.join(asset_dir, t.get("spec"))).read()
                N = int(txt.split()[0])
                fields = re.findall(r'"([^"]*)"', txt)
                if len(fields) < 5 * N:
                    raise SceneError("Tactile spec file is incomplete: " + str(t.get("spec")))
                pos, ipos, nrm, a0s, a1s = [], [], [], [], []
                for i in range(N):
                    f = fields[5 * i:5 * i + 5]
                    pos.append(R_it @ _vec(f[0]) + p_it)
                    ipos.append(_ivec(f[1]))
                    nrm.append(R_it @ _vec(f[2]) + p_it)
                    a0s.append(R_it @ _vec(f[3]) + p_it)
                    a1s.append(R_it @ _vec(f[4]) + p_it)
                sc.sensors.append(TactileSensor(
                    name=t.get("name", ""), body=b, pos=np.array(pos), axis0=np.array(a0s), axis1=np.array(a1s),
                    normal=np.array(nrm), image_pos=np.array(ipos, dtype=np.int64), candidates=cands, **coef))
                continue
            if ttype != "rect_array":
                raise SceneError(f"Tactile type {ttype} is not supported by the B200 path yet")
            p0 = _vec(t.get("rect_pos0"))
            p1 = _vec(t.get("rect_pos1"))
            ax0 = _vec(t.get("axis0"))
            ax1 = _vec(t.get("axis1"))
            res = _ivec(t.get("resolution"))
            l0 = float((p1 - p0) @ ax0)
            l1 = float((p1 - p0) @ ax1)
            if np.linalg.norm(p0 + l0 * ax0 + l1 * ax1 - p1) > 1e-5:
                raise SceneError("Tactile info for " + bname + " is incompatible")
            s0 = l0 / (res[0] - 1) * ax0
            s1 = l1 / (res[1] - 1) * ax1
            normal = np.cross(ax0, ax1)
            pos, ipos = [], []
            for i in range(res[0]):
                for j in range(res[1]):
                    pos.append(p0 + s0 * i + s1 * j)
                    ipos.append([i, j])
            M = len(pos)
            sc.sensors.append(TactileSensor(
                name=t.get("name", ""), body=b, pos=np.array(pos), axis0=np.tile(ax0, (M, 1)),
                axis1=np.tile(ax1, (M, 1)), normal=np.tile(normal, (M, 1)),
                image_pos=np.array(ipos, dtype=np.int64), candidates=cands, **coef))

    # contacts
    for cn in root:
        if cn.tag != "contact":
            continue
        for c in cn:
            coef = {}
            for key in ("kn", "kt", "mu", "damping"):
                v = _attr(c, default, c.tag, key)
                coef[key] = _f32(v) if v is not None else 0.0
            if c.tag == "ground_contact":
                bname = c.get("body")
                if bname not in body_map:
                    raise SceneError("Ground contact body name error: " + str(bname))
                sc.ground_contacts.append(dict(body=body_map[bname], **coef))
            elif c.tag == "general_primitive_contact":
                b1, b2 = c.get("general_body"), c.get("primitive_body")
                if b1 not in body_map:
                    raise SceneError("General contact body name error: " + str(b1))
                if b2 not in body_map:
                    raise SceneError("Primitive contact body name error: " + str(b2))
                if sc.shape[body_map[b2]] == SH_NONE:
                    raise SceneError("The second body in Contact should be primitive body.")
                sc.gp_contacts.append(dict(body1=body_map[b1], body2=body_map[b2], **coef))
            else:
                raise SceneError(f"contact type {c.tag} is not supported by the B200 path yet")

    # variables
    for vn in root:
        if vn.tag != "variable":
            continue
        for e in vn:
            if e.tag != "endeffector":
                continue
            jname = e.get("joint")
            if jname not in joint_map:
                raise SceneError("Endeffector joint name error: " + str(jname))
            # radius: rendering only (export_replay lists end-effectors with radius > 0), Simulation_Constructor.cpp:355-360
            rad = _attr(e, root.find("default"), "endeffector", "radius")
            sc.end_effectors.append({"joint": joint_map[jname], "pos": _vec(e.get("pos")),
                                     "name": e.get("name", ""), "radius": _f32(rad) if rad is not None else _f32("0.1")})
    for vn in root:
        if vn.tag == "virtual":
            for e in vn:
                sc.virtual_names.append(e.get("name", ""))
                # render-only objects: pose (pos, quat wxyz) kept for export_replay (Simulation_Constructor.cpp:380-420)
                pos = _vec(e.get("pos")) if e.get("pos") else np.zeros(3)
                quat = _vec(e.get("quat")) if (e.tag == "cuboid" and e.get("quat")) else np.array([1.0, 0.0, 0.0, 0.0])
                sc.virtual_pose.append(np.concatenate([pos, quat]))
    for s in sc.sensors:
        for k in s.candidates:
            if sc.shape[k] not in (SH_CUBOID, SH_CYLINDER, SH_SPHERE, SH_CAPSULE):
                raise SceneError("tactile candidates other than cuboids, cylinders, spheres and capsules are not supported by the B200 path yet")
    for gp in sc.gp_contacts:
        if sc.shape[gp["body2"]] not in (SH_CUBOID, SH_CYLINDER, SH_SPHERE, SH_CAPSULE):
            raise SceneError("primitive contact bodies other than cuboids, cylinders, spheres and capsules are not supported by the B200 path yet")
    if sc.integrator not in INTEGRATORS:
        raise SceneError("Integrator " + sc.integrator + " has not been implemented.")
    return sc


def joint_Q(jt: int, a0, a1, q):
    """Joint transform Q(q) (``Joint*.cpp update``)."""
    Q = np.eye(4)
    if jt == JT_REVOLUTE:
        c, s = math.cos(q[0]), math.sin(q[0])
        K = np.array([[0, -a0[2], a0[1]], [a0[2], 0, -a0[0]], [-a0[1], a0[0], 0]])
        Q[:3, :3] = np.eye(3) * c + s * K + (1 - c) * np.outer(a0, a0)
    elif jt == JT_PRISMATIC:
        Q[:3, 3] = a0 * q[0]
    elif jt == JT_PLANAR:
        Q[:3, 3] = a0 * q[0] + a1 * q[1]
    elif jt == JT_TRANSLATIONAL:
        Q[:3, 3] = q
    return Q


# ------------------------------------------------------------------ parameter updates (DH/Robot.cpp:571-650)
def primitive_inertia(shape: int, size, density: float) -> np.ndarray:
    """(Ixx, Iyy, Izz, m, m, m) of a primitive body: BodyCuboid.cpp:20-26, BodyCylinder.cpp:21-27, BodySphere.cpp:17-21,
    BodyCapsule.cpp:22-31 (the formulas compile_scene uses)."""
    inertia = np.zeros(6)
    if shape == SH_CUBOID:
        length = np.asarray(size, dtype=np.float64)
        mass = float(np.prod(length)) * density
        inertia[0] = mass / 12.0 * (length[1] * length[1] + length[2] * length[2])
        inertia[1] = mass / 12.0 * (length[0] * length[0] + length[2] * length[2])
        inertia[2] = mass / 12.0 * (length[0] * length[0] + length[1] * length[1])
    elif shape == SH_CYLINDER:
        radius, length = float(size[0]), float(size[1])
        mass = math.pi * radius * radius * length * density
        inertia[0] = mass * length * length / 12.0 + mass * radius * radius / 4.0
        inertia[1] = mass * length * length / 12.0 + mass * radius * radius / 4.0
        inertia[2] = mass * radius * radius / 2.0
    elif shape == SH_SPHERE:
        radius = float(size[0])
        mass = 4.0 / 3.0 * math.pi * radius * radius * radius * density
        inertia[:3] = 0.4 * mass * radius * radius
    elif shape == SH_CAPSULE:
        radius, length = float(size[0]), float(size[1])
        m_cy = density * length * math.pi * radius * radius
        m_hs = density * 2.0 / 3.0 * math.pi * radius * radius * radius
        mass = m_cy + 2.0 * m_hs
        inertia[0] = m_cy * (length * length / 12.0 + radius * radius / 4.0) + 2.0 * m_hs * (2.0 * radius * radius / 5.0 + length * length / 2.0 + 3.0 * length * radius / 8.0)
        inertia[1] = inertia[0]
        inertia[2] = m_cy * radius * radius / 2.0 + 2.0 * m_hs * 2.0 * radius * radius / 5.0
    else:
        raise SceneError("not a primitive body")
    inertia[3:] = mass
    return inertia


def _body_index(sc: Scene, body_name: str) -> int:
    """-1 for an unknown name: the reference walks its bodies and silently does nothing when none matches
    (DH/Robot.cpp:596-650)."""
    return sc.body_names.index(body_name) if body_name in sc.body_names else -1


def update_body_density(sc: Scene, body_name: str, density: float) -> None:
    """Simulation::update_body_density (DH/Robot.cpp:596-610): primitives recompute their mass matrix with the new
    density (Body*.cpp update_density); a mesh body re-runs process_mesh, i.e. mass and principal inertia scale with
    the density; abstract bodies ignore the call (Body.h:129)."""
    b = _body_index(sc, body_name)
    if b < 0:
        return
    density = float(density)
    if sc.shape[b] != SH_NONE:
        sc.inertia[b] = primitive_inertia(sc.shape[b], sc.size[b], density)
    elif b < len(sc.btype) and sc.btype[b] == "mesh":
        if b >= len(sc.body_density):
            raise SceneError("update_body_density on a mesh body needs a scene compiled from XML")
        sc.inertia[b] = sc.inertia[b] * (density / sc.body_density[b])
    if b < len(sc.body_density):
        sc.body_density[b] = density


def update_body_size(sc: Scene, body_name: str, body_size) -> None:
    """Simulation::update_body_size (DH/Robot.cpp:612-626): cuboid size = (lx, ly, lz) (BodyCuboid.cpp:289-294), cylinder
    size = (length, radius) (BodyCylinder.cpp:277-283): new sampled contact points and mass matrix; other bodies ignore
    the call (Body.h:130).  The density is the body's current one."""
    b = _body_index(sc, body_name)
    size = np.asarray(body_size, dtype=np.float64)
    if b < 0 or sc.shape[b] not in (SH_CUBOID, SH_CYLINDER):
        return
    if sc.shape[b] == SH_CUBOID:
        if size.shape[0] != 3:
            raise SceneError("update_body_size: a cuboid takes 3 values")
        mass_old, vol_old = sc.inertia[b][3], float(np.prod(sc.size[b]))
        new_size = size.copy()
    else:
        if size.shape[0] != 2:
            raise SceneError("update_body_size: a cylinder takes (length, radius)")
        mass_old, vol_old = sc.inertia[b][3], math.pi * sc.size[b][0] * sc.size[b][0] * sc.size[b][1]
        new_size = np.array([size[1], size[0], 0.0])
    density = sc.body_density[b] if b < len(sc.body_density) else mass_old / vol_old
    res = sc.body_res[b] if b < len(sc.body_res) else None
    if res is not None:
        sc.contact_points[b] = (_cuboid_points(new_size, res) if sc.shape[b] == SH_CUBOID
                                else _cylinder_points(new_size[0], new_size[1], int(res[0]), int(res[1])))
    elif len(sc.contact_points[b]):
        raise SceneError("update_body_size on a body whose sampled points are in use needs a scene compiled from XML")
    sc.size[b] = new_size
    sc.inertia[b] = primitive_inertia(sc.shape[b], new_size, density)


def update_joint_location(sc: Scene, joint_name: str, location) -> None:
    """Simulation::update_joint_location (DH/Joint/Joint.cpp:98-117): the joint's rest position in its parent (LOCAL
    frame) or in the world (WORLD frame); the children keep their transforms relative to this joint."""
    if joint_name not in sc.joint_names:
        return                             # as the reference: no joint of that name, nothing happens
    j = sc.joint_names.index(joint_name)
    p = np.asarray(location, dtype=np.float64)
    R0 = sc.joint_R0[j] if j < len(sc.joint_R0) else sc.E_pj0[j][:3, :3]
    frame = sc.joint_frame[j] if j < len(sc.joint_frame) else "LOCAL"
    par = sc.parent[j]
    if frame == "WORLD" and par >= 0:
        E_pj0 = sc.E_j0_0[par] @ SE(R0, p)
    else:
        E_pj0 = SE(R0, p)
    sc.E_pj0[j] = E_pj0
    if j < len(sc.E_j0_0):
        E_jp0 = Einv(E_pj0)
        sc.E_j0_0[j] = E_jp0 if par < 0 else E_jp0 @ sc.E_j0_0[par]
