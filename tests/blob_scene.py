"""Rebuilds a Scene (host tables) from a packed blob, so that tests on the GPU box -- where
/root/reference and its XML/mesh assets do not exist -- can drive the numpy oracle from the
committed golden fixtures.  Inverse of tactilesimulation_b200.layout.pack_scene."""
import numpy as np

from tactilesimulation_b200 import scene as S
from tactilesimulation_b200.layout import JI, GI, PI, AI, EI, SI, JD, CD, AD, ED, SD


def scene_from_blob(ibuf, dbuf):
    ib = np.asarray(ibuf).astype(np.int64)
    db = np.asarray(dbuf, dtype=np.float64)
    sc = S.Scene()
    nj, n, nu, nee, nm, ng, ngp, nact, nsens, max_iter, max_ls, npts = (int(x) for x in ib[2:14])
    sc.integrator = "BDF1"
    sc.h = float(db[0]); sc.gravity = db[1:4].copy(); sc.tol = float(db[4])
    sc.max_iter, sc.max_ls = max_iter, max_ls
    sc.E_g = np.eye(4); sc.E_g[:3, 2] = db[5:8]; sc.E_g[:3, 3] = db[8:11]
    sc.ndof_r, sc.ndof_u = n, nu
    P = db[ib[30]:ib[30] + 3 * npts].reshape(-1, 3)
    Mk = db[ib[31]:ib[31] + 3 * nm].reshape(-1, 3)
    sc.contact_points = [np.zeros((0, 3)) for _ in range(nj)]
    for j in range(nj):
        r = ib[ib[16] + j * JI: ib[16] + (j + 1) * JI]
        d = db[ib[24] + j * JD: ib[24] + (j + 1) * JD]
        sc.jtype.append(int(r[0])); sc.parent.append(int(r[1])); sc.qoff.append(int(r[2])); sc.ndof.append(int(r[3]))
        sc.shape.append(int(r[4]))
        E = np.eye(4); E[:3, :3] = d[0:9].reshape(3, 3); E[:3, 3] = d[9:12]
        sc.E_pj0.append(E)
        sc.axis0.append(d[12:15].copy()); sc.axis1.append(d[15:18].copy())
        sc.damping.append(float(d[18])); sc.lim_lo.append(float(d[19])); sc.lim_hi.append(float(d[20])); sc.lim_k.append(float(d[21]))
        E = np.eye(4); E[:3, :3] = d[22:31].reshape(3, 3); E[:3, 3] = d[31:34]
        sc.E_ji.append(E)
        sc.inertia.append(d[34:40].copy())
        sc.size.append(d[40:43] * 2.0 if int(r[4]) == S.SH_CUBOID else d[40:43].copy())
        sc.joint_names.append(f"joint{j}"); sc.body_names.append(f"body{j}")
    for i in range(ng):
        r = ib[ib[17] + i * GI: ib[17] + (i + 1) * GI]; d = db[ib[25] + i * CD: ib[25] + (i + 1) * CD]
        sc.contact_points[int(r[0])] = P[r[1]:r[1] + r[2]]
        sc.ground_contacts.append(dict(body=int(r[0]), kn=d[0], kt=d[1], mu=d[2], damping=d[3]))
        sc.has_ground = True
    for i in range(ngp):
        r = ib[ib[18] + i * PI: ib[18] + (i + 1) * PI]; d = db[ib[26] + i * CD: ib[26] + (i + 1) * CD]
        sc.contact_points[int(r[0])] = P[r[2]:r[2] + r[3]]
        sc.gp_contacts.append(dict(body1=int(r[0]), body2=int(r[1]), kn=d[0], kt=d[1], mu=d[2], damping=d[3]))
    for i in range(nact):
        r = ib[ib[19] + i * AI: ib[19] + (i + 1) * AI]; d = db[ib[27] + i * AD: ib[27] + (i + 1) * AD]
        nd = int(r[3])
        sc.actuators.append(dict(joint=int(r[0]), mode=int(r[1]), uoff=int(r[2]), ndof=nd, cmin=d[0:nd].copy(),
                                 cmax=d[3:3 + nd].copy(), P=d[6:6 + nd].copy(), D=d[9:9 + nd].copy()))
    for i in range(nee):
        r = ib[ib[20] + i * EI: ib[20] + (i + 1) * EI]; d = db[ib[28] + i * ED: ib[28] + (i + 1) * ED]
        sc.end_effectors.append(dict(joint=int(r[0]), pos=d[:3].copy(), name=""))
    for i in range(nsens):
        r = ib[ib[21] + i * SI: ib[21] + (i + 1) * SI]; d = db[ib[29] + i * SD: ib[29] + (i + 1) * SD]
        M = int(r[2])
        sc.sensors.append(S.TactileSensor(name=f"sensor{i}", body=int(r[0]), kn=d[0], kt=d[1], mu=d[2], damping=d[3],
                                          pos=Mk[r[1]:r[1] + M].copy(), axis0=np.tile(d[4:7], (M, 1)),
                                          axis1=np.tile(d[7:10], (M, 1)), normal=np.tile(d[10:13], (M, 1)),
                                          image_pos=np.zeros((M, 2), dtype=np.int64),
                                          candidates=[int(c) for c in r[4:4 + r[3]]]))
    return sc
