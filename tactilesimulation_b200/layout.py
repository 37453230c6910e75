"""Packs a :class:`Scene` into the (int32, float64) buffers described by
``csrc/scene_layout.h``; ``unpack_sizes`` reads the header back."""
from __future__ import annotations

import numpy as np

from . import scene as S

TS_MAGIC = 0x54533230
TS_VERSION = 5          # written; version-3 / 4 blobs (sensor records with 4 / 8 candidate slots) are still read
MAXB, MAXN, MAXCAND = 24, 16, 16   # largest kernel capacities (csrc/kernel_layout.h, variant 16): bodies, dofs, candidates
I_OFF_MARKER_IMAGE = 23           # header slot: int offset of the per-marker image positions (row, col), 0 = none
I_DOFF_MARKER_AXES = 22           # header slot: offset of the per-marker (axis0, axis1, normal) section, 0 = none
I_HEADER, D_HEADER = 32, 16
JI, GI, PI, AI, EI, SI = 8, 4, 4, 4, 2, 20
SI_BY_VERSION = {3: 8, 4: 12, 5: 20}
JD, CD, AD, ED, SD = 48, 4, 12, 4, 16


def pack_scene(sc: "S.Scene"):
    if sc.integrator not in S.INTEGRATORS:
        raise S.SceneError("Integrator " + sc.integrator + " has not been implemented.")
    if sc.nj > MAXB or sc.ndof_r > MAXN:
        raise S.SceneError(f"scene too large for this build: nj={sc.nj} (max {MAXB}), ndof_r={sc.ndof_r} (max {MAXN})")
    ints = [np.zeros(I_HEADER, dtype=np.int64)]
    dbls = [np.zeros(D_HEADER, dtype=np.float64)]
    hdr, dh = ints[0], dbls[0]

    def ioff():
        return sum(len(a) for a in ints)

    def doff():
        return sum(len(a) for a in dbls)

    # points pool: only bodies that actually use their sample points
    pts, pt_off = [], {}

    def points_of(b):
        if b not in pt_off:
            pt_off[b] = sum(len(p) for p in pts)
            pts.append(sc.contact_points[b])
        return pt_off[b], len(sc.contact_points[b])

    hdr[0:16] = [TS_MAGIC, TS_VERSION, sc.nj, sc.ndof_r, sc.ndof_u, len(sc.end_effectors), sc.n_markers,
                 len(sc.ground_contacts), len(sc.gp_contacts), len(sc.actuators), len(sc.sensors),
                 sc.max_iter, sc.max_ls, 0, S.INTEGRATORS[sc.integrator], 0]
    dh[0] = sc.h
    dh[1:4] = sc.gravity
    dh[4] = sc.tol
    dh[5:8] = sc.E_g[:3, 2]
    dh[8:11] = sc.E_g[:3, 3]

    hdr[16], hdr[24] = ioff(), doff()
    ji = np.zeros((sc.nj, JI), dtype=np.int64)
    jd = np.zeros((sc.nj, JD))
    for j in range(sc.nj):
        ji[j, :5] = [sc.jtype[j], sc.parent[j], sc.qoff[j], sc.ndof[j], sc.shape[j]]
        jd[j, 0:9] = sc.E_pj0[j][:3, :3].reshape(-1)
        jd[j, 9:12] = sc.E_pj0[j][:3, 3]
        jd[j, 12:15] = sc.axis0[j]
        jd[j, 15:18] = sc.axis1[j]
        jd[j, 18:22] = [sc.damping[j], sc.lim_lo[j], sc.lim_hi[j], sc.lim_k[j]]
        jd[j, 22:31] = sc.E_ji[j][:3, :3].reshape(-1)
        jd[j, 31:34] = sc.E_ji[j][:3, 3]
        jd[j, 34:40] = sc.inertia[j]
        if sc.shape[j] == S.SH_CUBOID:
            jd[j, 40:43] = sc.size[j] / 2.0
        else:
            jd[j, 40:43] = sc.size[j]
    ints.append(ji.reshape(-1))
    dbls.append(jd.reshape(-1))

    hdr[17], hdr[25] = ioff(), doff()
    gi = np.zeros((len(sc.ground_contacts), GI), dtype=np.int64)
    gd = np.zeros((len(sc.ground_contacts), CD))
    for i, g in enumerate(sc.ground_contacts):
        o, c = points_of(g["body"])
        gi[i, :3] = [g["body"], o, c]
        gd[i] = [g["kn"], g["kt"], g["mu"], g["damping"]]
    ints.append(gi.reshape(-1))
    dbls.append(gd.reshape(-1))

    hdr[18], hdr[26] = ioff(), doff()
    pi = np.zeros((len(sc.gp_contacts), PI), dtype=np.int64)
    pd = np.zeros((len(sc.gp_contacts), CD))
    for i, f in enumerate(sc.gp_contacts):
        o, c = points_of(f["body1"])
        pi[i] = [f["body1"], f["body2"], o, c]
        pd[i] = [f["kn"], f["kt"], f["mu"], f["damping"]]
    ints.append(pi.reshape(-1))
    dbls.append(pd.reshape(-1))

    hdr[19], hdr[27] = ioff(), doff()
    ai = np.zeros((len(sc.actuators), AI), dtype=np.int64)
    ad = np.zeros((len(sc.actuators), AD))
    for i, a in enumerate(sc.actuators):
        ai[i] = [a["joint"], a["mode"], a["uoff"], a["ndof"]]
        nd = a["ndof"]
        ad[i, 0:nd] = a["cmin"]
        ad[i, 3:3 + nd] = a["cmax"]
        ad[i, 6:6 + nd] = a["P"]
        ad[i, 9:9 + nd] = a["D"]
    ints.append(ai.reshape(-1))
    dbls.append(ad.reshape(-1))

    hdr[20], hdr[28] = ioff(), doff()
    ei = np.zeros((len(sc.end_effectors), EI), dtype=np.int64)
    ed = np.zeros((len(sc.end_effectors), ED))
    for i, e in enumerate(sc.end_effectors):
        ei[i, 0] = e["joint"]
        ed[i, :3] = e["pos"]
    ints.append(ei.reshape(-1))
    dbls.append(ed.reshape(-1))

    hdr[21], hdr[29] = ioff(), doff()
    si = np.zeros((len(sc.sensors), SI), dtype=np.int64)
    sd = np.zeros((len(sc.sensors), SD))
    moff = 0
    markers = []
    for i, s in enumerate(sc.sensors):
        if len(s.candidates) > MAXCAND:
            raise S.SceneError(f"more than {MAXCAND} tactile candidate bodies are not supported yet")
        si[i, :4] = [s.body, moff, len(s.pos), len(s.candidates)]
        si[i, 4:4 + len(s.candidates)] = s.candidates
        sd[i, :4] = [s.kn, s.kt, s.mu, s.damping]
        sd[i, 4:7] = s.axis0[0]
        sd[i, 7:10] = s.axis1[0]
        sd[i, 10:13] = s.normal[0]
        markers.append(s.pos)
        moff += len(s.pos)
    ints.append(si.reshape(-1))
    dbls.append(sd.reshape(-1))
    if sc.sensors:
        # image positions of the markers (host-side only: get_tactile_image_pos / get_tactile_flow_images)
        hdr[I_OFF_MARKER_IMAGE] = ioff()
        ints.append(np.concatenate([np.asarray(s.image_pos, dtype=np.int64).reshape(-1, 2) for s in sc.sensors], axis=0).reshape(-1))

    hdr[30] = doff()
    P = np.concatenate(pts, axis=0) if pts else np.zeros((0, 3))
    hdr[13] = len(P)
    dbls.append(P.reshape(-1))
    hdr[31] = doff()
    Mk = np.concatenate(markers, axis=0) if markers else np.zeros((0, 3))
    dbls.append(Mk.reshape(-1))
    # per-marker (axis0, axis1, normal): abstract sensors carry them per marker (DH/Sensor/TactileSensorAbstract.cpp)
    if sc.sensors:
        hdr[I_DOFF_MARKER_AXES] = doff()
        ax = np.concatenate([np.concatenate([s.axis0, s.axis1, s.normal], axis=1) for s in sc.sensors], axis=0)
        dbls.append(ax.reshape(-1))
    ibuf = np.concatenate(ints).astype(np.int32)
    dbuf = np.concatenate(dbls).astype(np.float64)
    return ibuf, dbuf


def unpack_sizes(ibuf):
    if int(ibuf[0]) != TS_MAGIC or int(ibuf[1]) not in SI_BY_VERSION:
        raise S.SceneError("not a tactilesimulation_b200 scene blob (magic/version mismatch)")
    nj, n, nu, nee, nm = (int(ibuf[i]) for i in (2, 3, 4, 5, 6))
    return dict(nj=nj, ndof_r=n, ndof_m=6 * nj, ndof_u=nu, ndof_var=3 * nee, n_markers=nm, ndof_tactile=3 * nm)


# ------------------------------------------------------------------ inverse of pack_scene
def scene_from_blob(ibuf, dbuf):
    """Rebuilds the host tables of a Scene from a packed blob (inverse of pack_scene), so that a scene
    can be shipped as two arrays (e.g. the golden fixtures) where the XML and meshes do not exist."""
    ib = np.asarray(ibuf).astype(np.int64)
    db = np.asarray(dbuf, dtype=np.float64)
    sc = S.Scene()
    nj, n, nu, nee, nm, ng, ngp, nact, nsens, max_iter, max_ls, npts = (int(x) for x in ib[2:14])
    sc.integrator = {v: k for k, v in S.INTEGRATORS.items()}[int(ib[14])]
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
    si = SI_BY_VERSION[int(ib[1])]
    for i in range(nsens):
        r = ib[ib[21] + i * si: ib[21] + (i + 1) * si]; d = db[ib[29] + i * SD: ib[29] + (i + 1) * SD]
        M = int(r[2])
        if ib[I_DOFF_MARKER_AXES] > 0:
            A = db[ib[I_DOFF_MARKER_AXES] + 9 * r[1]: ib[I_DOFF_MARKER_AXES] + 9 * (r[1] + M)].reshape(M, 9)
            a0, a1, nr = A[:, 0:3].copy(), A[:, 3:6].copy(), A[:, 6:9].copy()
        else:
            a0, a1, nr = np.tile(d[4:7], (M, 1)), np.tile(d[7:10], (M, 1)), np.tile(d[10:13], (M, 1))
        if ib[I_OFF_MARKER_IMAGE] > 0:
            ipos = ib[ib[I_OFF_MARKER_IMAGE] + 2 * r[1]: ib[I_OFF_MARKER_IMAGE] + 2 * (r[1] + M)].reshape(M, 2).copy()
        else:
            ipos = np.zeros((M, 2), dtype=np.int64)
        sc.sensors.append(S.TactileSensor(name=f"sensor{i}", body=int(r[0]), kn=d[0], kt=d[1], mu=d[2], damping=d[3],
                                          pos=Mk[r[1]:r[1] + M].copy(), axis0=a0, axis1=a1, normal=nr,
                                          image_pos=ipos,
                                          candidates=[int(c) for c in r[4:4 + r[3]]]))
    return sc
