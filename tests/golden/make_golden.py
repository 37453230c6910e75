#!/usr/bin/env python
"""Generates the golden fixtures in this directory by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/redmax_py + redmax_probe, built by oracle/build_ref.sh from /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Fixtures (npz): compiled scene blobs + seeded inputs + reference outputs
    q, qdot, var, tactile per step; active contact-point index sets per Force and
    marker->body ids per step; df_dq0 / df_dqdot0 / df_du of Simulation::backward();
    df_du of a StepSimFunction-style chain of backward_steps(5).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import redmax_py  # noqa: E402
import redmax_probe  # noqa: E402
from tactilesimulation_b200.scene import compile_scene  # noqa: E402

ASSETS = os.path.join(ROOT, "oracle", "_ref", "assets", "pusher")


def make_inputs(sim, T, seed, push=True):
    """SURVEY.md §8(d) config 2/3 inputs: q0[1]=-0.001, q0[4]~U(-.02,.02); u=tanh(N(0,1)) on
    dims 0-2, dims 3-4 resampled every 10 steps (p=.5 zero else U(-1,1)), dim 5 = 0."""
    rng = np.random.default_rng(seed)
    q0 = np.array(sim.get_q_init())
    q0[1] = -0.001
    q0[4] = rng.uniform(-0.02, 0.02)
    u = np.zeros((T, sim.ndof_u))
    ext = np.zeros(2)
    for t in range(T):
        u[t, :3] = np.tanh(rng.normal(size=3))
        if push:
            u[t, 0] = abs(u[t, 0])          # keep pushing so that contact happens
        if t % 10 == 0:
            ext = rng.uniform(-1, 1, 2) if rng.uniform() >= 0.5 else np.zeros(2)
        u[t, 3:5] = ext
    return q0, u


def pad_ids(lists, width):
    out = -np.ones((len(lists), width), dtype=np.int32)
    for i, l in enumerate(lists):
        out[i, :len(l)] = l
    return out


def episodic_case(xml, T, seed, push=True):
    sim = redmax_py.Simulation(xml)
    probe = redmax_probe.ProbeSimulation(xml)
    sc = compile_scene(xml)
    q0, u = make_inputs(sim, T, seed, push)
    n, nv, nt = sim.ndof_r, sim.ndof_var, sim.ndof_tactile
    for s in (sim, probe):
        s.set_state_init(q0, np.zeros(n))
        s.reset(True)
    q, qd, var, tac = [], [], [], []
    ground, gp, mb, newton = [], [], [], []
    probe.newton_counts()                        # (process-wide counters: drop what earlier cases left)
    for t in range(T):
        for s in (sim, probe):
            s.set_u(u[t])
            s.forward(1)
        assert np.array_equal(sim.get_q(), probe.get_q())
        newton.append(probe.newton_counts())     # (Newton iterations, line-search evaluations) of this step
        q.append(sim.get_q().copy())
        qd.append(sim.get_qdot().copy())
        var.append(sim.get_variables().copy())
        tac.append(sim.get_tactile_force_vector().copy())
        cs = probe.contact_sets()
        ground.append(cs["ground"][0])
        gp.append(cs["gp"][0])
        mb.append(cs["marker_body"][0])
    rng = np.random.default_rng(1000 + seed)
    df_dq = rng.normal(size=(T, n))
    df_dvar = rng.normal(size=(T, nv))
    df_dtac = 1e-3 * rng.normal(size=(T, nt))
    bi = sim.backward_info
    bi.set_flags(True, True, False, True)
    bi.df_dq = df_dq.reshape(-1)
    bi.df_dvar = df_dvar.reshape(-1)
    bi.df_dtactile = df_dtac.reshape(-1)
    bi.df_dq0 = np.zeros(n)
    bi.df_dqdot0 = np.zeros(n)
    bi.df_du = np.zeros(sim.ndof_u * T)
    sim.backward()
    br = sim.backward_results
    ib, db = sc.pack()
    return dict(ibuf=ib, dbuf=db, q0=q0, qd0=np.zeros(n), u=u, q=np.array(q), qd=np.array(qd), var=np.array(var),
                tactile=np.array(tac), ground_ids=pad_ids(ground, len(sc.contact_points[sc.ground_contacts[0]["body"]])),
                gp_ids=pad_ids(gp, len(sc.contact_points[sc.gp_contacts[0]["body1"]])),
                marker_body=np.array(mb, dtype=np.int32), df_dq=df_dq, df_dvar=df_dvar, df_dtactile=df_dtac,
                df_dq0=np.array(br.df_dq0), df_dqdot0=np.array(br.df_dqdot0),
                df_du=np.array(br.df_du).reshape(T, sim.ndof_u), newton=np.array(newton, dtype=np.int32))


def stepsim_case(xml, nsteps, frame_skip, seed):
    """The StepSimFunction call pattern (R/envs/redmax_torch_functions.py:112-174):
    forward(frame_skip, save_last_frame_var_only=True) per gym step, then a reverse chain of
    backward_steps(frame_skip) with cotangents on the last sub-step only."""
    sim = redmax_py.Simulation(xml)
    sc = compile_scene(xml)
    q0, u = make_inputs(sim, nsteps, seed)
    n, nv, nt, nu = sim.ndof_r, sim.ndof_var, sim.ndof_tactile, sim.ndof_u
    sim.set_state_init(q0, np.zeros(n))
    sim.reset(True)
    q, var, tac = [], [], []
    for t in range(nsteps):
        sim.set_u(u[t])
        sim.forward(frame_skip, False, False, True)
        q.append(sim.get_q().copy())
        var.append(sim.get_variables().copy())
        tac.append(sim.get_tactile_force_vector().copy())
    rng = np.random.default_rng(2000 + seed)
    df_dq = rng.normal(size=(nsteps, n))
    df_dvar = rng.normal(size=(nsteps, nv))
    df_dtac = 1e-3 * rng.normal(size=(nsteps, nt))
    df_du = np.zeros((nsteps, frame_skip, nu))
    for t in range(nsteps - 1, -1, -1):
        bi = sim.backward_info
        bi.set_flags(False, False, False, True)
        a = np.zeros(n * frame_skip); a[-n:] = df_dq[t]
        b = np.zeros(nv * frame_skip); b[-nv:] = df_dvar[t]
        c = np.zeros(nt * frame_skip); c[-nt:] = df_dtac[t]
        bi.df_dq, bi.df_dvar, bi.df_dtactile = a, b, c
        bi.df_du = np.zeros(nu * frame_skip)
        sim.backward_steps(frame_skip)
        df_du[t] = np.array(sim.backward_results.df_du).reshape(frame_skip, nu)
    ib, db = sc.pack()
    return dict(ibuf=ib, dbuf=db, q0=q0, qd0=np.zeros(n), u=u, frame_skip=frame_skip, q=np.array(q), var=np.array(var),
                tactile=np.array(tac), df_dq=df_dq, df_dvar=df_dvar, df_dtactile=df_dtac, df_du=df_du)


def multi_case(xml, q0, u, seed):
    """Scenes with several contact forces / sensors (DClaw): same content as episodic_case, with the contact
    index sets as per-force lists padded into [T, forces, max points] and marker->body ids over all sensors."""
    sim = redmax_py.Simulation(xml)
    probe = redmax_probe.ProbeSimulation(xml)
    sc = compile_scene(xml)
    T = u.shape[0]
    n, nv, nt = sim.ndof_r, sim.ndof_var, sim.ndof_tactile
    for s in (sim, probe):
        s.set_state_init(q0, np.zeros(n))
        s.reset(True)
    q, qd, var, tac, ground, gp, mb, newton = [], [], [], [], [], [], [], []
    probe.newton_counts()                        # (process-wide counters: drop what earlier cases left)
    for t in range(T):
        for s in (sim, probe):
            s.set_u(u[t])
            s.forward(1)
        assert np.array_equal(sim.get_q(), probe.get_q())
        newton.append(probe.newton_counts())
        q.append(sim.get_q().copy())
        qd.append(sim.get_qdot().copy())
        var.append(sim.get_variables().copy())
        tac.append(sim.get_tactile_force_vector().copy())
        cs = probe.contact_sets()
        ground.append([list(x) for x in cs["ground"]])
        gp.append([list(x) for x in cs["gp"]])
        mb.append(np.concatenate([np.asarray(x, dtype=np.int32) for x in cs["marker_body"]]))
    rng = np.random.default_rng(1000 + seed)
    df_dq = rng.normal(size=(T, n))
    df_dvar = rng.normal(size=(T, nv))
    df_dtac = 1e-3 * rng.normal(size=(T, nt))
    bi = sim.backward_info
    bi.set_flags(True, True, False, True)
    bi.df_dq, bi.df_dvar, bi.df_dtactile = df_dq.reshape(-1), df_dvar.reshape(-1), df_dtac.reshape(-1)
    bi.df_dq0, bi.df_dqdot0, bi.df_du = np.zeros(n), np.zeros(n), np.zeros(sim.ndof_u * T)
    sim.backward()
    br = sim.backward_results

    def pad3(lists, nforce, width):
        out = -np.ones((T, max(nforce, 1), max(width, 1)), dtype=np.int32)
        for t, per_force in enumerate(lists):
            for f, ids in enumerate(per_force):
                out[t, f, :len(ids)] = ids
        return out

    ib, db = sc.pack()
    gw = max([len(sc.contact_points[g["body"]]) for g in sc.ground_contacts] + [0])
    pw = max([len(sc.contact_points[f["body1"]]) for f in sc.gp_contacts] + [0])
    return dict(ibuf=ib, dbuf=db, q0=q0, qd0=np.zeros(n), u=u, q=np.array(q), qd=np.array(qd), var=np.array(var),
                tactile=np.array(tac), ground_ids_f=pad3(ground, len(sc.ground_contacts), gw),
                gp_ids_f=pad3(gp, len(sc.gp_contacts), pw), marker_body=np.array(mb, dtype=np.int32),
                df_dq=df_dq, df_dvar=df_dvar, df_dtactile=df_dtac, df_dq0=np.array(br.df_dq0),
                df_dqdot0=np.array(br.df_dqdot0), df_du=np.array(br.df_du).reshape(T, sim.ndof_u),
                newton=np.array(newton, dtype=np.int32))


def dclaw_case(T, seed, xml_name="dclaw_torque_control.xml"):
    """DClaw rotate-cap (R/envs/assets/dclaw_rotate/dclaw_torque_control.xml, BASELINE configs[3]): initial pose of
    R/envs/dclaw_rotate_env.py:76-77, actions U(-1,1)^9 with the middle joints biased to close on the cap."""
    xml = os.path.join(ROOT, "oracle", "_ref", "assets", "dclaw_rotate", xml_name)
    q0 = np.zeros(10)
    q0[[1, 4, 7]] = -0.5
    q0[[2, 5, 8]] = 0.8
    rng = np.random.default_rng(seed)
    u = rng.uniform(-1, 1, (T, 9))
    u[:, 1::3] = 0.6 + 0.4 * u[:, 1::3]
    return multi_case(xml, q0, u, seed)


def insertion_case(T, seed, xml_name="tactile_insertion.xml"):
    """TactileInsertion (R/envs/assets/tactile_insertion/tactile_insertion.xml, BASELINE configs[4]): position-controlled
    gripper base (translational + revolute), force-controlled fingers, free3d-euler box in a hole of four cuboids,
    two 13x10 pads.  Grasp as in R/envs/tactile_insertion_env.py:126-164 (fingers ramped closed), then the base is
    driven sideways / rotated / down so that the box meets the hole walls."""
    xml = os.path.join(ROOT, "oracle", "_ref", "assets", "tactile_insertion", xml_name)
    q0 = np.zeros(12)
    q0[2] = 0.2
    q0[4] = q0[5] = -0.03
    rng = np.random.default_rng(seed)
    u = np.zeros((T, 6))
    for t in range(T):
        g = min(1.0, (t + 1) / 15.0)
        m = max(0.0, (t - 25) / float(max(T - 25, 1)))
        u[t] = [0.004 * m, -0.003 * m, 0.2 - 0.002 * m, 0.15 * m, g, g]
    u[:, :4] += 1e-4 * rng.normal(size=(T, 4))
    return multi_case(xml, q0, u, seed)


def stable_grasp_case(T, seed):
    """StableGrasp (R/envs/assets/stable_grasp/stable_grasp.xml): position-controlled gripper closing on a bar of eleven
    rigidly connected boxes (free3d-euler), lifting it; 44 general-primitive and 11 ground contacts, two pads with 15
    candidate bodies each.  Stages 2-3 of R/envs/stable_grasp_env.py:197-246, shortened."""
    xml = os.path.join(ROOT, "oracle", "_ref", "assets", "stable_grasp", "stable_grasp.xml")
    q0 = np.zeros(12)
    q0[2] = 0.2029862
    q0[4] = q0[5] = -0.03
    rng = np.random.default_rng(seed)
    gp = 0.01                                  # grasp position along the bar: off-centre, the bar tilts when lifted
    q0[1] = gp
    stages = [np.array([0.0, gp, 0.2029862, 0.0, -0.03, -0.03]), np.array([0.0, gp, 0.2029862, 0.0, -0.008, -0.008]),
              np.array([0.0, gp, 0.2029862, 0.0, -0.008, -0.008]), np.array([0.0, gp, 0.2329862, 0.0, -0.008, -0.008])]
    steps = [20, 8, T - 28]
    u = []
    for st in range(3):
        for i in range(steps[st]):
            u.append((stages[st + 1] - stages[st]) / steps[st] * (i + 1) + stages[st])
    u = np.array(u)
    u[:, 2] += 0.003
    u[:, :4] += 1e-5 * rng.normal(size=(T, 4))
    return multi_case(xml, q0, u, seed)


def rolling_ball_case():
    """examples/RollingBallExp/test_sim_speed.py (BASELINE configs[0]): the unmodified tactile_pad.xml scene (BDF2 with
    an SDIRK2 start-up step, free3d-exp ball, sphere SDF, 2168 sampled pad points) under the script's action
    schedule, 350 steps.  The trajectory and the tactile field every 5 steps come from the 40x40-marker variant
    (identical dynamics: sensors do not act on the bodies); of the real 200x200 field four frames are kept as
    sparse (index, value) lists."""
    adir = os.path.join(ROOT, "oracle", "_ref", "assets", "tactile_pad")
    xml40, xml200 = os.path.join(adir, "tactile_pad_40x40.xml"), os.path.join(adir, "tactile_pad.xml")
    acts = [np.array([0., 0., 0.2]), np.array([0.1, 0., 0.2]), np.array([-0.2, 0., 0.2]), np.array([0., 0.1, 0.2]),
            np.array([0., -0.2, 0.2])]
    steps = [0, 100, 150, 200, 250, 350]
    u = np.zeros((350, 3))
    for i in range(5):
        u[steps[i]:steps[i + 1]] = acts[i]
    T = len(u)
    sim, probe, big = redmax_py.Simulation(xml40), redmax_probe.ProbeSimulation(xml40), redmax_py.Simulation(xml200)
    sc = compile_scene(xml40)
    for s in (sim, probe, big):
        s.reset(False)
    frames200 = (75, 150, 250, 349)
    q, qd, tac, ground, gp, mb, big_idx, big_val = [], [], [], [], [], [], [], []
    for t in range(T):
        for s in (sim, probe, big):
            s.set_u(u[t])
            s.forward(1)
        assert np.array_equal(sim.get_q(), probe.get_q()) and np.array_equal(sim.get_q(), big.get_q())
        q.append(sim.get_q().copy())
        qd.append(sim.get_qdot().copy())
        cs = probe.contact_sets()
        ground.append(cs["ground"][0])
        gp.append(cs["gp"][0])
        if t % 5 == 0:
            tac.append(sim.get_tactile_force_vector().copy())
            mb.append(np.asarray(cs["marker_body"][0], dtype=np.int32))
        if t in frames200:
            f = big.get_tactile_force_vector().copy()
            nz = np.nonzero(f)[0]
            big_idx.append(nz.astype(np.int32))
            big_val.append(f[nz])
    ib, db = sc.pack()
    width = max(len(x) for x in big_idx)
    bi = -np.ones((len(frames200), width), dtype=np.int32)
    bv = np.zeros((len(frames200), width))
    for k in range(len(frames200)):
        bi[k, :len(big_idx[k])] = big_idx[k]
        bv[k, :len(big_val[k])] = big_val[k]
    return dict(ibuf=ib, dbuf=db, q0=np.zeros(sim.ndof_r), qd0=np.zeros(sim.ndof_r), u=u, q=np.array(q), qd=np.array(qd),
                tactile=np.array(tac), tactile_every=5, marker_body=np.array(mb, dtype=np.int32),
                ground_ids=pad_ids(ground, 1), gp_ids=pad_ids(gp, max(len(x) for x in gp)),
                frames200=np.array(frames200), tactile200_idx=bi, tactile200_val=bv)


# a plate on a free2d joint in a tilted plane carrying a revolute arm (our own synthetic scene): ground contact at the
# arm, a 4x3 pad on the plate pressed by the arm, one end-effector, four force motors
FREE2D_TEMPLATE = '''<redmax model="free2d-plate">
    <option integrator="BDF1" timestep="5e-3" unit="m-kg" gravity="0. 0. -9.8"/>
    <ground pos="0 0 0" normal="0 0 1"/>
    <default>
        <ground_contact kn="2e3" kt="10" mu="0.8" damping="3"/>
        <tactile kn="50" kt="4" mu="1.0" damping="2"/>
        <motor ctrl="force" ctrl_range="-0.5 0.5"/>
    </default>
    <robot>
        <link name="plate">
            <joint name="plate_joint" type="free2d" pos="0 0 0.06" quat="0.9887711 0.1494381 0 0" damping="0.5"/>
            <body name="plate_body" type="cuboid" size="0.08 0.05 0.02" pos="0 0 0" quat="1 0 0 0" density="500." general_contact_resolution="2 2 2"/>
            <link name="arm">
                <joint name="arm_joint" type="revolute" axis="0 0 1" pos="0.04 0 0.01" quat="1 0 0 0" damping="0.002"/>
                <body name="arm_body" type="cuboid" size="0.06 0.02 0.02" pos="0.03 0 0" quat="1 0 0 0" density="500." general_contact_resolution="3 2 2"/>
            </link>
        </link>
    </robot>
    <actuator>
        <motor joint="plate_joint" ctrl="force"/>
        <motor joint="arm_joint" ctrl="force"/>
    </actuator>
    <sensor>
        <tactile body="plate_body" name="pad" type="rect_array" rect_pos0="0.03 -0.008 0.01" rect_pos1="0.039 0.008 0.01" axis0="1 0 0" axis1="0 1 0" resolution="4 3"/>
    </sensor>
    <contact>
        <ground_contact body="arm_body"/>
    </contact>
    <variable>
        <endeffector joint="arm_joint" pos="0.06 0 0" name="tip"/>
    </variable>
</redmax>
'''


def free2d_case(T, seed):
    """free2d joint (DH/Joint/JointFree2D.cpp) in a tilted plane, BDF1, forward + backward()."""
    d = os.path.join(ROOT, "oracle", "_ref", "assets", "synthetic")
    os.makedirs(d, exist_ok=True)
    xml = os.path.join(d, "free2d_plate.xml")
    open(xml, "w").write(FREE2D_TEMPLATE)
    rng = np.random.default_rng(seed)
    return multi_case(xml, np.array([0.01, -0.02, 0.3, 0.4]), rng.uniform(-1, 1, (T, 4)), seed)


# a pad pressed on a capsule lying on the ground and rolling it (our own synthetic scene: no reference asset has a capsule)
CAPSULE_TEMPLATE = '''<redmax model="capsule-press">
    <option integrator="BDF1" timestep="5e-3" unit="m-kg" gravity="0. 0. -9.8"/>
    <ground pos="0 0 0" normal="0 0 1"/>
    <default>
        <general_primitive_contact kn="20" kt="1" mu="1.0" damping="1"/>
        <tactile kn="1" kt="0.05" mu="2." damping="0.003"/>
    </default>
    <robot>
        <link name="pad">
            <joint name="pad_joint" type="translational" pos="0 0 0.045" quat="1 0 0 0" damping="1"/>
            <body name="pad_body" type="cuboid" size="0.05 0.05 0.01" pos="0 0 0" quat="1 0 0 0" density="1000." general_contact_resolution="8 8 2"/>
        </link>
    </robot>
    <robot>
        <link name="object">
            <joint name="object_joint" type="free3d-euler" pos="0.004 0. 0.015" quat="1 0 0 0"/>
            <body name="object" type="capsule" pos="0 0 0" radius="0.015" length="0.012" quat="0.7071068 0 0.7071068 0" density="300." general_contact_resolution="5 8"/>
        </link>
    </robot>
    <actuator>
        <motor joint="pad_joint" ctrl="force" ctrl_range="-1 1"/>
    </actuator>
    <sensor>
        <tactile body="pad_body" name="pad" type="rect_array" rect_pos0="-0.025 0.025 -0.005" rect_pos1="0.025 -0.025 -0.005" axis0="0 -1 0" axis1="1 0 0" resolution="12 12"/>
    </sensor>
    <contact>
        <ground_contact body="object" kn="5e3" kt="1" mu="0.8" damping="0.03"/>
        <general_primitive_contact general_body="pad_body" primitive_body="object"/>
    </contact>
</redmax>
'''


def capsule_case(T, seed):
    """capsule SDF (DH/Body/BodyCapsule.cpp) as contact primitive, tactile candidate and -- through its sampled points --
    ground contact body; free3d-euler joint; BDF1 forward + backward()."""
    d = os.path.join(ROOT, "oracle", "_ref", "assets", "synthetic")
    os.makedirs(d, exist_ok=True)
    xml = os.path.join(d, "capsule_press.xml")
    open(xml, "w").write(CAPSULE_TEMPLATE)
    sim, probe, sc = redmax_py.Simulation(xml), redmax_probe.ProbeSimulation(xml), compile_scene(xml)
    n, nt, nu = sim.ndof_r, sim.ndof_tactile, sim.ndof_u
    u = np.zeros((T, nu))
    u[:, 2] = 0.1
    u[T // 2:, 1] = 0.15
    u[T // 2:, 0] = 0.05
    q0 = np.zeros(n)
    q0[2] = -0.008
    for s_ in (sim, probe):
        s_.set_state_init(q0, np.zeros(n))
        s_.reset(True)
    q, qd, tac, ground, gp, mb = [], [], [], [], [], []
    for t in range(T):
        for s_ in (sim, probe):
            s_.set_u(u[t])
            s_.forward(1)
        q.append(sim.get_q().copy())
        qd.append(sim.get_qdot().copy())
        tac.append(sim.get_tactile_force_vector().copy())
        cs = probe.contact_sets()
        ground.append(cs["ground"][0])
        gp.append(cs["gp"][0])
        mb.append(np.asarray(cs["marker_body"][0], dtype=np.int32))
    rng = np.random.default_rng(1000 + seed)
    df_dq = rng.normal(size=(T, n))
    df_dtac = 1e-3 * rng.normal(size=(T, nt))
    bi = sim.backward_info
    bi.set_flags(True, True, False, True)
    bi.df_dq, bi.df_dtactile = df_dq.reshape(-1), df_dtac.reshape(-1)
    bi.df_dq0, bi.df_dqdot0, bi.df_du = np.zeros(n), np.zeros(n), np.zeros(nu * T)
    sim.backward()
    br = sim.backward_results
    ib, db = sc.pack()
    return dict(ibuf=ib, dbuf=db, q0=q0, qd0=np.zeros(n), u=u, q=np.array(q), qd=np.array(qd), tactile=np.array(tac),
                ground_ids=pad_ids(ground, len(sc.contact_points[1])), gp_ids=pad_ids(gp, len(sc.contact_points[0])),
                marker_body=np.array(mb, dtype=np.int32), cot_seed=1000 + seed,
                df_dq0=np.array(br.df_dq0), df_dqdot0=np.array(br.df_dqdot0), df_du=np.array(br.df_du).reshape(T, nu))


def randomized_case():
    """The domain-randomisation calls of the reference's envs before reset() -- update_joint_damping, update_body_size,
    update_endeffector_position, update_joint_location (R/envs/dclaw_rotate_env.py:173-178) and update_body_density
    (R/envs/stable_grasp_env.py:122) -- then a short rollout.  The fixture keeps the ORIGINAL scene blobs (+ names);
    the tests apply the same updates through tactilesimulation_b200.scene.update_* and compare the rollouts."""
    out = {}
    # ---- DClaw: cap radius / damping / location
    xml = os.path.join(ROOT, "oracle", "_ref", "assets", "dclaw_rotate", "dclaw_torque_control.xml")
    sim = redmax_py.Simulation(xml)
    sc = compile_scene(xml)
    upd = dict(damping=0.0015, size=np.array([0.03, 0.052]), ee=np.array([0.052, 0.0, 0.0]), loc=np.array([0.003, -0.004, 0.075]))
    sim.update_joint_damping("cap", upd["damping"])
    sim.update_body_size("cap", upd["size"])
    sim.update_endeffector_position("cap", upd["ee"])
    sim.update_joint_location("cap", upd["loc"])
    q0 = np.zeros(10)
    q0[[1, 4, 7]] = -0.5
    q0[[2, 5, 8]] = 0.8
    rng = np.random.default_rng(7)
    T = 30
    u = rng.uniform(-1, 1, (T, 9))
    u[:, 1::3] = 0.6 + 0.4 * u[:, 1::3]
    sim.set_state_init(q0, np.zeros(10))
    sim.reset(False)
    q, var, tac = [], [], []
    for t in range(T):
        sim.set_u(u[t])
        sim.forward(1)
        q.append(sim.get_q().copy())
        var.append(sim.get_variables().copy())
        if t % 5 == 4:
            tac.append(sim.get_tactile_force_vector().copy())
    d = sc.to_npz_dict()
    out.update(dclaw_ibuf=d["ibuf"], dclaw_dbuf=d["dbuf"], dclaw_joint_names=np.array(sc.joint_names), dclaw_body_names=np.array(sc.body_names),
               dclaw_ee_names=np.array([e["name"] for e in sc.end_effectors]),
               dclaw_damping=upd["damping"], dclaw_size=upd["size"], dclaw_ee=upd["ee"], dclaw_loc=upd["loc"], dclaw_q0=q0, dclaw_u=u,
               dclaw_q=np.array(q), dclaw_var=np.array(var), dclaw_tactile=np.array(tac))
    # ---- StableGrasp: box densities
    xml = os.path.join(ROOT, "oracle", "_ref", "assets", "stable_grasp", "stable_grasp.xml")
    sim = redmax_py.Simulation(xml)
    sc = compile_scene(xml)
    dens = rng.uniform(100.0, 1500.0, 11)
    for i in range(11):
        sim.update_body_density("box_%d" % i, float(dens[i]))
    g = stable_grasp_case.__globals__
    q0 = np.zeros(12)
    q0[2] = 0.2029862
    q0[4] = q0[5] = -0.03
    gp = 0.01
    q0[1] = gp
    stages = [np.array([0.0, gp, 0.2029862, 0.0, -0.03, -0.03]), np.array([0.0, gp, 0.2029862, 0.0, -0.008, -0.008]),
              np.array([0.0, gp, 0.2029862, 0.0, -0.008, -0.008]), np.array([0.0, gp, 0.2329862, 0.0, -0.008, -0.008])]
    T = 45
    steps = [20, 8, T - 28]
    u = []
    for st in range(3):
        for i in range(steps[st]):
            u.append((stages[st + 1] - stages[st]) / steps[st] * (i + 1) + stages[st])
    u = np.array(u)
    u[:, 2] += 0.003
    sim.set_state_init(q0, np.zeros(12))
    sim.reset(False)
    q, tac = [], []
    for t in range(T):
        sim.set_u(u[t])
        sim.forward(1)
        q.append(sim.get_q().copy())
        if t % 5 == 4:
            tac.append(sim.get_tactile_force_vector().copy())
    d = sc.to_npz_dict()
    out.update(sg_ibuf=d["ibuf"], sg_dbuf=d["dbuf"], sg_body_names=np.array(sc.body_names), sg_dens=dens, sg_q0=q0, sg_u=u,
               sg_q=np.array(q), sg_tactile=np.array(tac))
    return out


def perenv_case():
    """Per-ENVIRONMENT domain randomisation: K environments, each with its own parameters drawn like the reference's envs
    draw them per reset -- DClaw: cap joint damping, cap radius, end-effector position, cap joint location and initial
    finger pose (R/envs/dclaw_rotate_env.py:162-178); TactileInsertion: contact and tactile coefficients of both pads
    (R/envs/tactile_insertion_env.py:232-275) -- one reference Simulation per environment, short rollouts.  The fixture
    keeps the ORIGINAL scene blobs; the GPU test applies all K parameter sets to ONE batched Simulation."""
    out = {}
    rng = np.random.default_rng(11)
    # ---- DClaw
    K, T = 4, 30
    xml = os.path.join(ROOT, "oracle", "_ref", "assets", "dclaw_rotate", "dclaw_torque_control.xml")
    sc = compile_scene(xml)
    damping, radius = rng.uniform(0.01, 0.7, K), rng.uniform(0.02, 0.08, K)
    dxy = rng.uniform(-0.02, 0.02, (K, 2))
    q0 = np.zeros((K, 10))
    q0[:, [1, 4, 7]] = -0.5
    q0[:, [2, 5, 8]] = 0.8
    q0[:, :9] += 0.05 * rng.normal(size=(K, 9))
    u = rng.uniform(-1, 1, (T, K, 9))
    u[:, :, 1::3] = 0.6 + 0.4 * u[:, :, 1::3]
    q, var, tac = np.zeros((K, T, 10)), np.zeros((K, T, 12)), np.zeros((K, T // 5, 2718))
    for e in range(K):
        sim = redmax_py.Simulation(xml)
        sim.update_joint_damping("cap", damping[e])
        sim.update_body_size("cap", np.array([0.03, radius[e]]))
        sim.update_endeffector_position("cap", np.array([radius[e], 0.0, 0.0]))
        sim.update_joint_location("cap", np.array([dxy[e, 0], dxy[e, 1], 0.075]))
        sim.set_state_init(q0[e], np.zeros(10))
        sim.reset(False)
        for t in range(T):
            sim.set_u(u[t, e])
            sim.forward(1)
            q[e, t], var[e, t] = sim.get_q(), sim.get_variables()
            if t % 5 == 4:
                tac[e, t // 5] = sim.get_tactile_force_vector()
    d = sc.to_npz_dict()
    out.update(dclaw_ibuf=d["ibuf"], dclaw_dbuf=d["dbuf"], dclaw_joint_names=np.array(sc.joint_names), dclaw_body_names=np.array(sc.body_names),
               dclaw_ee_names=np.array([e_["name"] for e_ in sc.end_effectors]), dclaw_damping=damping, dclaw_radius=radius, dclaw_dxy=dxy,
               dclaw_q0=q0, dclaw_u=u, dclaw_q=q, dclaw_var=var, dclaw_tactile=tac)
    # ---- TactileInsertion
    K, T = 3, 45
    xml = os.path.join(ROOT, "oracle", "_ref", "assets", "tactile_insertion", "tactile_insertion.xml")
    sc = compile_scene(xml)
    cpar = np.stack([rng.uniform(2e3, 14e3, K), rng.uniform(20., 140., K), rng.uniform(0.5, 2.5, K), rng.uniform(0., 100., K)], axis=1)
    tpar = np.stack([rng.uniform(50, 450, K), rng.uniform(0.2, 2.3, K), rng.uniform(0.5, 2.5, K), rng.uniform(0., 100., K)], axis=1)
    q0 = np.zeros(12)
    q0[2] = 0.2
    q0[4] = q0[5] = -0.03
    u = np.zeros((T, 6))
    for t in range(T):
        g_ = min(1.0, (t + 1) / 15.0)
        m = max(0.0, (t - 25) / float(max(T - 25, 1)))
        u[t] = [0.004 * m, -0.003 * m, 0.2 - 0.002 * m, 0.15 * m, g_, g_]
    q, tac = np.zeros((K, T, 12)), np.zeros((K, T // 5, 780))
    for e in range(K):
        sim = redmax_py.Simulation(xml)
        for pad in ("tactile_pad_left", "tactile_pad_right"):
            sim.update_contact_parameters(pad, "box", *[float(x) for x in cpar[e]])
            sim.update_tactile_parameters(pad, *[float(x) for x in tpar[e]])
        sim.set_state_init(q0, np.zeros(12))
        sim.reset(False)
        for t in range(T):
            sim.set_u(u[t])
            sim.forward(1)
            q[e, t] = sim.get_q()
            if t % 5 == 4:
                tac[e, t // 5] = sim.get_tactile_force_vector()
    d = sc.to_npz_dict()
    out.update(ins_ibuf=d["ibuf"], ins_dbuf=d["dbuf"], ins_body_names=np.array(sc.body_names), ins_sensor_names=np.array([s_.name for s_ in sc.sensors]),
               ins_cpar=cpar, ins_tpar=tpar, ins_q0=q0, ins_u=u, ins_q=q, ins_tactile=tac)
    return out


def rolling_ball_bdf1_case(T, seed):
    """The rolling-ball scene (40x40 markers) under BDF1 with Simulation::backward(): the adjoint through the free3d-exp
    joint, the sphere SDF (ground, pad contact, tactile field) and the 2168-point pad.  Inputs: the script's action
    schedule compressed to T steps (press, then roll in x)."""
    adir = os.path.join(ROOT, "oracle", "_ref", "assets", "tactile_pad")
    src = os.path.join(adir, "tactile_pad_40x40.xml")
    xml = os.path.join(adir, "tactile_pad_40x40_bdf1.xml")
    txt = open(src).read()
    assert 'integrator="BDF2"' in txt
    open(xml, "w").write(txt.replace('integrator="BDF2"', 'integrator="BDF1"'))
    sim, probe, sc = redmax_py.Simulation(xml), redmax_probe.ProbeSimulation(xml), compile_scene(xml)
    n, nt, nu = sim.ndof_r, sim.ndof_tactile, sim.ndof_u
    u = np.zeros((T, nu))
    u[:, 2] = 0.2
    u[T // 2:, 0] = 0.1
    q0 = np.zeros(n)
    q0[2] = -0.012                      # pad just above the ball: contact within a few steps
    for s_ in (sim, probe):
        s_.set_state_init(q0, np.zeros(n))
        s_.reset(True)
    q, qd, tac, gp, mb = [], [], [], [], []
    for t in range(T):
        for s_ in (sim, probe):
            s_.set_u(u[t])
            s_.forward(1)
        q.append(sim.get_q().copy())
        qd.append(sim.get_qdot().copy())
        tac.append(sim.get_tactile_force_vector().copy())
        cs = probe.contact_sets()
        gp.append(cs["gp"][0])
        mb.append(np.asarray(cs["marker_body"][0], dtype=np.int32))
    rng = np.random.default_rng(1000 + seed)
    df_dq = rng.normal(size=(T, n))
    df_dtac = 1e-3 * rng.normal(size=(T, nt))
    bi = sim.backward_info
    bi.set_flags(True, True, False, True)
    bi.df_dq, bi.df_dtactile = df_dq.reshape(-1), df_dtac.reshape(-1)
    bi.df_dq0, bi.df_dqdot0, bi.df_du = np.zeros(n), np.zeros(n), np.zeros(nu * T)
    sim.backward()
    br = sim.backward_results
    ib, db = sc.pack()
    return dict(ibuf=ib, dbuf=db, q0=q0, qd0=np.zeros(n), u=u, q=np.array(q), qd=np.array(qd), tactile=np.array(tac),
                gp_ids=pad_ids(gp, max(max(len(x) for x in gp), 1)), marker_body=np.array(mb, dtype=np.int32),
                cot_seed=1000 + seed,      # df_dq = N(0,1) [T,n], df_dtactile = 1e-3 N(0,1) [T,nt] from default_rng(cot_seed), in this order
                df_dq0=np.array(br.df_dq0), df_dqdot0=np.array(br.df_dqdot0), df_du=np.array(br.df_du).reshape(T, nu))


def integrators_case(T, seed):
    """The TactilePush scene under the other two integrators of DH/Simulation.cpp:1076-1092 (forward only: the
    reference has no adjoint for them on this path either): options.integrator = BDF2 (SDIRK2 start-up step) and
    SDIRK2, same inputs as pusher13x10_episodic_s0.  The blob is the BDF1 one; only header slot 14 differs."""
    src = os.path.join(ASSETS, "pusher.xml")
    out = {}
    for name in ("BDF2", "SDIRK2"):
        xml = os.path.join(ASSETS, "pusher_%s.xml" % name.lower())
        txt = open(src).read()
        assert 'integrator="BDF1"' in txt
        open(xml, "w").write(txt.replace('integrator="BDF1"', 'integrator="%s"' % name))
        sim = redmax_py.Simulation(xml)
        probe = redmax_probe.ProbeSimulation(xml)
        assert sim.options.integrator == name
        q0, u = make_inputs(sim, T, seed)
        for s in (sim, probe):
            s.set_state_init(q0, np.zeros(sim.ndof_r))
            s.reset(False)
        q, qd, var, tac, gp = [], [], [], [], []
        for t in range(T):
            for s in (sim, probe):
                s.set_u(u[t])
                s.forward(1)
            q.append(sim.get_q().copy())
            qd.append(sim.get_qdot().copy())
            var.append(sim.get_variables().copy())
            tac.append(sim.get_tactile_force_vector().copy())
            gp.append(probe.contact_sets()["gp"][0])
        sc = compile_scene(xml)
        assert sc.integrator == name
        out.update({"q_" + name: np.array(q), "qd_" + name: np.array(qd), "var_" + name: np.array(var),
                    "tactile_" + name: np.array(tac), "gp_ids_" + name: pad_ids(gp, 66)})
    ib, db = compile_scene(src).pack()
    out.update(ibuf=ib, dbuf=db, q0=q0, qd0=np.zeros(len(q0)), u=u)
    return out


# a two-link arm on spherical joints (our own synthetic scene: no reference asset uses these joint types): ground
# contact at the tip, a 4x4 pad on the upper link pressed by the lower link, one end-effector, six force motors
SPHERICAL_TEMPLATE = '''<redmax model="spherical-chain">
    <option integrator="{integ}" timestep="5e-3" unit="m-kg" gravity="0. 0. -9.8"/>
    <ground pos="0 0 0" normal="0 0 1"/>
    <default>
        <ground_contact kn="2e3" kt="10" mu="0.8" damping="3"/>
        <tactile kn="50" kt="4" mu="1.0" damping="2"/>
        <motor ctrl="force" ctrl_range="-0.3 0.3"/>
    </default>
    <robot>
        <link name="upper">
            <joint name="shoulder" type="spherical-euler" pos="0 0 0.3" quat="1 0 0 0" damping="0.02"/>
            <body name="upper_body" type="cuboid" size="0.04 0.04 0.2" pos="0 0 -0.1" quat="1 0 0 0" density="800." general_contact_resolution="2 2 2"/>
            <link name="lower">
                <joint name="elbow" type="{elbow}" pos="0 0 -0.2" quat="1 0 0 0" damping="0.01"/>
                <body name="lower_body" type="cuboid" size="0.03 0.03 0.15" pos="0 0 -0.06" quat="1 0 0 0" density="800." general_contact_resolution="3 3 3"/>
            </link>
        </link>
    </robot>
    <actuator>
        <motor joint="shoulder" ctrl="force"/>
        <motor joint="elbow" ctrl="force"/>
    </actuator>
    <sensor>
        <tactile body="upper_body" name="pad" type="rect_array" rect_pos0="-0.015 -0.015 -0.1" rect_pos1="0.015 0.015 -0.1" axis0="1 0 0" axis1="0 1 0" resolution="4 4"/>
    </sensor>
    <contact>
        <ground_contact body="lower_body"/>
    </contact>
    <variable>
        <endeffector joint="elbow" pos="0 0 -0.135" name="tip"/>
    </variable>
</redmax>
'''


def spherical_xml(name, integ, elbow):
    d = os.path.join(ROOT, "oracle", "_ref", "assets", "synthetic")
    os.makedirs(d, exist_ok=True)
    xml = os.path.join(d, name + ".xml")
    open(xml, "w").write(SPHERICAL_TEMPLATE.format(integ=integ, elbow=elbow))
    return xml


def spherical_euler_case(T, seed):
    """spherical-euler joints (DH/Joint/JointSphericalEuler.cpp), BDF1, forward + backward()."""
    xml = spherical_xml("spherical_euler_bdf1", "BDF1", "spherical-euler")
    rng = np.random.default_rng(seed)
    return multi_case(xml, np.array([0.3, -0.2, 0.1, 0.2, 0.4, -0.3]), rng.uniform(-1, 1, (T, 6)), seed)


def spherical_exp_case(T, seed):
    """spherical-euler shoulder + spherical-exp elbow (DH/Joint/JointSphericalExp.cpp) under BDF2, forward only."""
    xml = spherical_xml("spherical_exp_bdf2", "BDF2", "spherical-exp")
    sim, probe, sc = redmax_py.Simulation(xml), redmax_probe.ProbeSimulation(xml), compile_scene(xml)
    rng = np.random.default_rng(seed)
    q0, u = np.array([0.3, -0.2, 0.1, 0.2, 0.4, -0.3]), rng.uniform(-1, 1, (T, 6))
    for s in (sim, probe):
        s.set_state_init(q0, np.zeros(6))
        s.reset(False)
    q, qd, var, tac, ground = [], [], [], [], []
    for t in range(T):
        for s in (sim, probe):
            s.set_u(u[t])
            s.forward(1)
        q.append(sim.get_q().copy())
        qd.append(sim.get_qdot().copy())
        var.append(sim.get_variables().copy())
        tac.append(sim.get_tactile_force_vector().copy())
        ground.append(probe.contact_sets()["ground"][0])
    ib, db = sc.pack()
    return dict(ibuf=ib, dbuf=db, q0=q0, qd0=np.zeros(6), u=u, q=np.array(q), qd=np.array(qd), var=np.array(var),
                tactile=np.array(tac), ground_ids=pad_ids(ground, len(sc.contact_points[1])))


def main():
    x13 = os.path.join(ASSETS, "pusher.xml")
    x32 = os.path.join(ASSETS, "pusher_32x13.xml")
    cases = {
        "pusher13x10_episodic_s0": lambda: episodic_case(x13, 60, 0),
        "pusher13x10_episodic_s1": lambda: episodic_case(x13, 60, 1, push=False),
        "pusher32x13_episodic_s0": lambda: episodic_case(x32, 30, 0),
        "pusher13x10_stepsim_s0": lambda: stepsim_case(x13, 8, 5, 0),
        "dclaw_episodic_s0": lambda: dclaw_case(40, 0),
        "insertion_episodic_s0": lambda: insertion_case(60, 0),
        "stable_grasp_episodic_s0": lambda: stable_grasp_case(50, 0),
        # synthetic sensor variants named by BASELINE.json configs[3] / [4] (oracle/build_ref.sh writes the XML files)
        "dclaw8x6_episodic_s0": lambda: dclaw_case(40, 0, "dclaw_torque_control_8x6.xml"),
        "insertion20x20_episodic_s0": lambda: insertion_case(60, 0, "tactile_insertion_20x20.xml"),
        "rollingball_bdf2_s0": rolling_ball_case,
        "pusher13x10_integrators_s0": lambda: integrators_case(40, 0),
        "spherical_euler_bdf1_s0": lambda: spherical_euler_case(60, 0),
        "rollingball_bdf1_adjoint_s0": lambda: rolling_ball_bdf1_case(60, 0),
        "free2d_plate_bdf1_s0": lambda: free2d_case(60, 0),
        "capsule_press_bdf1_s0": lambda: capsule_case(60, 0),
        "randomized_updates_s0": randomized_case,
        "perenv_updates_s0": perenv_case,
        "spherical_exp_bdf2_s0": lambda: spherical_exp_case(60, 0),
    }
    only = sys.argv[1:]          # optional: names of the fixtures to (re)generate
    if only:
        cases = {k: v for k, v in cases.items() if k in only}
    cases = {k: v() for k, v in cases.items()}
    for name, c in cases.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **c)
        print(name, os.path.getsize(path) // 1024, "KiB", "contacts/step:",
              [int((r >= 0).sum()) for r in c.get("gp_ids", np.zeros((0, 0)))][:12])


if __name__ == "__main__":
    main()
