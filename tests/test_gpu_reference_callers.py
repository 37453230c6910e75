"""north_star: "envs/redmax_torch_functions.py ... envs/redmax_torch_env.py and algorithms/gd.py run unchanged".

The reference's OWN files (staged unmodified under oracle/_ref/py by oracle/build_ref.sh; gym / tensorboardX / matplotlib
from tests/shims) are imported with sys.modules['redmax_py'] = tactilesimulation_b200.redmax and run on the GPU:
  * the unmodified EpisodicSimFunction / StepSimFunction reproduce the reference's golden gradients;
  * the unmodified TactilePushEnv (reset / step through StepSimFunction) gives the observations, rewards and action
    gradients of the SAME files run on the reference's own module (executed live on the box's CPU);
  * one GD.compute_reward_and_grad epoch of the unmodified gd.py gives the same episode reward and policy gradient;
  * the batched front-end (tactilesimulation_b200.envs.BatchedTactilePushEnv) replays episodes recorded from the
    reference's environment: same observations / rewards / action gradients per environment."""
import os

import numpy as np
import pytest
import torch

from tests import ref_callers as rc
from tests.conftest import GOLDEN, rel_err

pytestmark = pytest.mark.gpu
XML = os.path.join(rc.PY_DIR, "envs", "assets", "pusher", "pusher.xml")


def _need_ref():
    if not rc.available():
        pytest.fail("oracle/_ref/py (reference callers + assets) is missing on this box: run oracle/build_ref.sh before gpurun")


def _dropin():
    import tactilesimulation_b200.redmax as redmax
    return rc.load(redmax)


def test_unmodified_functions_reproduce_the_goldens_on_the_dropin():
    _need_ref()
    ns = _dropin()
    rc.check_episodic_function(ns, XML, np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz")), rel_err)
    rc.check_stepsim_function(ns, XML, np.load(os.path.join(GOLDEN, "pusher13x10_stepsim_s0.npz")), rel_err)


def test_unmodified_push_env_matches_itself_on_the_reference_module():
    _need_ref()
    o_ref, r_ref, g_ref = rc.run_push_env(rc.load(rc.reference_module()), steps=12)
    o_new, r_new, g_new = rc.run_push_env(_dropin(), steps=12)
    assert o_new.shape == o_ref.shape == (13, 393)
    assert np.abs(o_new[:, :3] - o_ref[:, :3]).max() <= 1e-9            # goal pose in the gripper frame
    assert rel_err(o_new[:, 3:], o_ref[:, 3:]) <= 1e-8                   # tactile field
    assert np.abs(o_ref[:, 3:]).max() > 0
    assert np.allclose(r_new, r_ref, rtol=1e-9, atol=1e-12)
    assert rel_err(g_new, g_ref) <= 1e-6


def test_unmodified_gd_epoch_matches_itself_on_the_reference_module(tmp_path):
    _need_ref()
    rew_ref, len_ref, gd_ref = rc.run_gd_epoch(rc.load(rc.reference_module()), str(tmp_path / "ref"))
    g_ref = np.concatenate([p.grad.reshape(-1).numpy() for p in gd_ref.actor.parameters() if p.grad is not None])
    rew_new, len_new, gd_new = rc.run_gd_epoch(_dropin(), str(tmp_path / "new"))
    g_new = np.concatenate([p.grad.reshape(-1).numpy() for p in gd_new.actor.parameters() if p.grad is not None])
    assert len_new == len_ref == [100]
    assert np.allclose(rew_new, rew_ref, rtol=1e-8)
    assert rel_err(g_new, g_ref) <= 1e-6 and np.abs(g_ref).max() > 0


def test_batched_frontend_replays_episodes_of_the_reference_env():
    """B = 4 episodes of the unmodified TactilePushEnv on the reference module (different seeds: initial state, goal,
    random pushes and actions recorded) replayed through BatchedTactilePushEnv in ONE batch."""
    _need_ref()
    from tactilesimulation_b200.envs import BatchedTactilePushEnv
    from tactilesimulation_b200.redmax import Simulation
    ns = rc.load(rc.reference_module())
    B, steps = 4, 12
    rec = []
    for e in range(B):
        env = ns.gym.make("TactilePush-v1", use_torch=True, gradient=True, observation_type="tactile_flatten")
        env.seed(10 + e)
        obs = [env.reset().detach().numpy().copy()]
        q0, goal = env.unwrapped.state_q.numpy().copy(), env.unwrapped.goal.numpy().copy()
        rng = np.random.RandomState(100 + e)
        acts, ext, rews, total = [], [], [], 0.0
        for k in range(steps):
            a = torch.tensor(rng.normal(size=3), dtype=torch.double, requires_grad=True)
            o, r, done, info = env.step(a)
            obs.append(o.detach().numpy().copy())
            ext.append(env.unwrapped.external_force.copy())
            rews.append(float(r.detach()))
            acts.append(a)
            total = total + r
        total.backward()
        rec.append(dict(q0=q0, goal=goal, obs=np.stack(obs), ext=np.stack(ext), rew=np.array(rews),
                        act=np.stack([a.detach().numpy() for a in acts]), grad=np.stack([a.grad.numpy() for a in acts])))
    sim = Simulation(XML, batch=B)
    benv = BatchedTactilePushEnv(sim, observation_type="tactile_flatten", gradient=True)
    dev = sim.device
    obs = [benv.reset(q0=torch.tensor(np.stack([r["q0"] for r in rec])), goal=torch.tensor(np.stack([r["goal"] for r in rec])))]
    acts, total, rews = [], 0.0, []
    for k in range(steps):
        a = torch.tensor(np.stack([r["act"][k] for r in rec]), device=dev, requires_grad=True)
        o, r, done, info = benv.step(a, external_force=torch.tensor(np.stack([rc_["ext"][k] for rc_ in rec])))
        obs.append(o)
        rews.append(r.detach().cpu().numpy())
        acts.append(a)
        total = total + r.sum()
    total.backward()
    for e in range(B):
        o_new = np.stack([o[e].detach().cpu().numpy() for o in obs])
        assert np.abs(o_new[:, :3] - rec[e]["obs"][:, :3]).max() <= 1e-9, e
        assert rel_err(o_new[:, 3:], rec[e]["obs"][:, 3:]) <= 1e-8, e
        assert np.allclose(np.array([r[e] for r in rews]), rec[e]["rew"], rtol=1e-9, atol=1e-12), e
        assert rel_err(np.stack([a.grad[e].cpu().numpy() for a in acts]), rec[e]["grad"]) <= 1e-6, e


def test_batched_insertion_frontend_matches_the_reference_env():
    """BatchedTactileInsertionEnv (BASELINE configs[4] is defined through this env) against the UNMODIFIED
    R/envs/tactile_insertion_env.py run on the reference module: grasp pose of generate_initial_pose, then B = 3 episodes
    (reset with given noises, two steps with given actions): observations, rewards, success flags."""
    _need_ref()
    from tactilesimulation_b200.envs import BatchedTactileInsertionEnv
    from tactilesimulation_b200.redmax import Simulation
    ns = rc.load(rc.reference_module())
    kw = dict(observation_type="tactile_flatten", observation_noise=False, normalize_tactile_obs=True, allow_translation=True,
              allow_rotation=True, action_type="relative", reward_type="delta", domain_randomization=False)
    # (the env builds its action ramp with torch.zeros((T, 6)): under torch's default float32 the position targets would be
    # rounded to 6e-8 relative, which the stiff position control turns into 1e-5 of tactile force; the comparison runs
    # the reference under float64 defaults, like R/examples/TactilePushExp/train_tactile_push_gd.py:13 sets them)
    torch.set_default_dtype(torch.float64)
    try:
        return _insertion_episodes(ns, kw)
    finally:
        torch.set_default_dtype(torch.float32)


def _insertion_episodes(ns, kw):
    from tactilesimulation_b200.envs import BatchedTactileInsertionEnv
    from tactilesimulation_b200.redmax import Simulation
    env = ns.gym.make("Insertion-v3", use_torch=True, verbose=False, render_tactile=False, **kw)
    B = 3
    rng = np.random.RandomState(4)
    pos = np.stack([rng.uniform(-0.006, 0.006, B), rng.uniform(-0.006, 0.006, B), rng.uniform(-0.0002, 0.0002, B)], axis=1)
    rot = rng.uniform(-np.pi / 18, np.pi / 18, B)
    gh = rng.uniform(-0.01, 0.005, B)
    acts = rng.uniform(-1, 1, (2, B, 3))
    rec = []
    for e in range(B):
        o = [env.reset(position_noise=pos[e], rotation_noise=rot[e], grasp_height_noise=gh[e]).detach().numpy().copy()]
        rs, ds = [], []
        for k in range(2):
            ob, r, d, info = env.step(torch.tensor(acts[k, e]))
            o.append(ob.detach().numpy().copy())
            rs.append(float(r))
            ds.append(bool(info["success"]))
        rec.append((o, rs, ds))
    xml = os.path.join(rc.PY_DIR, "envs", "assets", "tactile_insertion", "tactile_insertion.xml")
    benv = BatchedTactileInsertionEnv(Simulation(xml, batch=B), **kw)
    assert rel_err(benv.q_init_reference[0].cpu().numpy(), env.unwrapped.q_init_reference) <= 1e-8
    # The grasp pose is the end of 1 000 contact-rich sim-steps; the pads penetrate the box by ~1e-4 m, so the 1e-9 m left
    # of that rollout's rounding would show as 1e-5 in the RELATIVE tactile observations.  The episodes below start
    # from the reference's pose so that they test the env logic (and 45-step rollouts), not that amplification.
    benv.q_init_reference = torch.tensor(np.tile(env.unwrapped.q_init_reference, (B, 1)), device=benv.device)
    obs = [benv.reset(position_noise=torch.tensor(pos), rotation_noise=torch.tensor(rot), grasp_height_noise=torch.tensor(gh))]
    rews, dones = [], []
    for k in range(2):
        ob, r, d, info = benv.step(torch.tensor(acts[k]))
        obs.append(ob)
        rews.append(r.cpu().numpy())
        dones.append(d.cpu().numpy())
    for e in range(B):
        for k in range(3):
            assert rel_err(obs[k][e].cpu().numpy(), rec[e][0][k]) <= 1e-6, (e, k)
        assert np.allclose([r[e] for r in rews], rec[e][1], rtol=1e-7, atol=1e-9), e
        assert [bool(d[e]) for d in dones] == rec[e][2], e


@pytest.mark.parametrize("torque", [False, True])
def test_batched_dclaw_frontend_matches_the_reference_env(torque):
    """BatchedDClawRotateEnv (BASELINE configs[3]) against the UNMODIFIED R/envs/dclaw_rotate_env.py on the reference module:
    B = 3 environments, each with its OWN randomised cap (damping / radius / end-effector / joint location read back from
    the reference env's seeded draws), eight steps with given actions: observations, rewards, done / success flags."""
    _need_ref()
    from tactilesimulation_b200.envs import BatchedDClawRotateEnv
    from tactilesimulation_b200.redmax import Simulation
    ns = rc.load(rc.reference_module())
    B, steps = 3, 8
    rng = np.random.RandomState(21)
    acts = rng.uniform(-1.2, 1.2, (steps, B, 9))
    rec = []
    for e in range(B):
        env = ns.gym.make("TactileRotation-v1", use_torch=False, observation_type="tactile_flatten", render_tactile=False,
                          torque_control=torque, relative_control=True)
        env.seed(50 + e)
        # replay the env's own reset draws (dclaw_rotate_env.py:163-169) to know what it randomised
        r2 = np.random.RandomState(50 + e)
        qn = r2.randn(9) * 0.05
        damping, radius = r2.uniform(low=0.01, high=0.7), r2.uniform(low=0.02, high=0.08)
        dxy = r2.uniform(low=-0.02, high=0.02, size=2)
        obs = [np.asarray(env.reset()).copy()]
        rews, dones = [], []
        for k in range(steps):
            o, r, d, info = env.step(acts[k, e].copy())
            obs.append(np.asarray(o).copy())
            rews.append(float(r))
            dones.append((bool(d), bool(info["success"])))
        rec.append(dict(qn=qn, damping=damping, radius=radius, dxy=dxy, obs=np.stack(obs), rew=np.array(rews), done=dones))
    xml = os.path.join(rc.PY_DIR, "envs", "assets", "dclaw_rotate",
                       "dclaw_torque_control.xml" if torque else "dclaw_position_control.xml")
    benv = BatchedDClawRotateEnv(Simulation(xml, batch=B), observation_type="tactile_flatten", torque_control=torque)
    obs = [benv.reset(q_noise=torch.tensor(np.stack([r["qn"] for r in rec])), damping=np.array([r["damping"] for r in rec]),
                      radius=np.array([r["radius"] for r in rec]), dxy=np.stack([r["dxy"] for r in rec]))]
    rews, dones = [], []
    for k in range(steps):
        o, r, d, info = benv.step(torch.tensor(acts[k]))
        obs.append(o)
        rews.append(r.cpu().numpy())
        dones.append((d.cpu().numpy(), info["success"].cpu().numpy()))
    for e in range(B):
        o_new = np.stack([o[e].cpu().numpy() for o in obs])
        assert o_new.shape == rec[e]["obs"].shape == (steps + 1, 18 + 3 * 20 * 20 * 3)
        assert np.abs(o_new[:, :18] - rec[e]["obs"][:, :18]).max() <= 1e-8, e          # joint angles, fingertip positions
        assert rel_err(o_new[:, 18:], rec[e]["obs"][:, 18:]) <= 1e-6, e                # tactile flow images
        assert np.allclose(np.array([r[e] for r in rews]), rec[e]["rew"], rtol=1e-7, atol=1e-9), e
        assert [(bool(d[e]), bool(s_[e])) for d, s_ in dones] == rec[e]["done"], e


def test_batched_stable_grasp_frontend_matches_the_reference_env():
    """BatchedStableGraspEnv against the UNMODIFIED R/envs/stable_grasp_env.py on the reference module: B = 3 bars with
    their own box densities (read back from the reference env's seeded reset), the first grasp and two more with given
    actions: observations, rewards, success."""
    _need_ref()
    torch.set_default_dtype(torch.float64)          # (the env builds its targets with default-dtype tensors, see the insertion test)
    try:
        from tactilesimulation_b200.envs import BatchedStableGraspEnv
        from tactilesimulation_b200.redmax import Simulation
        ns = rc.load(rc.reference_module())
        B = 3
        rng = np.random.RandomState(9)
        acts = rng.uniform(-1, 1, (2, B, 1))
        rec = []
        for e in range(B):
            env = ns.gym.make("StableGrasp-v1", use_torch=True, observation_type="tactile_flatten", render_tactile=False)
            env.seed(70 + e)
            env.unwrapped.render = lambda *a, **k: None          # (the env opens the viewer when an episode succeeds)
            obs = [env.reset().detach().numpy().copy()]
            dens = env.unwrapped.block_densitys.copy()
            rews, succ = [float(env.unwrapped.reward_buf)], [bool(env.unwrapped.is_success)]
            for k in range(2):
                o, r, d, info = env.step(torch.tensor(acts[k, e]))
                obs.append(o.detach().numpy().copy())
                rews.append(float(r))
                succ.append(bool(info["success"]))
            rec.append(dict(dens=dens, obs=np.stack(obs), rew=np.array(rews), succ=succ, q0=env.unwrapped.qpos_init_reference.numpy().copy()))
        xml = os.path.join(rc.PY_DIR, "envs", "assets", "stable_grasp", "stable_grasp.xml")
        benv = BatchedStableGraspEnv(Simulation(xml, batch=B), observation_type="tactile_flatten")
        assert rel_err(benv.qpos_init_reference[0].cpu().numpy(), rec[0]["q0"]) <= 1e-8
        benv.qpos_init_reference = torch.tensor(np.stack([r["q0"] for r in rec]), device=benv.device)      # same start (see the insertion test)
        obs = [benv.reset(block_densitys=np.stack([r["dens"] for r in rec]))]
        rews, succ = [benv.reward_buf.cpu().numpy()], [benv.is_success.cpu().numpy()]
        for k in range(2):
            o, r, d, info = benv.step(torch.tensor(acts[k]))
            obs.append(o)
            rews.append(r.cpu().numpy())
            succ.append(info["success"].cpu().numpy())
        for e in range(B):
            o_new = np.stack([o[e].cpu().numpy() for o in obs])
            assert o_new.shape == rec[e]["obs"].shape == (3, 520)
            assert rel_err(o_new, rec[e]["obs"]) <= 1e-6, e
            assert np.allclose(np.array([r[e] for r in rews]), rec[e]["rew"], rtol=1e-6, atol=1e-9), e
            assert [bool(s_[e]) for s_ in succ] == rec[e]["succ"], e
    finally:
        torch.set_default_dtype(torch.float32)
