"""Test helper: import the reference's OWN python callers of the path -- R/envs/redmax_torch_functions.py,
redmax_torch_env.py, tactile_push_env.py, R/utils/*.py, R/algorithms/gd.py, staged unmodified under oracle/_ref/py by
oracle/build_ref.sh -- against a chosen `redmax_py` module: the reference's pybind module (oracle/_ref) or the drop-in
tactilesimulation_b200.redmax.  gym / tensorboardX / matplotlib are absent from this image: tests/shims holds minimal
stand-ins."""
import importlib
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
PY_DIR = os.path.join(REF_DIR, "py")
SHIMS = os.path.join(ROOT, "tests", "shims")
_PKGS = ("envs", "utils", "algorithms")


def available():
    return os.path.exists(os.path.join(PY_DIR, "envs", "redmax_torch_functions.py")) and \
        os.path.exists(os.path.join(PY_DIR, "envs", "assets", "pusher", "pusher.xml"))


def reference_module():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.modules.pop("redmax_py", None)
    return importlib.import_module("redmax_py")


def load(redmax_module):
    """Returns a namespace with the reference's modules imported against `redmax_module` as redmax_py."""
    for p in (SHIMS, PY_DIR):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in list(sys.modules):
        if name.split(".")[0] in _PKGS + ("gym", "tensorboardX", "matplotlib"):
            del sys.modules[name]
    sys.modules["redmax_py"] = redmax_module

    class NS:
        pass
    ns = NS()
    ns.functions = importlib.import_module("envs.redmax_torch_functions")
    ns.push_env = importlib.import_module("envs.tactile_push_env")
    ns.envs = importlib.import_module("envs")
    ns.gd = importlib.import_module("algorithms.gd")
    ns.gym = importlib.import_module("gym")
    return ns


def gd_config(logdir, num_episodes=1, num_epochs=1, device="cpu"):
    import yaml
    cfg = yaml.safe_load(open(os.path.join(PY_DIR, "cfg", "gd_tactile.yaml")))
    cfg["params"]["general"] = dict(seed=0, device=device, checkpoint=None, train=True, logdir=logdir,
                                    save_interval=0, log_interval=1, render_interval=0)
    cfg["params"]["config"]["num_episodes"] = num_episodes
    cfg["params"]["config"]["num_epochs"] = num_epochs
    return cfg


# ---- the checks shared by the CPU test (reference module) and the GPU test (drop-in module)
def check_episodic_function(ns, xml, g, rel_err, device="cpu"):
    """The unmodified EpisodicSimFunction (R/envs/redmax_torch_functions.py:11-109) reproduces the golden gradients."""
    import numpy as np
    import torch
    sim = sys.modules["redmax_py"].Simulation(xml)
    T = g["u"].shape[0]
    q0 = torch.tensor(g["q0"], dtype=torch.double, device=device, requires_grad=True)
    qd0 = torch.tensor(g["qd0"], dtype=torch.double, device=device, requires_grad=True)
    u = torch.tensor(g["u"], dtype=torch.double, device=device, requires_grad=True)
    masks = torch.ones(T, dtype=torch.bool)
    qs, vs, tacs = ns.functions.EpisodicSimFunction.apply(q0, qd0, u, masks, sim, True)
    assert rel_err(qs.detach().cpu().numpy(), g["q"]) <= 1e-9
    assert rel_err(vs.detach().cpu().numpy(), g["var"]) <= 1e-9
    assert rel_err(tacs.detach().cpu().numpy(), g["tactile"]) <= 1e-8
    loss = (qs * torch.tensor(g["df_dq"], device=device)).sum() + (vs * torch.tensor(g["df_dvar"], device=device)).sum() + \
        (tacs * torch.tensor(g["df_dtactile"], device=device)).sum()
    loss.backward()
    assert rel_err(u.grad.cpu().numpy(), g["df_du"]) <= 1e-6
    assert rel_err(q0.grad.cpu().numpy(), g["df_dq0"]) <= 1e-6
    assert rel_err(qd0.grad.cpu().numpy(), g["df_dqdot0"]) <= 1e-6
    assert np.isfinite(u.grad.cpu().numpy()).all()


def check_stepsim_function(ns, xml, g, rel_err, device="cpu"):
    """The unmodified StepSimFunction (R/envs/redmax_torch_functions.py:112-174): forward(frame_skip,
    save_last_frame_var_only) per gym step, backward_steps chain driven by torch autograd; the (num_steps, ndof_u)
    gradient it returns for an (ndof_u,) action (:170) is summed over the sub-steps by autograd."""
    import torch
    sim = sys.modules["redmax_py"].Simulation(xml)
    fs, ns_ = int(g["frame_skip"]), g["u"].shape[0]
    sim.set_state_init(g["q0"], g["qd0"])
    sim.reset(backward_flag=True)
    acts, loss = [], 0.0
    for t in range(ns_):
        a = torch.tensor(g["u"][t], dtype=torch.double, device=device, requires_grad=True)
        q, var, tac = ns.functions.StepSimFunction.apply(a, fs, sim, True)
        assert rel_err(q.detach().cpu().numpy(), g["q"][t]) <= 1e-9
        assert rel_err(var.detach().cpu().numpy(), g["var"][t]) <= 1e-9
        assert rel_err(tac.detach().cpu().numpy(), g["tactile"][t]) <= 1e-8
        loss = loss + (q * torch.tensor(g["df_dq"][t], device=device)).sum() + (var * torch.tensor(g["df_dvar"][t], device=device)).sum() + \
            (tac * torch.tensor(g["df_dtactile"][t], device=device)).sum()
        acts.append(a)
    loss.backward()
    for t in range(ns_):
        assert rel_err(acts[t].grad.cpu().numpy(), g["df_du"][t].sum(axis=0)) <= 1e-6, t


def run_push_env(ns, steps=10, seed=3, gradient=True):
    """TactilePushEnv.reset / step (R/envs/tactile_push_env.py:133-232) for `steps` gym steps with seeded actions.
    Returns per-step observations and rewards (numpy) and d(sum reward)/d(actions)."""
    import numpy as np
    import torch
    env = ns.gym.make("TactilePush-v1", use_torch=True, gradient=gradient, observation_type="tactile_flatten")
    env.seed(seed)
    obs = env.reset()
    rng = np.random.RandomState(seed)
    obs_l, rew_l, acts, total = [obs.detach().cpu().numpy().copy()], [], [], 0.0
    for k in range(steps):
        a = torch.tensor(rng.normal(size=3), dtype=torch.double, requires_grad=gradient)
        obs, reward, done, info = env.step(a)
        obs_l.append(obs.detach().cpu().numpy().copy())
        rew_l.append(float(reward.detach().cpu()))
        acts.append(a)
        total = total + reward
    grads = None
    if gradient:
        total.backward()
        grads = np.stack([a.grad.numpy() for a in acts])
    return np.stack(obs_l), np.array(rew_l), grads


def run_gd_epoch(ns, logdir, num_episodes=1, device="cpu"):
    """GD.__init__ + one compute_reward_and_grad epoch (R/algorithms/gd.py:28-127, 220-264) under the default dtype the
    reference's training script sets (R/examples/TactilePushExp/train_tactile_push_gd.py:13)."""
    import torch
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        gd = ns.gd.GD(gd_config(logdir, num_episodes=num_episodes, device=device))
        gd.total_num_steps = 0
        gd.actor_optimizer.zero_grad()
        rewards, lens = gd.compute_reward_and_grad(gd.actor, num_episodes)
    finally:
        torch.set_default_dtype(old)
        torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))      # GD.__init__ pins torch to one thread (gd.py:30)
    return rewards, lens, gd
