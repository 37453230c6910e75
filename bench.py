#!/usr/bin/env python
"""Benchmark of the hot path named by BASELINE.json: TactilePush (pusher scene, 32x13 tactile
pad), forward + reverse-time adjoint, horizon T sim-steps, batch B environments per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

* one "step" = one pass of the hot path over one batch: tsim_forward (T implicit steps with
  tape, q / var / tactile written every step: step loop + tape pass + tactile pass) + tsim_backward
  (cotangent pull-back pass + reverse sweep, cotangents on q, var and the tactile field) for B environments; for N > 1 followed by ONE NCCL all-reduce of a
  policy-gradient sized buffer.  unit of `value`: env-steps/s, 1 env-step = one
  simulation timestep of one environment including its share of the reverse sweep.
* `value`   : device-timed, inputs already resident in HBM.
* `e2e`     : the same metric through the public plugin API (EpisodicSimFunction.apply +
  loss.backward()) with pinned HOST inputs copied in and gradients/loss copied out every step.
* `roofline`: dominant kernel (fwd_kernel, the step loop) against the measured HBM peak, timed by CUDA
  events recorded inside the C ABI around every kernel (tsim_scene_kernel_times); `roofline.kernels` lists
  all five kernels of a step.  The step loop is fp64 latency-bound, its fraction is small by construction;
  the read-out / pull-back passes are the ones that stream.
* `cpu_baseline` / `--impl reference`: the unmodified reference C++ (oracle/_ref/redmax_py, built
  by oracle/build_ref.sh) on the host cores, one process per core, on a bounded sample.

--workload selects one of the other BASELINE.json configs (development / profiles arms; the default, `push`, is the
headline configs[2] and is the only line the driver reads):
    push_fwd   configs[1]  TactilePush 32x13 forward-only, B=1024, T=200, tactile every step
    dclaw      configs[3]  DClaw rotate cap, 3 x (8x6) pads (synthetic spec of oracle/build_ref.sh), B=2048, fwd+adjoint, T=200
    insertion  configs[4]  TactileInsertion 2 x (20x20) pads, forward-only Episodic rollout, T=45, tactile on 6 masked frames,
                           B=1024 per GPU (8192 over 8 GPUs)
    stepsim    the gd.py shape (SURVEY.md 8d "gym-steps/s"): BatchedTactilePushEnv 13x10, B=4096, 100 gym steps x frame_skip 5
               through StepSimFunction + loss.backward(); unit gym-steps/s
Every arm has its reference arm (--impl reference --workload W: the unmodified C++ on the host cores, same generator).

Nothing here reads /root/reference.  oracle/ is used only by the cpu_baseline / reference legs.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
SCENE_CASE = "pusher32x13_episodic_s0"          # packed scene blob of pusher.xml with resolution="32 13"
REF_XML = os.path.join(REF_DIR, "assets", "pusher", "pusher_32x13.xml")
METRIC = "env-steps/sec fwd+adjoint TactilePush 32x13 tactile, batch 4096"
POLICY_GRAD_SIZE = 84486                        # obs 3+1248 -> 64 -> 64 -> 3 (+logstd), SURVEY.md 8e
STEPSIM_METRIC = "gym-steps/sec fwd+backward TactilePush 13x10 through StepSimFunction (gd.py shape), batch 4096"
N_MARKERS = 416


# ------------------------------------------------------------------ synthetic inputs
def make_inputs(q_init, B, T, seed):
    """q0[1] = -0.001, q0[4] ~ U(-.02,.02) (tactile_push_env.py:135-136); u = tanh(N(0,1)) on dims
    0-2, dims 3-4 resampled every 10 steps (p=.5 zero else U(-1,1)) (:185-193), dim 5 = 0."""
    rng = np.random.default_rng(seed)
    n = q_init.shape[0]
    q0 = np.tile(q_init, (B, 1))
    q0[:, 1] = -0.001
    q0[:, 4] = rng.uniform(-0.02, 0.02, B)
    u = np.zeros((T, B, 6))
    u[:, :, :3] = np.tanh(rng.normal(size=(T, B, 3)))
    ext = np.zeros((B, 2))
    for t in range(T):
        if t % 10 == 0:
            ext = np.where(rng.uniform(size=(B, 1)) < 0.5, rng.uniform(-1, 1, (B, 2)), 0.0)
        u[t, :, 3:5] = ext
    goal = np.zeros((B, 3))
    goal[:, 0] = rng.uniform(0.15, 0.25, B)
    goal[:, 1] = rng.uniform(-0.2, 0.2, B)
    goal[:, 2] = rng.uniform(goal[:, 1] * np.pi - np.pi / 16, goal[:, 1] * np.pi + np.pi / 16)
    return q0, np.zeros((B, n)), u, goal


def reward_cotangents(q_traj, var, goal, scale):
    """d(-sum reward)/d(q, var) of the TactilePush reward (tactile_push_env.py:206-211), torch ops."""
    import torch
    dq = torch.zeros_like(q_traj)
    dq[:, :, 3:5] = 2.0 * (q_traj[:, :, 3:5] - goal[None, :, 0:2]) / (0.01 ** 2) * 0.01
    dq[:, :, 6] = 2.0 * (q_traj[:, :, 6] - goal[None, :, 2]) / ((np.pi / 36.0) ** 2) * 0.1
    dv = torch.empty_like(var)
    d = 2.0 * (var[:, :, 0:3] - var[:, :, 3:6]) / (0.02 ** 2)
    dv[:, :, 0:3] = d
    dv[:, :, 3:6] = -d
    return dq * scale, dv * scale


# ------------------------------------------------------------------ workloads (BASELINE.json configs)
WORKLOADS = {
    "push": dict(case=SCENE_CASE, xml=("pusher", "pusher_32x13.xml"), B=4096, T=200, grad=True, metric=METRIC,
                 label="TactilePush 32x13 fwd+adjoint horizon {T}, batch {B}/GPU (BASELINE configs[2])"),
    "push_fwd": dict(case=SCENE_CASE, xml=("pusher", "pusher_32x13.xml"), B=1024, T=200, grad=False,
                     metric="env-steps/sec forward-only TactilePush 32x13 tactile, batch 1024",
                     label="TactilePush 32x13 forward-only horizon {T}, batch {B}/GPU, tactile every step (BASELINE configs[1])"),
    "dclaw": dict(case="dclaw8x6_episodic_s0", xml=("dclaw_rotate", "dclaw_torque_control_8x6.xml"), B=2048, T=200, grad=True,
                  metric="env-steps/sec fwd+adjoint DClaw rotate cap 3x(8x6) tactile, batch 2048",
                  label="DClaw rotate cap, 9-DoF 3-finger + cap, 3 x (8x6) pads (synthetic), fwd+adjoint horizon {T}, batch {B}/GPU (BASELINE configs[3])"),
    "insertion": dict(case="insertion20x20_episodic_s0", xml=("tactile_insertion", "tactile_insertion_20x20.xml"), B=1024, T=45,
                      grad=False, metric="env-steps/sec forward-only TactileInsertion 2x(20x20) tactile rollout, batch 8192 over 8 GPUs",
                      label="TactileInsertion 2 x (20x20) pads (synthetic), forward-only Episodic rollout of {T} sim-steps, tactile on 6 "
                            "masked frames, batch {B}/GPU (BASELINE configs[4]: 8192 over 8 GPUs)"),
}
INSERTION_FRAMES = (6, 20, 26, 32, 38, 44)      # tactile_masks of tactile_insertion_env.py:75-77 (initial frame 15, 5 samples)


def workload_inputs(name, g, B, T, seed):
    """Seeded synthetic inputs of a workload: (q0 [B,n], qd0 [B,n], u [T,B,nu], goal or None).  The reference arm calls
    this with B = 1 and the environment's own seed."""
    if name in ("push", "push_fwd"):
        return make_inputs(g["q0"], B, T, seed)
    rng = np.random.default_rng(seed)
    n, nu = g["q0"].shape[0], g["u"].shape[1]
    q0 = np.tile(g["q0"], (B, 1))
    if name == "dclaw":
        # dclaw_rotate_env.py:74-77,162-166: fingers at (-0.5, 0.8) + N(0, 0.05) on the nine joints; actions U(-1,1)^9
        q0[:, :9] += 0.05 * rng.normal(size=(B, 9))
        u = rng.uniform(-1.0, 1.0, (T, B, nu))
        return q0, np.zeros((B, n)), u, None
    if name == "insertion":
        # tactile_insertion_env.py:343-357: position-target ramp of the gripper base + grasp force; here the golden's
        # schedule (fingers closing, then the base moving the box against the hole), perturbed per environment
        gu = g["u"]
        u = np.stack([gu[t % len(gu)] for t in range(T)])[:, None, :].repeat(B, axis=1)
        u[:, :, :4] += 2e-4 * rng.normal(size=(T, B, 4))
        return q0, np.zeros((B, n)), u, None
    raise ValueError(name)


def workload_cotangents(name, q_traj, var, goal, scale):
    """Cotangents on (q, var) of a fwd+adjoint workload (torch, device): the TactilePush reward gradient; for DClaw the
    gradient of -sum(cap angle) (the rotate-cap objective, dclaw_rotate_env.py:_get_reward) plus a fingertip term."""
    import torch
    if name == "push":
        return reward_cotangents(q_traj, var, goal, scale)
    dq = torch.zeros_like(q_traj)
    dq[:, :, 9] = -scale
    dv = 2.0 * var * scale if var is not None else None
    return dq, dv


# ------------------------------------------------------------------ clocks sampler (pynvml)
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------ reference (CPU) arm
def _ref_worker(args):
    """One process = one core: E environments x T steps of a workload through the unmodified reference redmax_py
    (EpisodicSimFunction call pattern, redmax_torch_functions.py:35-109): forward (+tape) and, for the fwd+adjoint
    workloads, backward()."""
    xml, seeds, T, name, case = args
    sys.path.insert(0, REF_DIR)
    import redmax_py
    sim = redmax_py.Simulation(xml)
    n, nv, nt, nu = sim.ndof_r, sim.ndof_var, sim.ndof_tactile, sim.ndof_u
    g = np.load(os.path.join(GOLDEN, case + ".npz"))
    grad = WORKLOADS[name]["grad"] if name in WORKLOADS else True
    frames = set(INSERTION_FRAMES) if name == "insertion" else None
    t0 = time.perf_counter()
    steps = 0
    for seed in seeds:
        q0, qd0, u, goal = workload_inputs(name, g, 1, T, seed)
        sim.set_state_init(q0[0], qd0[0])
        sim.reset(grad)
        qs, vs = np.zeros((T, n)), np.zeros((T, nv))
        for t in range(T):
            sim.set_u(u[t, 0])
            sim.forward(1)
            qs[t] = sim.get_q()
            if nv:
                vs[t] = sim.get_variables()
            if frames is None or t in frames:
                sim.get_tactile_force_vector()
        if grad:
            if name == "push":
                dq = np.zeros((T, n))
                dq[:, 3:5] = 2.0 * (qs[:, 3:5] - goal[0, 0:2]) / (0.01 ** 2) * 0.01
                dq[:, 6] = 2.0 * (qs[:, 6] - goal[0, 2]) / ((np.pi / 36.0) ** 2) * 0.1
                d = 2.0 * (vs[:, 0:3] - vs[:, 3:6]) / (0.02 ** 2)
                dv = np.concatenate([d, -d], axis=1)
            else:
                dq = np.zeros((T, n))
                dq[:, 9] = -1.0
                dv = 2.0 * vs
            bi = sim.backward_info
            bi.set_flags(True, True, False, True)
            bi.df_dq = dq.reshape(-1)
            bi.df_dvar = dv.reshape(-1)
            bi.df_dtactile = np.full(nt * T, 1e-3)
            bi.df_dq0, bi.df_dqdot0, bi.df_du = np.zeros(n), np.zeros(n), np.zeros(nu * T)
            sim.backward()
        steps += T
    return steps, time.perf_counter() - t0


def ref_xml(name):
    d, f = WORKLOADS[name]["xml"]
    return os.path.join(REF_DIR, "assets", d, f)


def ref_pass(pool, cores, envs_per_core, T, seed0, name="push"):
    """One bounded CPU sample: cores x envs_per_core environments x T steps.  Returns (env-steps, wall s)."""
    jobs = [(ref_xml(name), [seed0 + c * envs_per_core + e for e in range(envs_per_core)], T, name, WORKLOADS[name]["case"])
            for c in range(cores)]
    t0 = time.perf_counter()
    res = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall


def ref_available():
    if not os.path.exists(REF_XML):
        return False
    return any(f.startswith("redmax_py") and f.endswith(".so") for f in os.listdir(REF_DIR))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _ref_stepsim_worker(args):
    """One process: E episodes of the UNMODIFIED TactilePushEnv + StepSimFunction (the gd.py shape: 100 gym steps x
    frame_skip 5, loss = -sum reward, loss.backward()) on the reference module.  Returns (gym-steps, seconds)."""
    seeds, steps = args
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from tests import ref_callers as rc
    import torch
    torch.set_num_threads(1)
    ns = rc.load(rc.reference_module())
    env = ns.gym.make("TactilePush-v1", use_torch=True, gradient=True, observation_type="tactile_flatten")
    t0 = time.perf_counter()
    n = 0
    for seed in seeds:
        env.seed(seed)
        env.reset()
        rng = np.random.RandomState(seed)
        total, acts = 0.0, []
        for k in range(steps):
            a = torch.tensor(rng.normal(size=3), dtype=torch.double, requires_grad=True)
            obs, r, done, info = env.step(a)
            total = total + r
            acts.append(a)
            n += 1
        (-total).backward()
    return n, time.perf_counter() - t0


def run_reference(a):
    """--impl reference: the reference's own CPU implementation on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    sys.path.insert(0, REF_DIR)
    import redmax_py  # noqa: F401  (loaded once in the parent: the forked workers inherit it, the driver's hook sees the .so)
    cores = host_cores()
    name = a.workload
    build = "oracle/build_ref.sh: g++ -O3 -DNDEBUG (the reference's CMake Release flags)"
    if name == "stepsim":
        gsteps = a.gym_steps
        sample = f"{cores} processes x {a.ref_envs_per_core} episodes x {gsteps} gym steps (unmodified tactile_push_env.py + StepSimFunction, fwd + loss.backward())"
        with mp.get_context("fork").Pool(cores) as pool:
            for w in range(min(a.warmup, 1)):
                pool.map(_ref_stepsim_worker, [([9000 + c], 5) for c in range(cores)])
            tot, wall = 0, 0.0
            for k in range(a.steps):
                t0 = time.perf_counter()
                res = pool.map(_ref_stepsim_worker, [([1000 * k + c * a.ref_envs_per_core + e for e in range(a.ref_envs_per_core)], gsteps)
                                                     for c in range(cores)])
                wall += time.perf_counter() - t0
                tot += sum(r[0] for r in res)
        val, unit, metric = tot / wall, "gym-steps/s", STEPSIM_METRIC
        label = f"TactilePush 13x10 gd.py shape: {gsteps} gym steps x frame_skip 5, fwd + loss.backward(), bounded sample of the batch-{a.batch or 4096} job"
        T = gsteps
    else:
        wl = WORKLOADS[name]
        T = a.horizon or wl["T"]
        envs_per_core = a.ref_envs_per_core
        what = "fwd+tape+backward, tactile read every step" if wl["grad"] else \
            ("forward only, tactile read on %d masked frames" % len(INSERTION_FRAMES) if name == "insertion" else "forward only, tactile read every step")
        sample = f"{cores} processes x {envs_per_core} env x T={T} sim-steps per step ({what})"
        with mp.get_context("fork").Pool(cores) as pool:
            for w in range(a.warmup):
                ref_pass(pool, cores, 1, min(T, 20), 10_000 + w, name)
            tot_steps, wall = 0, 0.0
            for k in range(a.steps):
                s_, w_ = ref_pass(pool, cores, envs_per_core, T, 1000 * k, name)
                tot_steps += s_
                wall += w_
        val, unit, metric = tot_steps / wall, "env-steps/s", wl["metric"]
        label = wl["label"].format(T=T, B=a.batch or wl["B"]) + ": bounded sample"
    line = {"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * wall / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "horizon": T},
            "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": "reference", "sample": sample, "build": build},
            "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ B200 arm
def run_b200(a):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from tactilesimulation_b200.redmax import Simulation
    from tactilesimulation_b200.sim import BatchedSim
    from tactilesimulation_b200.torch_functions import EpisodicSimFunction
    from tactilesimulation_b200.layout import scene_from_blob

    name = a.workload
    wl = WORKLOADS[name]
    grad = wl["grad"]
    g = np.load(os.path.join(GOLDEN, wl["case"] + ".npz"))
    B, T = a.batch or wl["B"], a.horizon or wl["T"]
    lanes = a.lanes if (a.lanes and int(g["ibuf"][3]) <= 8) else None      # (8 lanes: the 8-dof kernel variant only)
    core = BatchedSim((g["ibuf"], g["dbuf"]), device=dev, lanes=lanes)
    n, nu, nvar, ntac = core.ndof_r, core.ndof_u, core.ndof_var, core.ndof_tactile
    from tactilesimulation_b200.distributed import rank_seed
    # weak scaling: every rank owns B environments of the global batch [rank*B, (rank+1)*B), own RNG stream
    q0_h, qd0_h, u_h, goal_h = workload_inputs(name, g, B, T, seed=rank_seed(1234, rank))
    q0, qd0 = torch.tensor(q0_h, device=dev), torch.tensor(qd0_h, device=dev)
    u = torch.tensor(u_h, device=dev)
    goal = torch.tensor(goal_h, device=dev) if goal_h is not None else None
    # tactile frames: every step, or the masked frames of the insertion rollout (tactile_insertion_env.py:75-77)
    frames = [t for t in INSERTION_FRAMES if t < T] if name == "insertion" else list(range(T))
    tac_rows = None if len(frames) == T else [frames.index(t) if t in frames else -1 for t in range(T)]
    dtac = torch.full((T, B, ntac), 1e-3, dtype=torch.float64, device=dev) if grad else None   # tactile cotangent, resident
    gradbuf = torch.zeros(POLICY_GRAD_SIZE, dtype=torch.float64, device=dev)
    scale = 1.0 / (B * world)
    # preallocated outputs are reused by every step (180 GB HBM: tactile [T,B,3M] is 8.2 GB at B=4096)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    fwd_ms, bwd_ms = [], []

    def one_step(timed):
        q, qd = q0.clone(), qd0.clone()
        if timed:
            ev[0].record()
        out = core.forward(q, qd, u, T, grad=grad, tac_rows=tac_rows)
        if timed:
            ev[1].record()
        if not grad:                       # forward-only rollout: no adjoint, no gradient exchange
            if timed:
                ev[2].record()
                ev[3].record()
            return out, None
        dq, dv = workload_cotangents(name, out["q_traj"], out["var"], goal, scale)
        if timed:
            ev[2].record()
        bw = core.backward(out, u, T, dq, dv, dtac, want_q0=True)
        if timed:
            ev[3].record()
        gradbuf[:nu] = bw["df_du"].sum(dim=(0, 1))
        if world > 1:
            dist.all_reduce(gradbuf)
        return out, bw

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        out, bw = one_step(False)
    barrier()
    st = None
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(a.steps):
        out, bw = one_step(True)
        # per-kernel durations are read after the loop for the last step only (events are reused);
        # reading them here would add a host sync to the timed region.
    t_end.record()
    barrier()
    ms_total = t_start.elapsed_time(t_end)
    clocks = sampler.stop()
    fwd_call_ms = ev[0].elapsed_time(ev[1])
    bwd_call_ms = ev[2].elapsed_time(ev[3])
    kt = core.kernel_times()           # CUDA events recorded inside the C ABI around each kernel of the last step
    fwd_ms = kt["fwd_kernel"]
    nan = bool(torch.isnan(bw["df_du"] if grad else out["q_traj"]).any().item())
    ms_t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_total = float(ms_t.item())
    env_steps = B * T * world * a.steps
    value = env_steps / (ms_total * 1e-3)
    touched = int((out["tactile"].abs().amax(dim=2) > 0).sum().item())    # env-steps whose pad touches (after timing)
    from tactilesimulation_b200 import _lib as tsim_lib
    fp64_peak = tsim_lib.fp64_peak(local_rank) if rank == 0 else None
    del out, bw
    torch.cuda.empty_cache()

    # ---- e2e: public plugin API with pinned host buffers in, gradients + loss out, every step
    sim = Simulation(scene_from_blob(g["ibuf"], g["dbuf"]), batch=B, device=dev, lanes=lanes)
    pin = lambda x: torch.tensor(x).pin_memory()
    q0_p, qd0_p, u_p = pin(q0_h), pin(qd0_h), pin(u_h)
    masks = torch.zeros(T, dtype=torch.bool)
    masks[frames] = True
    g_u = torch.empty((T, B, nu), dtype=torch.float64).pin_memory()
    g_q0 = torch.empty((B, n), dtype=torch.float64).pin_memory()
    loss_h = torch.empty((), dtype=torch.float64).pin_memory()
    q_fin = torch.empty((B, n), dtype=torch.float64).pin_memory()

    def e2e_step():
        if not grad:
            # forward-only rollout through the plugin API: host state / actions in, final state + a tactile metric out
            qs, vs, tacs = EpisodicSimFunction.apply(q0_p.to(dev, non_blocking=True), qd0_p.to(dev, non_blocking=True),
                                                     u_p.to(dev, non_blocking=True), masks, sim, False)
            q_fin.copy_(qs[-1], non_blocking=True)
            loss_h.copy_(tacs.abs().sum(), non_blocking=True)
            torch.cuda.synchronize()
            return
        q0d = q0_p.to(dev, non_blocking=True).requires_grad_(True)
        qd0d = qd0_p.to(dev, non_blocking=True).requires_grad_(True)
        ud = u_p.to(dev, non_blocking=True).requires_grad_(True)
        qs, vs, tacs = EpisodicSimFunction.apply(q0d, qd0d, ud, masks, sim, True)
        if name == "push":
            r_pos = (((qs[:, :, 3:5] - goal[None, :, 0:2]) / 0.01) ** 2).sum() * 0.01
            r_rot = (((qs[:, :, 6] - goal[None, :, 2]) / (np.pi / 36.0)) ** 2).sum() * 0.1
            r_touch = ((vs[:, :, 0:3] - vs[:, :, 3:6]) ** 2).sum() / (0.02 ** 2)
            loss = (r_pos + r_rot + r_touch) * scale + 1e-3 * tacs.sum()
        else:
            loss = (-(qs[:, :, 9]).sum() + (vs ** 2).sum()) * scale + 1e-3 * tacs.sum()
        loss.backward()
        gradbuf[:nu] = ud.grad.sum(dim=(0, 1))
        if world > 1:
            dist.all_reduce(gradbuf)
        g_u.copy_(ud.grad, non_blocking=True)
        g_q0.copy_(q0d.grad, non_blocking=True)
        loss_h.copy_(loss.detach(), non_blocking=True)
        torch.cuda.synchronize()
        sim.clearBackwardCache()

    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    for _ in range(min(a.warmup, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    t_start.record()
    for _ in range(e2e_steps):
        e2e_step()
    t_end.record()
    barrier()
    e2e_ms = t_start.elapsed_time(t_end)
    e2e_wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = max(e2e_ms, e2e_wall)
    ms_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    e2e_value = B * T * world * e2e_steps / (float(ms_t.item()) * 1e-3)
    h2d = q0_p.numel() * 8 + qd0_p.numel() * 8 + u_p.numel() * 8
    d2h = (g_u.numel() * 8 + g_q0.numel() * 8 + 8) if grad else (q_fin.numel() * 8 + 8)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (forward): algorithmic bytes per env-step, SURVEY.md 8(d)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    else:
        peak, peak_src = 6650.0, "fallback"
    M = core.n_markers
    items = B * T
    # algorithmic HBM bytes per env-step of each kernel: the step loop moves u in, q / qd / var and
    # the H block of the tape out; G0 / G1 / gains are written by tape_kernel.  The two passes that stream are credited
    # with the bytes they MOVE: the tactile pass zeroes the whole field (cudaMemsetAsync, inside its timer) and re-writes
    # the env-steps whose pad touches; the pull-back pass reads the tactile cotangent of the touched env-steps only.
    per_step = {"fwd_kernel": 8 * nu + 16 * n + 8 * nvar + (8 * n * n if grad else 0),
                "tape_kernel": 8 * nu + 32 * n + 16 * n * n + 8 * nu,
                "bwd_kernel": 8 * n + 16 * n + 8 * (3 * n * n + nu) + 8 * nu}
    launch_bytes = {k: v * items for k, v in per_step.items()}
    launch_bytes["tac_kernel"] = items * 16 * n + len(frames) * B * 24 * M + touched * 24 * M
    launch_bytes["vjp_kernel"] = items * (16 * n + 8 * nvar + 16 * n) + touched * 24 * M
    fwd_bytes = launch_bytes["fwd_kernel"]
    achieved = fwd_bytes / (fwd_ms * 1e-3) / 1e9
    kernels = {k: {"ms": kt[k], "algorithmic_bytes_per_launch": launch_bytes[k],
                   "achieved_gbs": launch_bytes[k] / (kt[k] * 1e-3) / 1e9, "frac": launch_bytes[k] / (kt[k] * 1e-3) / 1e9 / peak}
               for k in launch_bytes if kt.get(k)}
    if "tac_kernel" in kernels:
        kernels["tac_kernel"]["note"] = "memset of the field + env-steps in touch re-written (%d of %d)" % (touched, len(frames) * B)
    if "vjp_kernel" in kernels:
        kernels["vjp_kernel"]["note"] = "tactile cotangent read for the env-steps in touch only (%d of %d)" % (touched, items)
    # figures captured under ncu (tools/gpu_ncu_r02.sh) are quoted only for the kernel sources they were captured on
    from tactilesimulation_b200.build import source_sha
    traffic, flops, cap_note = None, None, "no ncu capture for these kernel sources (profiles/r02_counters.json)"
    cpath = os.path.join(ROOT, "profiles", "r02_counters.json")
    if os.path.exists(cpath):
        try:
            cap = json.load(open(cpath))
            if cap.get("source_sha") == source_sha() and cap.get("B") == B and cap.get("T") == T and cap.get("workload", "push") == name:
                traffic = cap.get("fwd_kernel_dram_bytes_per_launch")
                flops = cap.get("fwd_kernel_fp64_flops_per_launch")
                cap_note = "ncu capture of this source (sha %s), B=%d T=%d" % (cap["source_sha"], B, T)
        except Exception:
            pass
    fp64 = {"peak_gflops": fp64_peak["gflops"], "peak_source": "tsim_debug_fp64_peak (DFMA chains, %d SMs, best of 3)" % fp64_peak["sms"],
            "flops_per_launch": flops, "achieved_gflops": (flops / (fwd_ms * 1e-3) / 1e9) if flops else None,
            "frac": (flops / (fwd_ms * 1e-3) / 1e9 / fp64_peak["gflops"]) if flops else None,
            "flops_source": "smsp__sass_thread_inst_executed_op_{dfma x2, dmul, dadd}_pred_on of fwd_kernel; " + cap_note}
    roofline = {"bound": "hbm", "kernel": "fwd_kernel", "achieved": achieved, "peak": peak, "peak_source": peak_src,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": cap_note,
                "algorithmic_bytes_per_launch": fwd_bytes, "kernel_ms": fwd_ms,
                "forward_call_ms": fwd_call_ms, "adjoint_call_ms": bwd_call_ms, "kernels": kernels, "fp64": fp64,
                "note": "fwd_kernel (the step loop) is fp64 latency / issue bound: its HBM fraction is small by construction "
                        "(SURVEY.md 8d), roofline.fp64 is the roofline that binds; the passes that stream are tac_kernel / vjp_kernel"}

    # ---- CPU baseline beside it: unmodified reference C++ on the host cores, bounded sample
    cpu = None
    if world == 1 and ref_available() and not a.no_cpu_baseline:       # reported at N=1 only (rank 0)
        import multiprocessing as mp
        cores = host_cores()
        with mp.get_context("fork").Pool(cores) as pool:
            ref_pass(pool, cores, 1, 10, 777, name)
            s, w = ref_pass(pool, cores, a.ref_envs_per_core, T, 0, name)
        cpu = {"value": s / w, "unit": "env-steps/s", "cores": cores, "kind": "reference",
               "sample": f"{cores} processes x {a.ref_envs_per_core} env x T={T} sim-steps, "
                         + ("fwd+tape+backward" if grad else "forward only") + " (oracle/_ref/redmax_py, -O3 -DNDEBUG)"}

    stream_gb = (2 if grad else 1) * len(frames) * B * ntac * 8 / 1e9 + (T * B * core.tape_doubles * 8 / 1e9 if grad else 0.0)
    line = {"metric": wl["metric"], "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["label"].format(T=T, B=B),
                       "batch_per_gpu": B, "horizon": T, "markers": M, "lanes_per_env": core.lanes or (8 if n <= 8 else 16),
                       "l2": ("inputs_larger_than_l2" if stream_gb > 0.3 else "l2_flushed_by_the_step's_own_streams")
                             + " (tactile field%s: %.2f GB written%s per step)" % (" + cotangent + tape" if grad else "", stream_gb, " / read" if grad else ""),
                       "collective": ("1 NCCL all-reduce of %d f64 per step" % POLICY_GRAD_SIZE if (world > 1 and grad) else
                                      ("none (forward-only rollout, env shards are independent)" if world > 1 else "none (1 GPU)"))},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": ("EpisodicSimFunction.apply + loss.backward(), pinned host q0/qdot0/actions in, grads + loss out" if grad
                                                else "EpisodicSimFunction.apply (grad_mode False), pinned host q0/qdot0/actions in, final state + tactile metric out")},
            # launches of this library's kernels inside the timed region (vjp_kernel runs two phases per step)
            "gpu_launches": sum((2 if k == "vjp_kernel" else 1) for k, v in kt.items() if v) * a.steps,
            "roofline": roofline, "cpu_baseline": cpu, "nan": nan}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_stepsim(a):
    """The gd.py shape (gym-steps/s): B TactilePush 13x10 environments through the batched front-end
    (tactilesimulation_b200.envs.BatchedTactilePushEnv over the batched StepSimFunction): reset, `gym_steps` gym steps of
    frame_skip 5 sim-steps each (one forward launch per gym step, tactile / variables on the last sub-step), loss = -mean
    episode reward, loss.backward() (one backward_steps launch chain per gym step).  One bench step = one episode of the
    whole batch.  value: device-resident actions; e2e: per gym step the actions come from pinned host memory and the
    observation + reward go back to the host (the policy-on-host loop of gd.py)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from tactilesimulation_b200.envs import BatchedTactilePushEnv
    from tactilesimulation_b200.layout import scene_from_blob
    from tactilesimulation_b200.redmax import Simulation
    from tactilesimulation_b200.distributed import rank_seed
    g = np.load(os.path.join(GOLDEN, "pusher13x10_episodic_s0.npz"))
    B, G = a.batch or 4096, a.gym_steps
    sim = Simulation(scene_from_blob(g["ibuf"], g["dbuf"]), batch=B, device=dev, lanes=a.lanes)
    sim.set_q_init(np.tile(g["q0"], (B, 1)))
    env = BatchedTactilePushEnv(sim, observation_type="tactile_flatten", gradient=True, tactile_rows=13, tactile_cols=10,
                                seed=rank_seed(1234, rank))
    gen = torch.Generator(device=dev).manual_seed(rank_seed(99, rank))
    acts = torch.randn((G, B, 3), generator=gen, device=dev, dtype=torch.float64)
    acts_p = acts.cpu().pin_memory()
    gradbuf = torch.zeros(POLICY_GRAD_SIZE, dtype=torch.float64, device=dev)
    obs_h = torch.empty((B, 393), dtype=torch.float64).pin_memory()
    rew_h = torch.empty((B,), dtype=torch.float64).pin_memory()
    launches = [0]

    def episode(host_loop):
        env.reset()
        total, leaves = 0.0, []
        for k in range(G):
            a_k = (acts_p[k].to(dev, non_blocking=True) if host_loop else acts[k]).requires_grad_(True)
            obs, r, done, info = env.step(a_k)
            if host_loop:
                obs_h.copy_(obs.detach(), non_blocking=True)
                rew_h.copy_(r.detach(), non_blocking=True)
                torch.cuda.synchronize()
            total = total + r
            leaves.append(a_k)
        loss = -(total.mean()) / world
        loss.backward()
        gradbuf[:3] = torch.stack([x.grad for x in leaves]).sum(dim=(0, 1))
        if world > 1:
            dist.all_reduce(gradbuf)
        sim.clearBackwardCache()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 1)):
        episode(False)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0e.record()
    for _ in range(a.steps):
        loss = episode(False)
    t1e.record()
    barrier()
    ms = t0e.elapsed_time(t1e)
    clocks = sampler.stop()
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = float(ms_t.item())
    value = B * G * world * a.steps / (ms * 1e-3)
    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    episode(True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        episode(True)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    ms_t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    e2e_value = B * G * world * e2e_steps / (float(ms_t.item()) * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and ref_available() and not a.no_cpu_baseline:
        import multiprocessing as mp
        cores = host_cores()
        # (spawn: the workers run torch autograd, which cannot be used in a fork of a process that already has)
        with mp.get_context("spawn").Pool(cores) as pool:
            pool.map(_ref_stepsim_worker, [([900 + c], 2) for c in range(cores)])      # imports + first call
            t0 = time.perf_counter()
            res = pool.map(_ref_stepsim_worker, [([c], G) for c in range(cores)])
            w = time.perf_counter() - t0
        cpu = {"value": sum(r[0] for r in res) / w, "unit": "gym-steps/s", "cores": cores, "kind": "reference",
               "sample": f"{cores} processes x 1 episode x {G} gym steps: the unmodified tactile_push_env.py + StepSimFunction on oracle/_ref/redmax_py, fwd + loss.backward()"}
    line = {"metric": STEPSIM_METRIC, "value": value, "unit": "gym-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"TactilePush 13x10 gd.py shape: {G} gym steps x frame_skip 5 per episode, fwd + loss.backward(), batch {B}/GPU",
                       "batch_per_gpu": B, "gym_steps": G, "frame_skip": 5, "env_steps_per_s": value * 5,
                       "l2": "l2_flushed_by_the_episode's_own_tape (%.2f GB per episode)" % (G * 5 * B * sim.core.tape_doubles * 8 / 1e9)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "gym-steps/s", "h2d_bytes_per_step": G * B * 3 * 8, "d2h_bytes_per_step": G * (B * 393 + B) * 8,
                    "steps": e2e_steps, "api": "BatchedTactilePushEnv.step per gym step: actions from pinned host memory, observation + reward to the host, loss.backward()"},
            "gpu_launches": a.steps * G * 5, "cpu_baseline": cpu, "nan": bool(torch.isnan(loss).item())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="push", choices=sorted(WORKLOADS) + ["stepsim"],
                    help="push (default, the headline BASELINE configs[2]) | push_fwd | dclaw | insertion | stepsim")
    ap.add_argument("--batch", type=int, default=0, help="environments per GPU (0 = the workload's own: 4096 for push)")
    ap.add_argument("--horizon", type=int, default=0, help="sim-steps per episode (0 = the workload's own: 200 for push)")
    ap.add_argument("--gym-steps", type=int, default=100, help="stepsim: gym steps per episode (R/envs/__init__.py:9-13)")
    ap.add_argument("--lanes", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-envs-per-core", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        if not ref_available():
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/redmax_py not built (run oracle/build_ref.sh where /root/reference exists)"}))
            return
        run_reference(a)
    elif a.workload == "stepsim":
        run_stepsim(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
