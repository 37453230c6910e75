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
    """One process = one core: E environments x T steps, forward(+tape) and backward() through the
    unmodified reference redmax_py (EpisodicSimFunction call pattern, redmax_torch_functions.py:35-109)."""
    xml, seeds, T = args
    sys.path.insert(0, REF_DIR)
    import redmax_py
    sim = redmax_py.Simulation(xml)
    n, nv, nt, nu = sim.ndof_r, sim.ndof_var, sim.ndof_tactile, sim.ndof_u
    q_init = np.array(sim.get_q_init())
    t0 = time.perf_counter()
    steps = 0
    for seed in seeds:
        q0, qd0, u, goal = make_inputs(q_init, 1, T, seed)
        sim.set_state_init(q0[0], qd0[0])
        sim.reset(True)
        qs, vs = np.zeros((T, n)), np.zeros((T, nv))
        for t in range(T):
            sim.set_u(u[t, 0])
            sim.forward(1)
            qs[t] = sim.get_q()
            vs[t] = sim.get_variables()
            sim.get_tactile_force_vector()
        dq = np.zeros((T, n))
        dq[:, 3:5] = 2.0 * (qs[:, 3:5] - goal[0, 0:2]) / (0.01 ** 2) * 0.01
        dq[:, 6] = 2.0 * (qs[:, 6] - goal[0, 2]) / ((np.pi / 36.0) ** 2) * 0.1
        d = 2.0 * (vs[:, 0:3] - vs[:, 3:6]) / (0.02 ** 2)
        bi = sim.backward_info
        bi.set_flags(True, True, False, True)
        bi.df_dq = dq.reshape(-1)
        bi.df_dvar = np.concatenate([d, -d], axis=1).reshape(-1)
        bi.df_dtactile = np.full(nt * T, 1e-3)
        bi.df_dq0, bi.df_dqdot0, bi.df_du = np.zeros(n), np.zeros(n), np.zeros(nu * T)
        sim.backward()
        steps += T
    return steps, time.perf_counter() - t0


def ref_pass(pool, cores, envs_per_core, T, seed0):
    """One bounded CPU sample: cores x envs_per_core environments x T steps.  Returns (env-steps, wall s)."""
    jobs = [(REF_XML, [seed0 + c * envs_per_core + e for e in range(envs_per_core)], T) for c in range(cores)]
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


def run_reference(a):
    """--impl reference: the reference's own CPU implementation on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    sys.path.insert(0, REF_DIR)
    import redmax_py  # noqa: F401  (loaded once in the parent: the forked workers inherit it, the driver's hook sees the .so)
    cores = host_cores()
    T = a.horizon
    envs_per_core = a.ref_envs_per_core
    sample = f"{cores} processes x {envs_per_core} env x T={T} sim-steps per step (fwd+tape+backward, tactile read every step)"
    with mp.get_context("fork").Pool(cores) as pool:
        for w in range(a.warmup):
            ref_pass(pool, cores, 1, min(T, 20), 10_000 + w)
        tot_steps, tot_wall = 0, 0.0
        for k in range(a.steps):
            s, w = ref_pass(pool, cores, envs_per_core, T, 1000 * k)
            tot_steps += s
            tot_wall += w
    val = tot_steps / tot_wall
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "env-steps/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * tot_wall / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"TactilePush 32x13 fwd+adjoint horizon {T}, bounded sample of the batch-{a.batch} job",
                       "horizon": T, "markers": N_MARKERS},
            "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": cores, "kind": "reference", "sample": sample,
                             "build": "oracle/build_ref.sh: g++ -O3 -DNDEBUG (the reference's CMake Release flags)"},
            "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    g = np.load(os.path.join(GOLDEN, SCENE_CASE + ".npz"))
    B, T = a.batch, a.horizon
    core = BatchedSim((g["ibuf"], g["dbuf"]), device=dev, lanes=a.lanes)
    n, nu, nvar, ntac = core.ndof_r, core.ndof_u, core.ndof_var, core.ndof_tactile
    from tactilesimulation_b200.distributed import rank_seed
    # weak scaling: every rank owns B environments of the global batch [rank*B, (rank+1)*B), own RNG stream
    q0_h, qd0_h, u_h, goal_h = make_inputs(g["q0"], B, T, seed=rank_seed(1234, rank))
    q0, qd0 = torch.tensor(q0_h, device=dev), torch.tensor(qd0_h, device=dev)
    u, goal = torch.tensor(u_h, device=dev), torch.tensor(goal_h, device=dev)
    dtac = torch.full((T, B, ntac), 1e-3, dtype=torch.float64, device=dev)   # tactile cotangent, resident
    gradbuf = torch.zeros(POLICY_GRAD_SIZE, dtype=torch.float64, device=dev)
    scale = 1.0 / (B * world)
    # preallocated outputs are reused by every step (180 GB HBM: tactile [T,B,3M] is 8.2 GB at B=4096)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    fwd_ms, bwd_ms = [], []

    def one_step(timed):
        q, qd = q0.clone(), qd0.clone()
        if timed:
            ev[0].record()
        out = core.forward(q, qd, u, T, grad=True)
        if timed:
            ev[1].record()
        dq, dv = reward_cotangents(out["q_traj"], out["var"], goal, scale)
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
    nan = bool(torch.isnan(bw["df_du"]).any().item())
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
    sim = Simulation(scene_from_blob(g["ibuf"], g["dbuf"]), batch=B, device=dev, lanes=a.lanes)
    pin = lambda x: torch.tensor(x).pin_memory()
    q0_p, qd0_p, u_p = pin(q0_h), pin(qd0_h), pin(u_h)
    masks = torch.ones(T, dtype=torch.bool)
    g_u = torch.empty((T, B, nu), dtype=torch.float64).pin_memory()
    g_q0 = torch.empty((B, n), dtype=torch.float64).pin_memory()
    loss_h = torch.empty((), dtype=torch.float64).pin_memory()

    def e2e_step():
        q0d = q0_p.to(dev, non_blocking=True).requires_grad_(True)
        qd0d = qd0_p.to(dev, non_blocking=True).requires_grad_(True)
        ud = u_p.to(dev, non_blocking=True).requires_grad_(True)
        qs, vs, tacs = EpisodicSimFunction.apply(q0d, qd0d, ud, masks, sim, True)
        r_pos = (((qs[:, :, 3:5] - goal[None, :, 0:2]) / 0.01) ** 2).sum() * 0.01
        r_rot = (((qs[:, :, 6] - goal[None, :, 2]) / (np.pi / 36.0)) ** 2).sum() * 0.1
        r_touch = ((vs[:, :, 0:3] - vs[:, :, 3:6]) ** 2).sum() / (0.02 ** 2)
        loss = (r_pos + r_rot + r_touch) * scale + 1e-3 * tacs.sum()
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
    d2h = g_u.numel() * 8 + g_q0.numel() * 8 + 8

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
    per_step = {"fwd_kernel": 8 * nu + 16 * n + 8 * nvar + 8 * n * n,
                "tape_kernel": 8 * nu + 32 * n + 16 * n * n + 8 * nu,
                "bwd_kernel": 8 * n + 16 * n + 8 * (3 * n * n + nu) + 8 * nu}
    launch_bytes = {k: v * items for k, v in per_step.items()}
    launch_bytes["tac_kernel"] = items * (16 * n + 24 * M) + touched * 24 * M
    launch_bytes["vjp_kernel"] = items * (16 * n + 8 * nvar + 16 * n) + touched * 24 * M
    fwd_bytes = launch_bytes["fwd_kernel"]
    achieved = fwd_bytes / (fwd_ms * 1e-3) / 1e9
    kernels = {k: {"ms": kt[k], "algorithmic_bytes_per_launch": launch_bytes[k],
                   "achieved_gbs": launch_bytes[k] / (kt[k] * 1e-3) / 1e9, "frac": launch_bytes[k] / (kt[k] * 1e-3) / 1e9 / peak}
               for k in launch_bytes if kt.get(k)}
    kernels["tac_kernel"]["note"] = "memset of the field + env-steps in touch re-written (%d of %d)" % (touched, items)
    kernels["vjp_kernel"]["note"] = "tactile cotangent read for the env-steps in touch only (%d of %d)" % (touched, items)
    # figures captured under ncu (tools/gpu_ncu_r02.sh) are quoted only for the kernel sources they were captured on
    from tactilesimulation_b200.build import source_sha
    traffic, flops, cap_note = None, None, "no ncu capture for these kernel sources (profiles/r02_counters.json)"
    cpath = os.path.join(ROOT, "profiles", "r02_counters.json")
    if os.path.exists(cpath):
        try:
            cap = json.load(open(cpath))
            if cap.get("source_sha") == source_sha() and cap.get("B") == B and cap.get("T") == T:
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
            ref_pass(pool, cores, 1, 10, 777)
            s, w = ref_pass(pool, cores, a.ref_envs_per_core, T, 0)
        cpu = {"value": s / w, "unit": "env-steps/s", "cores": cores, "kind": "reference",
               "sample": f"{cores} processes x {a.ref_envs_per_core} env x T={T} sim-steps, fwd+tape+backward (oracle/_ref/redmax_py)"}

    line = {"metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"TactilePush 32x13 fwd+adjoint horizon {T}, batch {B}/GPU (BASELINE configs[2])",
                       "batch_per_gpu": B, "horizon": T, "markers": M, "lanes_per_env": a.lanes,
                       "l2": "inputs_larger_than_l2 (tactile field + cotangent: %.1f GB per step)" % (2 * T * B * ntac * 8 / 1e9),
                       "collective": "1 NCCL all-reduce of %d f64 per step" % POLICY_GRAD_SIZE if world > 1 else "none (1 GPU)"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "EpisodicSimFunction.apply + loss.backward(), pinned host q0/qdot0/actions in, grads + loss out"},
            # launches of this library's kernels inside the timed region (vjp_kernel runs two phases per step)
            "gpu_launches": sum((2 if k == "vjp_kernel" else 1) for k, v in kt.items() if v) * a.steps,
            "roofline": roofline, "cpu_baseline": cpu, "nan": nan}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--horizon", type=int, default=200, help="sim-steps per episode")
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
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
