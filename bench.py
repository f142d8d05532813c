"""Headline benchmark: safeguarded env-steps/s (forward + backward) of the SHAC rollout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--env ...] [--safeguard ...]

One "step" = one SHAC episode pass over the batch: buffer.reset, collect_trajectories (H safeguarded
environment steps for every environment, policy + target critic inside) and update_policy (loss,
backward through the whole horizon, clip, Adam) -- src/learning_algorithms/shac.py:87-90,96-151 of the
reference; the critic fit is excluded.  Workload at N=1: BASELINE.json configs[1],
BalanceCartPole + boundary projection, 4096 environments, H = 64, policy [128,128], critic [128,128,128].
Multi-GPU: environments are sharded (4096 per rank, weak scaling); the only collective is the NCCL
all-reduce of the policy gradients (+ 4 scalars).
Prints ONE JSON line (see the contract in the task statement / DESIGN.md §Measurement).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ENV_DEFAULT, SG_DEFAULT = "BalanceCartPole", "boundary_projection"
N_PER_GPU, HORIZON = 4096, 64
POLICY_ARCH, VF_ARCH = [128, 128], [128, 128, 128]
ALG_BYTES = {"BalancePendulum": 105, "BalanceCartPole": 177, "SwingUpCartPole": 177, "BalanceQuadrotor": 273,
             "ManageHousehold": 309, "NavigateSeeker": 193}     # 4(7n+4m+2o+2)+1, fp32


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--env", default=ENV_DEFAULT)
    p.add_argument("--safeguard", default=SG_DEFAULT, choices=["none", "boundary_projection", "ray_mask"])
    p.add_argument("--num-envs", type=int, default=N_PER_GPU, help="environments per GPU")
    p.add_argument("--horizon", type=int, default=HORIZON)
    p.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    p.add_argument("--fused", default="auto", choices=["auto", "on", "off"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--policy-arch", default=None, help="comma-separated hidden widths (default 128,128)")
    p.add_argument("--vf-arch", default=None, help="comma-separated hidden widths (default 128,128,128)")
    p.add_argument("--gate", default="intended", choices=["reference", "intended", "always"],
                   help="intended: the guard acts where the linearised reachable set meets the feasible box (safeguard.py:63-67 as the "
                        "paper describes it); reference: the shipped inverted where() (Q1), under which the guard fires only where the "
                        "program is infeasible -- the reference itself raises there, so it is not a benchmarkable steady state")
    p.add_argument("--rng-order", default="batched", choices=["batched", "reference"],
                   help="batched: 3 RNG draws per episode (same distributions); reference: the reference's per-step order")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ------------------------------------------------------------------------------------------------------
CPU_ENVS_PER_WORKER = 64          # per-process N in {64, 256, 1024}; the CPU path is linear in N
STATE_SPREAD = 0.9                # both arms: initial states uniform in 0.9 x the feasible state box


def _cpu_persistent_worker(conn, job):
    """One host process = one environment shard of the oracle port; runs one episode per "step" message."""
    env_name, sg_kind, n_envs, horizon, seed, gate = job
    import torch
    torch.set_num_threads(1)
    from oracle.dynamics import OracleEnv
    from oracle.safeguard import OracleSafeguard
    from oracle import shac as oshac
    torch.manual_seed(seed)
    env = OracleEnv(env_name, n_envs, 240)
    env.reset(seed=seed)
    # the safeguard bakes in the linearisation of environment 0 at CONSTRUCTION (Q4, safeguard.py:54).  The reference constructs
    # it on an un-reset environment (torch.empty state); both arms here construct it at the zero state
    env.state = torch.zeros_like(env.state)
    step_fn = env if sg_kind == "none" else OracleSafeguard(env, sg_kind, gate=gate, strict=False)
    obs_dim = env.observation().shape[1]
    pol = oshac.OraclePolicy(obs_dim, env.action_dim, POLICY_ARCH).double()
    vf = oshac.OracleValue(obs_dim, VF_ARCH).double()
    opt = torch.optim.Adam(pol.parameters(), lr=4e-4)
    counts = {"infeasible": 0, "nonfinite": 0}

    def env_step(action, t):
        obs, rew, term, trunc = step_fn.step(action)
        if step_fn is not env:              # infeasible programs whose solution the gate actually uses
            used = {"reference": ~step_fn.projectable, "intended": step_fn.projectable}.get(gate, torch.ones_like(step_fn.projectable))
            counts["infeasible"] += int((step_fn.infeasible & used).sum())
        counts["nonfinite"] += int((~torch.isfinite(obs).all(dim=1)).sum())
        return obs, rew, term | trunc

    conn.send("ready")
    k = 0
    while conn.recv() == "step":
        # every step (= one SHAC episode) starts from a fresh sample of the same initial-state distribution as the GPU arm
        k += 1
        env.reset(seed=seed + 1000 * k)
        env.state = env.state_center + (torch.rand(n_envs, env.state_dim, dtype=torch.float64) * 2 - 1) * env.state_half * STATE_SPREAD
        env.cut_computation_graph()
        opt.zero_grad()
        eps = torch.randn(horizon, n_envs, env.action_dim, dtype=torch.float64)
        counts["infeasible"] = counts["nonfinite"] = 0
        try:
            out = oshac.rollout_loss(env_step, env.observation(), pol, vf, horizon, eps)
            out["loss"].backward()
            if bool(torch.isfinite(out["loss"])):        # (a non-finite episode is counted below and must not poison the weights)
                torch.nn.utils.clip_grad_norm_(pol.parameters(), 1.0)
                opt.step()
            conn.send(("done", counts["infeasible"], counts["nonfinite"]))
        except Exception as exc:        # a failed worker-episode is REPORTED (worker_errors) and its work is not counted
            conn.send(("error", repr(exc)[:200]))


def cpu_port_persistent(env_name, sg_kind, horizon, steps, warmup, budget_s, n_w=CPU_ENVS_PER_WORKER, gate="intended"):
    """The reference's CPU path (oracle port: per-environment HiGHS solves + torch autograd, float64) on all host cores.
    A step = one SHAC episode (collect + backward + clip + Adam) of a BOUNDED sample of the workload: every core owns
    `n_w` environments (process start-up and imports are outside the timed steps).  The first warm-up step also
    calibrates the step time; warm-up and timed steps are capped so that the run fits `budget_s`.
    Returns dict(value, steps, warmup, total_s, workers, n_w, errors, infeasible_env_steps, nonfinite_env_steps)."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    ctx = mp.get_context("spawn")
    procs = []
    for i in range(workers):
        a, b = ctx.Pipe()
        pr = ctx.Process(target=_cpu_persistent_worker, args=(b, (env_name, sg_kind, n_w, horizon, 100 + i, gate)), daemon=True)
        pr.start()
        procs.append((pr, a))
    for _, c in procs:
        c.recv()
    errors, tally = [], {"failed_workers": 0, "infeasible": 0, "nonfinite": 0}

    def step():
        t0 = time.perf_counter()
        for _, c in procs:
            c.send("step")
        for _, c in procs:
            msg = c.recv()
            if msg[0] == "done":
                tally["infeasible"] += msg[1]
                tally["nonfinite"] += msg[2]
            else:
                errors.append(msg[1])
                tally["failed_workers"] += 1
        return time.perf_counter() - t0

    calib = step()                                   # the first warm-up step
    warmup = max(1, min(warmup, int(0.35 * budget_s / max(calib, 1e-3))))
    for _ in range(warmup - 1):
        step()
    steps = max(1, min(steps, int((budget_s - warmup * calib) / max(calib, 1e-3))))
    errors.clear()
    for key in tally:
        tally[key] = 0
    walls = [step() for _ in range(steps)]
    for pr, c in procs:
        c.send("stop")
    for pr, _ in procs:
        pr.join(timeout=10)
    total = sum(walls)
    done_env_steps = (workers * steps - tally["failed_workers"]) * horizon * n_w
    return dict(value=done_env_steps / total, steps=steps, warmup=warmup, total_s=total, workers=workers, n_w=n_w, errors=errors,
                infeasible_env_steps=tally["infeasible"], nonfinite_env_steps=tally["nonfinite"],
                ms_per_env_step_per_core=1e3 * total * workers / max(done_env_steps, 1))


def _cpu_config(args, r):
    """What the CPU arm actually ran (the GPU arm's config keys + the documented differences)."""
    cfg = workload_config(args, max(1, args.gpus))
    cfg.update({"num_envs_cpu": r["workers"] * r["n_w"], "envs_per_worker": r["n_w"], "cpu_workers": r["workers"],
                "cpu_note": "the CPU path is linear in the number of environments (BASELINE.md section 3): the arm runs "
                            f"{r['workers']} processes x {r['n_w']} environments and the GPU/CPU ratio is per env-step; float64 like the reference",
                "initial_states": f"uniform in {STATE_SPREAD} x feasible state box, fresh sample every step (both arms)"})
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_port_persistent(args.env, args.safeguard, args.horizon, args.steps, max(3, args.warmup), budget_s=200.0, gate=args.gate)
    v = r["value"]
    per_n = {str(r["n_w"]): {"env_steps_per_s": v, "ms_per_env_step_per_core": r["ms_per_env_step_per_core"]}}
    if os.environ.get("SGBPO_REF_SWEEP"):            # optional: the other per-process sizes of BASELINE.md section 3 (slow)
        for n_w in (256,):
            q = cpu_port_persistent(args.env, args.safeguard, args.horizon, 1, 1, budget_s=1.0, n_w=n_w, gate=args.gate)
            per_n[str(n_w)] = {"env_steps_per_s": q["value"], "ms_per_env_step_per_core": q["ms_per_env_step_per_core"]}
    line = {
        "impl": "reference", "metric": "safeguarded env-steps/s (fwd+bwd)", "value": v, "unit": "env-steps/s",
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": r["warmup"], "steps_requested": args.steps,
        "warmup_requested": args.warmup, "ms_per_step": 1e3 * r["total_s"] / r["steps"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _cpu_config(args, r),
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": r["workers"], "kind": "port",
                         "sample": f"each step = {r['workers']} processes x {r['n_w']} envs x H={args.horizon}, one SHAC episode of the "
                                   "same workload (oracle port of the reference path: per-env HiGHS solves + torch autograd, "
                                   "float64); warm-up and timed steps capped to a 200 s budget",
                         "ms_per_env_step_per_core": r["ms_per_env_step_per_core"], "per_envs_per_worker": per_n},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "worker_errors": {"count": len(r["errors"]), "first": r["errors"][0] if r["errors"] else None},
        "gates": {"infeasible_env_steps": r["infeasible_env_steps"], "nonfinite_env_steps": r["nonfinite_env_steps"],
                  "env_steps": r["workers"] * r["n_w"] * args.horizon * r["steps"]},
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": f"{args.env}+{args.safeguard}, {args.num_envs} envs/GPU x {world} GPU, H={args.horizon}, "
                        f"policy {POLICY_ARCH}, critic {VF_ARCH}, SHAC collect_trajectories+update_policy",
            "env": args.env, "safeguard": args.safeguard, "num_envs_per_gpu": args.num_envs, "horizon": args.horizon,
            "gate": args.gate, "rng_order": args.rng_order,
            "initial_states": f"uniform in {STATE_SPREAD} x feasible state box, fresh sample every step (both arms)", "l2_policy": "working set per episode > L2 not guaranteed; "
            "a 256 MiB buffer is written between timed episodes to flush L2"}


# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def _nvml_loop(self) -> bool:
        """The counters nvidia-smi prints, read in-process through NVML (nvidia_ml_py): spawning nvidia-smi every 50 ms enumerates
        every GPU of the box and takes driver locks, which stalls rank 0's launches -- and, through the gradient all-reduce, every
        other rank's timed episode -- for milliseconds.  Returns False when NVML is not usable (then nvidia-smi is polled)."""
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            return False
        self.source = "nvml"
        while not self.stop_flag:
            try:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                get = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = int(get(handle))
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)
        return True

    def run(self):
        self.source = "nvidia-smi"
        if self._nvml_loop():
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, val in zip(names, out[2:]):
                    if "Active" in val and "Not" not in val:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "source": getattr(self, "source", "nvidia-smi")}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from safegbpo_b200 import _lib
    from safegbpo_b200.envs import balance_cartpole, balance_pendulum, swingup_cartpole, balance_quadrotor, manage_household
    from safegbpo_b200.safeguards.boundary_projection import BoundaryProjectionSafeguard
    from safegbpo_b200.safeguards.ray_mask import RayMaskSafeguard
    from safegbpo_b200.learning_algorithms.shac import SHAC

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    torch.set_default_dtype(dtype)
    torch.set_default_device(dev)
    torch.manual_seed(1234 + rank)
    env_cls = {"BalanceCartPole": balance_cartpole.BalanceCartPoleEnv, "BalancePendulum": balance_pendulum.BalancePendulumEnv,
               "SwingUpCartPole": swingup_cartpole.SwingUpCartPoleEnv, "BalanceQuadrotor": balance_quadrotor.BalanceQuadrotorEnv,
               "ManageHousehold": manage_household.ManageHouseholdEnv}.get(args.env)
    N, H = args.num_envs, args.horizon
    if args.env == "NavigateSeeker":
        from safegbpo_b200.envs.navigate_seeker import NavigateSeekerEnv
        env = NavigateSeekerEnv(num_envs=N, num_steps=400, num_obstacles=1, min_radius=2.0, max_radius=4.0)
        env.reset(seed=1)
    else:
        env = env_cls(num_envs=N, num_steps={"ManageHousehold": 720}.get(args.env, 240))
    if args.safeguard == "boundary_projection":
        env = BoundaryProjectionSafeguard(env, gate=args.gate, strict=False)
    elif args.safeguard == "ray_mask":
        env = RayMaskSafeguard(env, linear_projection=True, zonotopic_approximation=True, passthrough=False, gate=args.gate, strict=False)
    algo = SHAC(env=env, policy_kwargs={"net_arch": POLICY_ARCH}, policy_optim_kwargs={"lr": 4e-4},
                vf_kwargs={"net_arch": VF_ARCH}, vf_optim_kwargs={"lr": 2e-3}, regularisation_coefficient=0.1,
                len_trajectories=H, fused=args.fused, rng_order=args.rng_order)
    if world > 1:      # replicated networks: broadcast rank 0's parameters; gradients are summed over ranks
        from safegbpo_b200 import distributed as sgdist
        sgdist.attach(algo)          # broadcast + policy AND critic gradient all-reduce hooks
    base = env.env if hasattr(env, "env") else env
    n = base.state_dim
    # synthetic batched initial states, pinned on the host (e2e leg copies them in every step)
    gen = torch.Generator(device="cpu").manual_seed(99 + rank)
    half = base.state_set.generator[0].diagonal().cpu()
    ctr = base.state_set.center[0].cpu()
    s0_host = (ctr + (torch.rand(N, n, generator=gen, device="cpu", dtype=dtype) * 2 - 1) * half * STATE_SPREAD).pin_memory()
    if hasattr(base, "obstacles"):     # obstacle courses: the states reset() drew together with their goals and obstacles
        s0_host = base.state.detach().to("cpu", dtype).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    launches = {"n": 0}

    def episode(from_host: bool):
        if from_host:      # e2e: the caller hands over host initial states; row 0 of the buffer is rebuilt from them
            base.state = s0_host.to(dev, non_blocking=True)
            base.steps, base._reset_pending = 0, False
            o0 = base.observation
            with torch.no_grad():
                algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
        else:
            algo.buffer.reset()
        algo.policy_optim.zero_grad()
        avg_reward = algo.collect_trajectories()
        loss = algo.update_policy()
        return loss, avg_reward

    def timed(k, from_host):
        """K episodes back to back as a training loop issues them (the only host wait is the episode's own read of its reward
        scalar), bracketed by barrier + synchronize on both sides.  Every episode sits between two CUDA events on the launching
        stream and the L2-flush fill between two episodes sits OUTSIDE them, so the sum of the K intervals is the device time of
        exactly K episodes."""
        events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        for e0, e1 in events:
            flush.fill_(1.0)                          # L2 flush between timed iterations (not timed)
            e0.record()
            episode(from_host)
            e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return [e0.elapsed_time(e1) for e0, e1 in events]

    base.state = s0_host.to(dev)
    base.steps, base._reset_pending = 0, False
    o0 = base.observation
    algo.buffer.reset(o0, algo.target_value_function(o0).squeeze(1))
    n_warm = max(3, args.warmup)
    for _ in range(n_warm):
        episode(False)
    # the whole-horizon engine captures its CUDA graphs over its first episodes (2 eager, one that captures the forward / reverse
    # graphs, one that captures the merged episode graph): a short requested warm-up must not leave a capture inside the timed region
    eng = getattr(algo, "_fused_engine", None)
    while eng is not None and getattr(eng, "use_graphs", False) and eng._episodes < eng.graph_warmup + 3:
        episode(False)
        n_warm += 1
    sampler = ClockSampler(local)      # rank 0 samples its own GPU (one nvidia-smi poller per job, not per rank)
    if rank == 0:
        sampler.start()
    t_dev = timed(args.steps, False)
    t_e2e = timed(args.steps, True)
    extended = False
    t_guard = time.time()
    while True:                 # very short runs: keep the same load going (untimed) until nvidia-smi has sampled it
        need = rank == 0 and len(sampler.samples) < 3 and time.time() - t_guard < 3.0
        if world > 1:           # episodes contain a collective: every rank must run the same number of them
            flag = torch.tensor([1 if need else 0], device=dev, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            need = bool(int(flag))
        if not need:
            break
        episode(False)
        extended = True
    sampler.stop_flag = True
    total_ms = sum(t_dev)
    if world > 1:
        t = torch.tensor([total_ms, sum(t_e2e)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    else:
        e2e_ms = sum(t_e2e)
    in_sync = None
    if world > 1:
        # data parallelism check: one FULL SHAC iteration (untimed: collect, policy update, critic fit with its per-batch gradient
        # all-reduce, Polyak) on every rank, after which policy, critic and target critic must be bit-identical everywhere
        from safegbpo_b200 import distributed as sgdist
        algo._learn_episode(0)
        in_sync = sgdist.replicas_in_sync(algo)
    env_steps = N * H * args.steps * world
    value = env_steps / (total_ms * 1e-3)
    e2e_value = env_steps / (e2e_ms * 1e-3)
    engine = algo._fused_engine
    if engine is not None:
        roof = fused_roofline(args, engine, episode, dtype)
    else:
        roof = step_kernel_roofline(args, base, env, dev, dtype)
    gates = correctness_gates(args, algo, env, base, episode, dtype, dev) if rank == 0 else None
    if world > 1:           # the gate episodes contain the gradient all-reduce: every rank runs them
        if rank != 0:
            for _ in range(3):
                episode(True)
    stress = None
    if rank == 0:
        try:
            stress = step_kernel_roofline(args, base, env, dev, dtype, n_envs=1 << 20)
            stress.pop("launches_per_step", None)
        except Exception as exc:      # the stress variant is additional evidence, never fatal
            stress = {"error": str(exc)[:200]}
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        line = {
            "metric": "safeguarded env-steps/s (fwd+bwd)", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": int(s0_host.numel() * s0_host.element_size()),
                    "d2h_bytes_per_step": 2 * (4 if dtype == torch.float32 else 8)},
            "gpu_launches": roof.pop("launches_per_step", None),
            "roofline": roof, "clocks": {**sampler.summary(), "n_samples": len(sampler.samples),
                                         "sampled_beyond_timed_region": extended},
            "roofline_stress": stress,
            "replicas_in_sync": in_sync,
            "gates": gates,
            "mode": "fused" if engine is not None else "eager (per-step fused safeguard+simulator kernel, torch MLP)",
        }
        if not args.no_cpu_baseline and world == 1:
            r = cpu_port_persistent(args.env, args.safeguard, H, 1, 1, budget_s=30.0, gate=args.gate)
            line["cpu_baseline"] = {"value": r["value"], "unit": "env-steps/s", "cores": r["workers"], "kind": "port",
                                    "sample": f"{r['steps']} step(s) of {r['workers']} processes x {r['n_w']} envs x H={H} (one SHAC episode "
                                              f"each, {r['total_s']:.1f} s wall, after one warm-up step), oracle port of the reference path, float64",
                                    "ms_per_env_step_per_core": r["ms_per_env_step_per_core"], "envs_per_worker": r["n_w"],
                                    "num_envs_cpu": r["workers"] * r["n_w"], "worker_errors": len(r["errors"]),
                                    "infeasible_env_steps": r["infeasible_env_steps"], "nonfinite_env_steps": r["nonfinite_env_steps"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def correctness_gates(args, algo, env, base, episode, dtype, dev):
    """Correctness beside the throughput: flag-word statistics of extra untimed episodes (infeasible
    fraction, non-finite actions, interventions, safety violations among feasible guarded steps) and a 64-environment slice of
    one safeguarded step, forward + backward, against the float64 oracle (max relative errors, verdict mismatches)."""
    import torch
    from safegbpo_b200 import _lib, ops
    out = {}
    engine = algo._fused_engine
    if engine is not None:
        tot = {"env_steps": 0, "guard_used": 0, "infeasible": 0, "nonfinite": 0, "intervened": 0, "safety_violations": 0}
        k = env.consts() if hasattr(env, "consts") else None
        for _ in range(3):
            episode(True)
            torch.cuda.synchronize()
            fl = engine.last["flags"]
            guard = (fl & _lib.FL_GUARD_USED) != 0
            infeas = (fl & _lib.FL_INFEASIBLE) != 0
            ok = guard & ~infeas
            viol = torch.zeros_like(ok)
            if k is not None and k.state_constrained and k.state_model not in (2, 3):      # (no device verdict for the lifted / unreduced models)
                viol |= ok & ((fl & _lib.FL_NEXT_IN_SET) == 0)
            if k is not None and k.action_constrained:
                viol |= ok & ((fl & _lib.FL_ACTION_IN_SET) == 0)
            tot["env_steps"] += fl.numel()
            tot["guard_used"] += int(guard.sum())
            tot["infeasible"] += int((guard & infeas).sum())
            tot["nonfinite"] += int(((fl & _lib.FL_NONFINITE) != 0).sum())
            tot["intervened"] += int(((fl & _lib.FL_INTERVENED) != 0).sum())
            tot["safety_violations"] += int(viol.sum())
        tot["infeasible_fraction_of_guarded"] = tot["infeasible"] / max(tot["guard_used"], 1)
        if args.safeguard == "ray_mask":
            # the reference's radial maps as shipped (ray_mask.py:108-113) return action_dist / safe_dist WITHOUT the
            # paper's factor safe_dist, so its executed actions are not confined to the safe set; the path reproduces the shipped
            # arithmetic (parity), and the exits are counted here, not gated
            tot["set_exits_reference_quirk_Q2"] = tot.pop("safety_violations")
        out["episodes"] = tot
    try:
        out["oracle_slice"] = oracle_slice_check(args, env, base, dtype, dev)
    except Exception as exc:             # the oracle leg is a checker: its absence is reported, never hidden
        out["oracle_slice"] = {"skipped": repr(exc)[:200]}
    ep = out.get("episodes", {})
    sl = out.get("oracle_slice", {})
    out["pass"] = bool(ep.get("nonfinite", 0) == 0 and ep.get("safety_violations", 0) == 0
                       and sl.get("verdict_mismatches", 0) == 0 and sl.get("max_rel_err", 0.0) <= sl.get("tolerance", 1.0))
    return out


def oracle_slice_check(args, env, base, dtype, dev, n=64):
    """One safeguarded step of `n` environments, forward + backward, on the GPU kernels (the benchmark's dtype and constants)
    against the float64 oracle port on the CPU; the gate is forced on for the slice so that every program is solved."""
    import copy
    import torch
    from safegbpo_b200 import _lib, ops
    from oracle.dynamics import OracleEnv
    from oracle.safeguard import OracleSafeguard
    if args.safeguard == "none" or args.env not in ("BalanceCartPole", "SwingUpCartPole", "BalancePendulum"):
        return {"skipped": "slice check covers the pendulum / cartpole safeguards (the other oracles take minutes per batch)"}
    f64 = torch.float64
    prev_dtype = torch.get_default_dtype()
    torch.set_default_dtype(f64)
    try:
        with torch.device("cpu"):          # the oracle is a CPU program; the bench's default device is the GPU
            gcpu = torch.Generator(device="cpu").manual_seed(7)
            oenv = OracleEnv(args.env, n, 240)
            oenv.reset(seed=0)
            spread = 0.1 if args.env == "BalancePendulum" else STATE_SPREAD
            st = oenv.state_center + (torch.rand(n, oenv.state_dim, dtype=f64, generator=gcpu, device="cpu") * 2 - 1) * oenv.state_half * spread
            a = torch.rand(n, oenv.action_dim, dtype=f64, generator=gcpu, device="cpu") * 2 - 1
            kind = {"boundary_projection": "boundary_projection", "ray_mask": "ray_mask"}[args.safeguard]
            oenv.state = torch.zeros_like(st)        # Q4: both sides bake in the linearisation at the zero (pre-reset) state
            sg = OracleSafeguard(oenv, kind, gate="always", strict=False)
            oenv.state = st.clone()
            dirs = sg.draw_directions() if kind == "ray_mask" and oenv.state_constrained else None
            w = oenv.noise_sample()
            s_o, a_o = st.clone().requires_grad_(True), a.clone().requires_grad_(True)
            oenv.state = s_o
            obs, rew, _, _ = sg.step(a_o, directions=dirs, noise=w)
            ok = ~sg.infeasible & torch.isfinite(obs).all(dim=1)
            (rew[ok].sum() + obs[ok].sum()).backward()
    finally:
        torch.set_default_dtype(prev_dtype)
    k = copy.copy(env.consts())
    k.gate_mode = _lib.GATE["always"]
    s_d = st.to(dev, dtype).requires_grad_(True)
    a_d = a.to(dev, dtype).requires_grad_(True)
    s_next, a_exec, obs_d, rew_d, flags = ops.safeguarded_step(s_d, a_d, w.to(dev, dtype), base.env_id, k,
                                                               dirs=None if dirs is None else dirs.to(dev, dtype))
    okd = ok.to(dev)
    (rew_d[okd].sum() + obs_d[okd].sum()).backward()
    torch.cuda.synchronize()

    def rel(x, y):
        x, y = x.detach().double().cpu()[ok], y.detach().double()[ok]
        return float((x - y).abs().max() / (y.abs().max() + 1e-30))

    infeasible_d = ((flags & _lib.FL_INFEASIBLE) != 0).cpu()
    errs = {"safe_action": rel(a_exec, sg.safe_action), "next_state": rel(s_next, oenv.state), "obs": rel(obs_d, obs),
            "grad_state": rel(s_d.grad, s_o.grad), "grad_action": rel(a_d.grad, a_o.grad)}
    tol = 1e-5 if dtype == torch.float32 else 1e-9
    if kind == "ray_mask" and dtype == torch.float32:
        tol = 2e-4          # the ray-mask maps divide by a chord length (tests/test_gpu_step_parity.py)
    return {"n_envs": n, "feasible": int(ok.sum()), "verdict_mismatches": int((infeasible_d != sg.infeasible).sum()),
            "rel_err": errs, "max_rel_err": max(errs.values()), "tolerance": tol, "gate": "always (every program solved)"}


def _peaks():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        return {}


def fused_roofline(args, engine, episode, dtype):
    """Roofline of the dominant kernel of the fused path, timed live with CUDA events on the launching
    stream (a few extra profiled episodes after the timed region, so the events do not perturb it)."""
    import torch
    engine.profile = True
    for _ in range(5):
        episode(False)
    torch.cuda.synchronize()
    engine.profile = False
    ms = engine.kernel_times()
    dominant = max(ms, key=lambda k: ms[k] or 0.0)
    N, H = args.num_envs, args.horizon
    scale = 1 if dtype == torch.float32 else 2
    peaks = _peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # algorithmic bytes: B_alg per env-step covers fwd+bwd of the rollout buffers; split per kernel
    b_alg = ALG_BYTES[args.env] * scale
    n_s, n_a = engine.base.state_dim, engine.base.action_dim
    jac_bytes = (2 * n_s + n_a) * 4 * scale * N * H          # the step-Jacobian kernel re-reads (s_t, a_t, w_t): part of the reverse share
    per_kernel_bytes = {"rollout_fwd": 0.45 * b_alg * N * H, "rollout_bwd": 0.55 * b_alg * N * H - (jac_bytes if ms.get("rollout_jac") else 0),
                        "rollout_jac": jac_bytes, "mlp_rows": (engine.base.obs_dim + 1) * 4 * scale * N * H}
    achieved = per_kernel_bytes[dominant] / (ms[dominant] * 1e-3) / 1e9
    total_alg = b_alg * N * H
    all_ms = sum(v for v in ms.values() if v)
    traffic = None          # DRAM bytes per launch of the dominant kernel from the committed ncu capture of THIS workload
    try:
        t = json.loads((ROOT / "profiles" / "r2_traffic.json").read_text())
        w = t["workload"]
        if (w["env"], w["safeguard"], w["num_envs_per_gpu"], w["horizon"], w["dtype"]) == \
                (args.env, args.safeguard, args.num_envs, args.horizon, args.dtype):
            traffic = t["bytes_per_launch"].get(dominant)
    except Exception:
        pass
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "kernel": dominant, "kernel_ms": ms, "alg_bytes_per_launch": per_kernel_bytes[dominant],
            "all_kernels_achieved_gbs": total_alg / (all_ms * 1e-3) / 1e9,
            "engines": {"forward": {0: "fma-tiles", 1: "tcgen05", 2: "warp-mma"}[engine.engine_fwd], "reverse": "parallel-tcgen05" if engine.engine_bwd else "sequential-fma"},
            "note": "rollout_bwd = the whole reverse pass after the step Jacobians (row-batched tcgen05 policy backward in Jacobian "
                    "mode, per-environment scan, row-batched tcgen05 policy backward in gradient mode).  This workload (4096 envs x "
                    "64 SEQUENTIAL steps with a 34 kFLOP policy inside every step) is latency/issue-bound in its forward kernel, not "
                    "HBM-bound: see profiles/r2_ncu_summary.md; traffic exceeds the algorithmic bytes on purpose (the forward kernel "
                    "stashes the two ELU outputs, 1 KB per env-step, so that the reverse pass is batched over all H x N rows); the "
                    "HBM-bound regime of the per-step kernels is reported under roofline_stress",
            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "launches_per_step": engine.launches_per_episode}


def step_kernel_roofline(args, base, env, dev, dtype, n_envs=None):
    """Roofline of the per-step safeguard+simulator kernels (step_fwd_kernel + step_bwd_kernel, policy bypassed:
    the stress variant of SURVEY §8d), each launched through the C ABI on preallocated buffers and timed alone
    with CUDA events on the launching stream (no torch autograd around them)."""
    import ctypes as C
    import torch
    from safegbpo_b200 import _lib
    N = n_envs or args.num_envs
    consts = env.consts() if hasattr(env, "consts") else base._plain
    n, m, o = base.state_dim, base.action_dim, base.obs_dim
    half = base.state_set.generator[0].diagonal()
    s = (base.state_set.center[0] + (torch.rand(N, n) * 2 - 1) * half * 0.9).contiguous()
    a = (torch.rand(N, m) * 2 - 1).contiguous()
    w = torch.zeros_like(s)
    aux = base._aux()
    aux = None if aux is None else aux[:1].expand(N, -1).contiguous()
    shared = base._shared()
    sh = C.byref(shared) if shared is not None else None
    dirs = None
    if consts.sg_kind == _lib.SG_RAY_MASK and consts.rm_zonotopic and consts.state_constrained:
        d = torch.rand(N, m, 2 * m) * 2 - 1
        dirs = (d / torch.linalg.vector_norm(d, dim=1, keepdim=True)).contiguous()
    s_next, a_exec, obs, rew = torch.empty(N, n), torch.empty(N, m), torch.empty(N, o), torch.empty(N)
    flags = torch.empty(N, dtype=torch.int32)
    g = [torch.ones(N, n), torch.ones(N, m), torch.ones(N, o), torch.ones(N)]
    gs, ga = torch.empty(N, n), torch.empty(N, m)
    lib, p, dc, st = _lib.lib(), _lib.ptr, _lib.dtype_code(dtype), _lib.stream_ptr()

    def fwd():
        _lib.check(lib.sgbpo_step_fwd(base.env_id, dc, N, C.byref(consts), sh, p(s), p(a), p(w), p(aux), p(dirs), None,
                                      p(s_next), p(a_exec), p(obs), p(rew), p(flags), st), "sgbpo_step_fwd")

    def bwd():
        _lib.check(lib.sgbpo_step_bwd(base.env_id, dc, N, C.byref(consts), sh, p(s), p(a), p(w), p(aux), p(dirs), None,
                                      p(g[0]), p(g[1]), p(g[2]), p(g[3]), p(gs), p(ga), st), "sgbpo_step_bwd")

    def timed(fn, reps=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_f, ms_b = timed(fwd), timed(bwd)
    ms = ms_f + ms_b
    peaks = _peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_alg = ALG_BYTES[args.env] * (1 if dtype == torch.float32 else 2) * N
    achieved = bytes_alg / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "kernel": "step_fwd_kernel + step_bwd_kernel (one launch pair per env step, C ABI, back-to-back launches)",
            "n_envs": N, "ms_fwd": ms_f, "ms_bwd": ms_b, "env_steps_per_s": N / (ms * 1e-3),
            "note": "instruction-bound (~1.5 k instructions per environment step for 84 + 93 bytes), not HBM-bound",
            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "launches_per_step": 2 * args.horizon}


def _apply_archs(a):
    global POLICY_ARCH, VF_ARCH
    if a.policy_arch:
        POLICY_ARCH = [int(x) for x in a.policy_arch.split(",")]
    if a.vf_arch:
        VF_ARCH = [int(x) for x in a.vf_arch.split(",")]


if __name__ == "__main__":
    if os.environ.get("SGBPO_BENCH_DEBUG"):          # hang diagnosis: dump every thread's python stack after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["SGBPO_BENCH_DEBUG"]), repeat=False, file=sys.stderr)
    a = parse()
    _apply_archs(a)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
