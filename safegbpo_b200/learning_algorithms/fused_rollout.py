"""Whole-horizon fused engine behind ``SHAC.collect_trajectories`` / ``SHAC.update_policy``.

Host side of csrc/sgbpo_rollout*.cu: resolves the episode schedule (reset-all of Q6, terminal rows,
household series), draws the random inputs (the reference's order, or one Philox launch), launches

    [sgbpo_episode_inputs] -> sgbpo_rollout_fwd -> sgbpo_mlp_rows (target critic over all rows)
                           -> sgbpo_episode_stats -> sgbpo_shac_loss                                  [collect_trajectories]
    sgbpo_rollout_jac -> sgbpo_mlp_rows_vjp (terminal rows) -> sgbpo_rollout_bwd                      [update_policy backward]

and maps the results back onto the reference's objects (CoupledBuffer tensors, ``policy.*.grad``).  Episodes whose schedule allows
it are replayed as CUDA graphs (two per episode, then one with the critic branch forked beside the reverse pass).  Everything
numeric runs in libsgbpo.so; with the sequential reverse engine (float64, other widths) two plain library calls remain: the
critic's input gradient where the network does not fit the kernel, and dW_hid = h1^T dz2 (cuBLAS through torch.matmul).
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from .. import _lib, rng
from .._lib import Mlp, Rollout, RolloutGrads, check, dtype_code, stream_ptr

FUSED_ENVS = {"BalancePendulum", "BalanceCartPole", "SwingUpCartPole", "BalanceQuadrotor", "ManageHousehold", "NavigateSeeker"}
# environments whose whole-horizon path exists only on the tensor-core engines (float32, policy [o -> 128 -> 128 -> m]): the
# household's per-step series and the seeker's obstacle columns / per-reset goal + obstacles are handled by rollout_tc_fwd_kernel
TC_ONLY_ENVS = {"ManageHousehold", "NavigateSeeker"}
WARP_MMA_ENVS = {"BalancePendulum", "BalanceCartPole", "SwingUpCartPole", "BalanceQuadrotor"}   # forward engine 2 (csrc/sgbpo_rollout_wm.cuh)
WIDTHS = (32, 64, 128)


def _mlp_layout(model: nn.Sequential):
    """(linears, norms) of a Linear/ELU/LayerNorm stack, or None if the module is not of that form."""
    mods = list(model)
    if len(mods) < 4 or (len(mods) - 1) % 3:
        return None
    n_hidden = (len(mods) - 1) // 3
    lin, norm = [], []
    for l in range(n_hidden):
        a, b, c = mods[3 * l: 3 * l + 3]
        if not (isinstance(a, nn.Linear) and isinstance(b, nn.ELU) and isinstance(c, nn.LayerNorm)):
            return None
        if b.alpha != 1.0 or abs(c.eps - 1e-5) > 0 or not c.elementwise_affine:
            return None
        lin.append(a)
        norm.append(c)
    if not isinstance(mods[-1], nn.Linear):
        return None
    lin.append(mods[-1])
    widths = {l.out_features for l in lin[:-1]}
    if len(widths) != 1 or next(iter(widths)) not in WIDTHS:
        return None
    return lin, norm


SMEM_LIMIT = 227 * 1024


def net_smem_floats(in_dim: int, width: int, n_hidden: int, out_dim: int) -> int:
    """NetSmem<T, W>::floats of csrc/sgbpo_mlp.cuh (weights as [k][W+1] + biases + LayerNorm parameters)."""
    ld = width + 1
    f = in_dim * ld + (n_hidden - 1) * width * ld + width * out_dim + 3 * n_hidden * width + out_dim
    return (f + 7) & ~7


def mlp_rows_fits(dtype, in_dim: int, width: int, n_hidden: int, out_dim: int = 1) -> bool:
    """At least one warp of 4 rows next to the parameters (csrc/sgbpo_rollout_impl.cuh::pick_shape)."""
    size = 4 if dtype == torch.float32 else 8
    return (net_smem_floats(in_dim, width, n_hidden, out_dim) + (in_dim + 2 * width) * 4 + 8) * size + 32 <= SMEM_LIMIT


def mlp_vjp_fits(dtype, in_dim: int, width: int, n_hidden: int, out_dim: int = 1) -> bool:
    """sgbpo_mlp_rows_vjp: one warp of 4 rows with its activation tile and n_hidden stash tiles next to the parameters."""
    size = 4 if dtype == torch.float32 else 8
    return (net_smem_floats(in_dim, width, n_hidden, out_dim) + (in_dim + (1 + n_hidden) * width) * 4 + 8) * size + 32 <= SMEM_LIMIT


def rollout_fits(dtype, obs_dim: int, width: int, n_hidden: int, act_dim: int) -> bool:
    """At least one warp of 2 rows of the backward kernel (4 tiles of W + the observation tile) fits."""
    size = 4 if dtype == torch.float32 else 8
    return (net_smem_floats(obs_dim, width, n_hidden, act_dim) + (obs_dim + 4 * width) * 2 + 8) * size + 32 <= SMEM_LIMIT


def _tall_skinny_gram(a: Tensor, b: Tensor) -> Tensor:
    """a^T b for (M, W) x (M, Wb) operands with M >> W: a plain library GEMM, split over M so that more than a
    handful of CTAs work on it (cuBLAS picks a 64x64 tile without split-K for this shape)."""
    M, W = a.shape
    Wb = b.shape[1]
    split = 1
    for cand in (256, 128, 64, 32, 16, 8):
        if M % cand == 0 and M // cand >= 256:
            split = cand
            break
    if split == 1:
        return a.t() @ b
    return torch.bmm(a.view(split, M // split, W).transpose(1, 2), b.view(split, M // split, Wb)).sum(dim=0)


class PackedMlp:
    """ctypes descriptor pointing at an MLP's torch parameters (the kernels read torch's layout directly)."""

    def __init__(self, model: nn.Sequential, log_std: Optional[Tensor] = None):
        self.lin, self.norm = _mlp_layout(model)
        self.log_std = log_std
        self.n_hidden = len(self.norm)
        self.width = self.lin[0].out_features
        self.in_dim, self.out_dim = self.lin[0].in_features, self.lin[-1].out_features
        self.refresh()

    def refresh(self):
        """Re-read the data pointers (parameters are updated in place by the optimisers, so this is only
        needed if a parameter tensor was replaced)."""
        d = Mlp()
        d.in_dim, d.width, d.n_hidden, d.out_dim = self.in_dim, self.width, self.n_hidden, self.out_dim
        for l, lin in enumerate(self.lin):
            if not (lin.weight.is_contiguous() and lin.bias.is_contiguous()):
                raise _lib.SgbpoError("MLP parameters must be contiguous")
            d.weight[l], d.bias[l] = lin.weight.data_ptr(), lin.bias.data_ptr()
        for l, norm in enumerate(self.norm):
            d.ln_weight[l], d.ln_bias[l] = norm.weight.data_ptr(), norm.bias.data_ptr()
        if self.log_std is not None:
            d.log_std = self.log_std.data_ptr()
        self.desc = d


class FusedRollout:
    def __init__(self, algo):
        self.algo = algo
        self.wrapper = algo.env if hasattr(algo.env, "consts") else None
        self.base = algo.env.env if self.wrapper is not None else algo.env
        self.rng_mode = "reference"        # "reference": per-step draws in the reference's order; "batched": 2 draws
        self.policy_pack = PackedMlp(algo.policy.model, algo.policy.log_std)
        self.vf_pack = PackedMlp(algo.target_value_function.model)
        vp = self.vf_pack
        # critic too large for shared memory in this dtype (e.g. float64 [128]x3): evaluate it with torch
        self.vf_native = mlp_rows_fits(self.base.state.dtype, vp.in_dim, vp.width, vp.n_hidden)
        self.vf_vjp_native = mlp_vjp_fits(self.base.state.dtype, vp.in_dim, vp.width, vp.n_hidden)
        self._grad_bufs = None
        self.last = None
        # Engines (include/sgbpo.h, sgbpo_rollout.engine): forward 0 = FMA register tiles, 1 = tcgen05 kernel (one CTA per 128
        # environments: wins once those tiles fill the GPU), 2 = warp-MMA kernel (8 environments per warp / pair of warps: wins
        # below that); reverse 0 = sequential FMA kernel, 1 = parallel tensor-core reverse pass.  SGBPO_ROLLOUT_ENGINE =
        # 0 | 1 | 2 | 3 forces (fwd, reverse) = (0, 0) | (1, 1) | (0, 1) | (2, 1) for A/B measurements and tests.
        self.engine_fwd, self.engine_bwd = self.pick_engine(self.base, self.policy_pack)
        self.engine = {0: 0, 1: 1, 2: 4}[self.engine_fwd] | (self.engine_bwd << 1)       # sgbpo_rollout.engine bits
        lib = _lib.lib()
        self.jac_stride = lib.sgbpo_rollout_jac_stride(self.base.env_id) if self.engine_bwd else 0
        self.pol_jac_stride = lib.sgbpo_rollout_pol_jac_stride(self.base.env_id) if self.engine_bwd else 0
        # [episode_inputs,] rollout_fwd, mlp_rows(_tc), episode_stats, shac_loss, mlp_rows_vjp, [rollout_jac,] rollout_bwd,
        # grad_norm, adam_apply (reverse engine 1: rollout_jac, policy_rows_bwd<jac>, rollout_scan, policy_rows_bwd<grad>
        # instead of rollout_bwd)
        self.use_graphs = True                 # replay the episode as two CUDA graphs when its schedule allows
        self.graph_warmup = 2                  # eager episodes before capture (optimizer state, cuBLAS handles)
        self._episodes = 0
        self._g = None                         # static buffers + captured graphs
        self.profile = False                   # record CUDA events around each of our kernels
        self.speculative_backward = True       # graph path: enqueue the backward graph right behind the forward graph
        self.merged_graph = os.environ.get("SGBPO_MERGED_GRAPH", "1") != "0"      # ... as ONE graph with the critic branch forked off
        self.forward_only = False              # model-free consumers (consumers.RolloutCollector): no activation stash, no reverse pass
        self.kernel_ms = {"rollout_fwd": [], "mlp_rows": [], "rollout_jac": [], "rollout_bwd": []}

    @property
    def launches_per_episode(self) -> int:
        return (10 if self.engine_bwd else 7) + (1 if self._native_inputs() else 0)     # (clip + Adam: one launch for small policies)

    TC_FORWARD_MIN_ENVS = 12288       # measured cross-over of the warp-MMA and tcgen05 forward kernels on one B200 (profiles/r2_engines.md)

    @staticmethod
    def pick_engine(base, pack: "PackedMlp") -> tuple[int, int]:
        force = os.environ.get("SGBPO_ROLLOUT_ENGINE", "")
        if base.state.dtype != torch.float32 or (force == "0" and base.ENV_NAME not in TC_ONLY_ENVS):
            return 0, 0
        if not _lib.lib().sgbpo_rollout_engine_supported(base.env_id, 0, C.byref(pack.desc)):
            return 0, 0
        if force == "1" or base.ENV_NAME in TC_ONLY_ENVS:
            return 1, 1
        if force == "2":
            return 0, 1
        wm_ok = base.ENV_NAME in WARP_MMA_ENVS and base.obs_dim < 16
        if force == "3":
            return (2 if wm_ok else 0), 1
        # forward: tcgen05 CTA tiles (128 environments per CTA) once they fill the GPU, below that the warp-MMA kernel
        # (8 environments per warp, no block-wide barrier in the time loop); FMA tiles where neither applies
        return (1 if base.num_envs >= FusedRollout.TC_FORWARD_MIN_ENVS else (2 if wm_ok else 0)), 1

    # ---- native input generation (rng_order = "batched"): one Philox launch instead of ~20 library launches ----------------
    def _native_inputs(self) -> bool:
        return (self.rng_mode == "batched" and rng._source is None and os.environ.get("SGBPO_NATIVE_RNG", "1") != "0"
                and self.base._reset_distribution() is not None)

    def _inputs_args(self, H: int, eps, noise, dirs, reset_states, counter) -> "_lib.EpisodeInputs":
        base = self.base
        a = _lib.EpisodeInputs()
        a.N, a.H, a.n, a.m = base.num_envs, H, base.state_dim, base.action_dim
        a.n_resets = 0 if reset_states is None else reset_states.shape[0]
        a.want_dirs = int(dirs is not None)
        a.seed = torch.initial_seed() & 0xFFFFFFFFFFFFFFFF
        a.episode = counter.data_ptr()
        a.eps, a.noise = eps.data_ptr(), noise.data_ptr()
        a.dirs = None if dirs is None else dirs.data_ptr()
        a.reset_states = None if reset_states is None else reset_states.data_ptr()
        ctr = base.noise_set.center[0].double().cpu().numpy()
        half = base.noise_set.generator[0].diagonal().double().cpu().numpy()
        c, G = base._reset_distribution()
        if G.shape[1] > 24:
            raise _lib.SgbpoError("reset distribution with more than 24 generators")
        a.n_reset_gens = G.shape[1]
        for i in range(base.state_dim):
            a.noise_lo[i], a.noise_span[i], a.reset_center[i] = float(ctr[i] - half[i]), float(2.0 * half[i]), float(c[i])
            for j in range(G.shape[1]):
                a.reset_gen[i][j] = float(G[i, j])
        return a

    def _stash_shape(self, H: int, N: int, W: int):
        """ELU-output stashes: [H][N][W] for the sequential reverse pass; for the parallel one 128-row tiles of 16-byte column
        chunks (csrc stash_index), padded to whole tiles."""
        return (-(-(H * N) // 128) * 128, W) if self.engine_bwd else (H, N, W)

    def _timed(self, name, fn):
        if not self.profile:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.kernel_ms[name].append((e0, e1))
        return out

    def kernel_times(self) -> dict:
        """Average duration (ms) of each kernel over the recorded launches (call after a synchronize)."""
        return {k: (sum(a.elapsed_time(b) for a, b in v) / len(v) if v else None) for k, v in self.kernel_ms.items()}

    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def try_create(algo) -> Optional["FusedRollout"]:
        env = algo.env
        base = env.env if hasattr(env, "consts") else env
        if getattr(base, "ENV_NAME", None) not in FUSED_ENVS:
            return None
        if not torch.cuda.is_available() or base.state.device.type != "cuda":
            return None
        if base.state.dtype not in (torch.float32, torch.float64):
            return None
        pol = algo.policy
        if not pol.stochastic or _mlp_layout(pol.model) is None or _mlp_layout(algo.target_value_function.model) is None:
            return None
        if len(_mlp_layout(pol.model)[1]) != 2:          # fused backward: two hidden layers
            return None
        if algo.buffer.store_safe_actions:
            return None
        pw = _mlp_layout(pol.model)[0][0].out_features
        if base.ENV_NAME in TC_ONLY_ENVS:
            if FusedRollout.pick_engine(base, PackedMlp(pol.model, pol.log_std)) != (1, 1):
                return None
        elif not rollout_fits(base.state.dtype, base.obs_dim, pw, 2, base.action_dim):
            return None
        if hasattr(env, "consts"):
            try:
                env.consts()
            except NotImplementedError:
                return None
        return FusedRollout(algo)

    def _consts(self):
        return self.wrapper.consts() if self.wrapper is not None else self.base._plain

    # ------------------------------------------------------------------------------------------------------
    def _schedule(self, H: int):
        """Episode schedule (pure host logic, no draws): which steps trigger the reset-all of Q6
        (simulator.py:100-106), which buffer rows are terminal, and the shared step counter per step."""
        base = self.base
        pending, steps = base._reset_pending, base.steps
        key = (bool(pending), int(steps), H, int(base.num_steps))
        cached = self.__dict__.get("_sched_cache")
        if cached is not None and cached[0] == key:          # (host time per episode matters: the lists are read-only downstream)
            return cached[1]
        reset_at, terminal_rows, pre_steps, post_steps = [], [False] * (H + 2), [], []
        for t in range(H):
            pre_steps.append(steps)               # household: T_out index used by execute_action (household.py:176)
            steps += 1
            if pending:
                reset_at.append(t)
                steps = 0
            post_steps.append(steps)              # household: series index of observation / reward
            pending = steps >= base.num_steps
            terminal_rows[t + 1] = pending
        terminal_rows[H] = True
        out = (reset_at, terminal_rows, pre_steps, post_steps, pending, steps)
        self.__dict__["_sched_cache"] = (key, out)
        return out

    def _max_resets(self, H: int) -> int:
        """Upper bound on reset-all events inside one horizon (a reset follows every num_steps steps)."""
        return H // max(1, int(self.base.num_steps)) + 1

    def _draw_inputs(self, H: int):
        """Random inputs in the reference's order or, with rng_mode="batched", in three
        batched draws (same distributions, different stream order)."""
        base, wrapper = self.base, self.wrapper
        N, n, m = base.num_envs, base.state_dim, base.action_dim
        dt, dev = base.state.dtype, base.state.device
        want_dirs = (wrapper is not None and wrapper.SG_KIND == _lib.SG_RAY_MASK
                     and getattr(wrapper, "zonotopic_approximation", False) and wrapper.state_constrained)
        reset_at, terminal_rows, pre_steps, post_steps, pending_end, steps_end = self._schedule(H)
        reset_idx, aux_src = [-1] * H, [-1] * H
        resets, reset_aux = [], []
        household = base._shared() is not None
        per_reset_aux = hasattr(base, "obstacles")          # NavigateSeeker: reset() re-samples goal and obstacles on the host
        if self._native_inputs():
            z = lambda *shape: torch.empty(*shape, dtype=dt, device=dev)
            eps_t, noise_t = z(H, N, m), z(H, N, n)
            dirs_t = z(H, N, m, 2 * m) if want_dirs else None
            rs = z(self._max_resets(H), N, n)
            counter = torch.tensor([self._episodes], dtype=torch.int64, device=dev)
            args = self._inputs_args(H, eps_t, noise_t, dirs_t, rs, counter)
            check(_lib.lib().sgbpo_episode_inputs(dtype_code(dt), C.byref(args), stream_ptr()), "sgbpo_episode_inputs")
            resets = [rs[r] for r in range(len(reset_at))]
            if resets:
                base._adopt_reset(resets[-1])
        elif self.rng_mode == "batched":
            eps_t = rng.randn(H, N, m).to(dt)
            u = rng.rand(H, N, n).to(dt)
            ctr = base.noise_set.center[0].to(dt)
            half = base.noise_set.generator[0].diagonal().to(dt)
            noise_t = torch.addcmul(ctr - half, u, 2.0 * half)      # = ctr + half (2u - 1), same form as the graph body
            dirs_t = None
            if want_dirs:
                d = rng.rand(H, N, m, 2 * m).to(dt) * 2 - 1
                dirs_t = d / torch.linalg.vector_norm(d, dim=2, keepdim=True)
            # the batched mode draws a FIXED number of reset states per episode (used or not), so the episode is the
            # same sequence of draws whether it runs eagerly or as a replayed CUDA graph
            if per_reset_aux:           # (never graph-replayed: the resets are drawn when they happen, with their goal / obstacles)
                for _ in reset_at:
                    base.reset()
                    resets.append(base.state.detach().to(dt))
                    reset_aux.append(base._aux().to(dt))
            else:
                snap = base._reset_snapshot()
                for r in range(self._max_resets(H)):
                    base.reset()
                    if r < len(reset_at):
                        resets.append(base.state.detach().to(dt))
                base._reset_restore(snap)
                if resets:
                    base._adopt_reset(resets[-1])
        else:
            eps, noise, dirs = [], [], []
            for t in range(H):
                eps.append(rng.randn(N, m).to(dt))
                if want_dirs:
                    dirs.append(wrapper._directions().to(dt))
                noise.append(base.noise_set.sample().to(dt))
                if t in reset_at:
                    base.reset()                  # draws in the reference's position of the stream
                    resets.append(base.state.detach().to(dt))
                    if per_reset_aux:
                        reset_aux.append(base._aux().to(dt))
            eps_t, noise_t = torch.stack(eps), torch.stack(noise)
            dirs_t = torch.stack(dirs) if dirs else None
        cur = -1
        for t in range(H):
            if t in reset_at:
                cur = reset_at.index(t)
                reset_idx[t] = cur
            aux_src[t] = cur
        shared_t = None
        if household:
            tout = torch.tensor([float(base._t_out[s]) for s in pre_steps], dtype=dt, device=dev)
            noise_t = noise_t.clone()
            noise_t[:, :, 1] = tout.unsqueeze(1)                      # household.py:175-176
            rows = []
            for s in post_steps:
                rows.append(np.concatenate([base._load[s:s + 5], base._pv[s:s + 5], base._t_out[s:s + 5], base._price[s:s + 5],
                                            [base._load[s], base._pv[s], base._price[s]]]))
            shared_t = torch.tensor(np.stack(rows), dtype=dt, device=dev)
        base.steps, base._reset_pending = steps_end, pending_end
        resets_t = torch.stack(resets) if resets else None
        self._reset_aux_t = torch.stack(reset_aux).contiguous() if reset_aux else None
        return eps_t, noise_t, dirs_t, resets_t, reset_idx, aux_src, terminal_rows, shared_t


    # ======================================================================================================
    # CUDA-graph path: an episode whose schedule has no reset inside the horizon is always the same sequence
    # of launches on the same buffers, so it is captured once (two graphs: collect / update) and replayed.
    # ======================================================================================================
    def _graphable(self, reset_at) -> bool:
        base = self.base
        resets_ok = not reset_at or (self.rng_mode == "batched" and len(reset_at) <= self._max_resets(self.algo.len_trajectories)
                                     and self._max_resets(self.algo.len_trajectories) <= 4)
        return (self.use_graphs and not self.profile and rng._source is None and resets_ok
                and base._shared() is None and not hasattr(base, "obstacles") and self._episodes >= self.graph_warmup
                and getattr(self.algo.policy_optim, "defaults", {}).get("capturable", False))

    def _alloc_static(self, H: int):
        base, algo = self.base, self.algo
        N, n, m, o, W = base.num_envs, base.state_dim, base.action_dim, base.obs_dim, self.policy_pack.width
        dt, dev = base.state.dtype, base.state.device
        z = lambda *shape, dtype=dt: torch.zeros(*shape, dtype=dtype, device=dev)
        want_dirs = (self.wrapper is not None and self.wrapper.SG_KIND == _lib.SG_RAY_MASK
                     and getattr(self.wrapper, "zonotopic_approximation", False) and self.wrapper.state_constrained)
        g = dict(H=H, eps=z(H, N, m), u=z(H, N, n), noise=z(H, N, n), dirs=z(H, N, m, 2 * m) if want_dirs else None,
                 s_start=z(N, n), states=z(H + 1, N, n), obs=z(H + 2, N, o), actions=z(H + 2, N, m), exec_actions=z(H, N, m),
                 rewards=z(H + 2, N), values=z(H + 2, N), flags=z(H, N, dtype=torch.int32),
                 term=z(H + 2, N, dtype=torch.bool), reward_sum=z(()), n_interv=z((), dtype=torch.int64),
                 bad=z((), dtype=torch.int32), aux0=None if base._aux() is None else z(N, base._aux().shape[1]),
                 h1=None if self.engine_bwd else z(H, N, W), e1=z(*self._stash_shape(H, N, W)), e2=z(*self._stash_shape(H, N, W)),
                 stats=z(H, N, 4),
                 dz2=None if self.engine_bwd else z(H, N, W), jac=z(H, N, self.jac_stride) if self.engine_bwd else None,
                 pol_jac=z(H, N, self.pol_jac_stride) if self.engine_bwd else None, dmu=z(H, N, m) if self.engine_bwd else None, loss=z(()),
                 fwd_graph=None, bwd_graph=None,
                 scal=z(2, dtype=torch.float64), scal_host=torch.zeros(2, dtype=torch.float64, device="cpu").pin_memory(),
                 loss_host=torch.zeros((), dtype=dt, device="cpu").pin_memory(),
                 done_host=torch.full((1,), -1, dtype=torch.int64, device="cpu").pin_memory(), branch=torch.cuda.Stream(), ep_graph=None,
                 ep_sync=False,
                 side=torch.cuda.Stream(), ev_fwd=torch.cuda.Event(), ev_side=torch.cuda.Event())
        # per-episode schedule (reset-all events of Q6, terminal rows, discounts): two small pinned host blocks
        # copied into static device tensors before every replay, so episodes WITH resets replay the same graphs
        RM = self._max_resets(H)
        KM = RM + 1
        g["RM"], g["KM"] = RM, KM
        g["sched_i_host"] = torch.zeros(2 * H + (H + 2) + (H + 2) + KM, dtype=torch.int32, device="cpu").pin_memory()
        g["sched_f_host"] = torch.zeros((H + 2) + H + 1 + KM, dtype=dt, device="cpu").pin_memory()
        g["sched_i"] = z(g["sched_i_host"].numel(), dtype=torch.int32)
        g["sched_f"] = z(g["sched_f_host"].numel())
        si, sf = g["sched_i"], g["sched_f"]
        g["reset_idx"], g["aux_src"] = si[0:H], si[H:2 * H]
        g["term_of_row"], g["term_int"], g["term_rows"] = si[2 * H:3 * H + 2], si[3 * H + 2:4 * H + 4], si[4 * H + 4:]
        g["disc_t"], g["grw"], g["norm_t"], g["term_w"] = sf[0:H + 2], sf[H + 2:2 * H + 2], sf[2 * H + 2:2 * H + 3], sf[2 * H + 3:]
        g["reset_states"] = z(RM, N, n)
        g["rng_ctr"] = z(1, dtype=torch.int64)
        g["rng_ctr_host"] = torch.zeros(1, dtype=torch.int64, device="cpu").pin_memory()
        g["inputs_args"] = self._inputs_args(H, g["eps"], g["noise"], g["dirs"], g["reset_states"], g["rng_ctr"]) \
            if self._native_inputs() else None
        g["gobs_term"] = z(KM, N, o)
        g["roww"] = z(KM, N)
        # (Linear2.weight.grad first: the reverse kernels flush it with 16-byte vector reductions, so it must be 16-byte aligned)
        sizes = ([W * W] if self.engine_bwd else []) + [o * W, W * m, 2 * W, 2 * W, 2 * W, m, m]
        g["flat"] = z(sum(sizes))
        parts = torch.split(g["flat"], sizes)
        g["dw_hid"] = parts[0].view(W, W) if self.engine_bwd else z(W, W)
        g["parts"] = parts[1:] if self.engine_bwd else parts
        g["ctr"] = base.noise_set.center[0].to(dt).clone()
        g["half"] = base.noise_set.generator[0].diagonal().to(dt).clone()
        g["noise_lo"], g["noise_span"] = g["ctr"] - g["half"], 2.0 * g["half"]
        r = Rollout()
        r.N, r.H = N, H
        p = lambda t: None if t is None else t.data_ptr()
        r.eps, r.noise, r.dirs, r.aux0 = p(g["eps"]), p(g["noise"]), p(g["dirs"]), p(g["aux0"])
        r.states, r.obs, r.actions, r.exec_actions = p(g["states"]), p(g["obs"]), p(g["actions"]), p(g["exec_actions"])
        r.rewards, r.flags = p(g["rewards"]), p(g["flags"])
        r.reset_states, r.reset_idx, r.aux_src = p(g["reset_states"]), p(g["reset_idx"]), p(g["aux_src"])
        r.stash_h1, r.stash_e1, r.stash_e2, r.stash_stats = p(g["h1"]), p(g["e1"]), p(g["e2"]), p(g["stats"])
        r.engine, r.jac, r.pol_jac, r.dmu = self.engine, p(g["jac"]), p(g["pol_jac"]), p(g["dmu"])
        g["r"] = r
        gr = RolloutGrads()
        gr.grw, gr.gobs_term, gr.term_of_row = p(g["grw"]), p(g["gobs_term"]), p(g["term_of_row"])
        gr.dz2_out = p(g["dz2"])
        (gr.g_w_in, gr.g_w_out, gr.g_b_hid, gr.g_gamma, gr.g_beta, gr.g_b_out, gr.g_log_std) = [x.data_ptr() for x in g["parts"][:7]]
        gr.g_w_hid = p(g["dw_hid"]) if self.engine_bwd else None
        g["gr"] = gr
        g_w_in, g_w_out, g_b_hid, g_gamma, g_beta, g_b_out, g_log_std = g["parts"][:7]
        pack = self.policy_pack
        g["grads"] = {
            pack.lin[0].weight: g_w_in.view(W, o), pack.lin[0].bias: g_b_hid.view(2, W)[0],
            pack.norm[0].weight: g_gamma.view(2, W)[0], pack.norm[0].bias: g_beta.view(2, W)[0],
            pack.lin[1].weight: g["dw_hid"], pack.lin[1].bias: g_b_hid.view(2, W)[1],
            pack.norm[1].weight: g_gamma.view(2, W)[1], pack.norm[1].bias: g_beta.view(2, W)[1],
            pack.lin[2].weight: g_w_out.view(m, W), pack.lin[2].bias: g_b_out,
            algo.policy.log_std: g_log_std,
        }
        self._g = g
        return g

    def _fwd_body(self, g):
        """Everything of collect_trajectories that touches the device, on static buffers (capturable)."""
        self._fwd_head(g)
        self._fwd_tail(g)

    def _episode_body(self, g):
        """One whole episode as ONE graph with a fork: after the forward kernel the critic values, episode scalars and loss value
        (branch stream, ending in the host copies of the scalars) run concurrently with the reverse pass (capture stream).  Both
        branches only need the forward kernel's buffers; the branch's tcgen05 critic shares the SMs with the small-footprint
        kernels of the reverse pass (critic input gradient, step Jacobians, scan), and the row-batched reverse kernels take over
        the SMs as its CTAs retire."""
        self._fwd_head(g)
        main, br = torch.cuda.current_stream(), g["branch"]
        br.wait_stream(main)
        with torch.cuda.stream(br):
            self._fwd_tail(g)
            g["scal_host"].copy_(g["scal"], non_blocking=True)
            g["loss_host"].copy_(g["loss"], non_blocking=True)
            g["done_host"].copy_(g["rng_ctr"], non_blocking=True)      # last: its arrival tells the host the scalars are there
        self._bwd_body(g)
        if g["ep_sync"]:       # data-parallel replicas: the gradient all-reduce is a node of the episode graph (no per-episode host call)
            import torch.distributed as dist
            for b in ([g["flat"]] if self.engine_bwd else [g["flat"], g["dw_hid"]]):
                dist.all_reduce(b, op=dist.ReduceOp.AVG)
        main.wait_stream(br)

    def _graph_sync(self) -> bool:
        """Whether the policy-gradient all-reduce of a data-parallel job is captured at the end of the episode graph (opt-in,
        SGBPO_GRAPH_ALLREDUCE=1): the hook installed by distributed.attach is the stock one and NCCL is the backend.  Measured on
        2 B200s: 1.080 -> 1.044 ms per episode (the per-episode host call of torch.distributed costs more than the 13 us
        collective: the host falls behind the device, profiles/r2_configs.md).  Off by default: with the collective captured, the
        process did not return from its exit path (process-group teardown with a live graph) within the test's time limit, and
        there was no GPU budget left to make that robust."""
        import torch.distributed as dist
        hook = getattr(self.algo, "grad_sync", None)
        return (os.environ.get("SGBPO_GRAPH_ALLREDUCE", "0") == "1" and hook is not None and getattr(hook, "graph_capturable", False)
                and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and dist.get_backend() == "nccl")

    def _fwd_head(self, g):
        base, algo, H = self.base, self.algo, g["H"]
        N, o = base.num_envs, base.obs_dim
        dt = base.state.dtype
        lib = _lib.lib()
        g["states"][0].copy_(g["s_start"])      # (separate input buffer: the body must be idempotent for capture)
        if g["inputs_args"] is not None:        # eps, noise, directions and reset states in ONE launch of ours
            check(lib.sgbpo_episode_inputs(dtype_code(dt), C.byref(g["inputs_args"]), stream_ptr()), "sgbpo_episode_inputs")
        elif self.rng_mode == "batched":
            g["eps"].normal_()
            g["u"].uniform_()
            if g["dirs"] is not None:
                g["dirs"].uniform_(-1.0, 1.0)
        else:                                   # the reference's per-step order: policy draw, directions, noise
            for t in range(H):
                g["eps"][t].normal_()
                if g["dirs"] is not None:
                    g["dirs"][t].uniform_(-1.0, 1.0)
                g["u"][t].uniform_()
        if g["inputs_args"] is None:
            torch.addcmul(g["noise_lo"], g["u"], g["noise_span"], out=g["noise"])
            if g["dirs"] is not None:
                g["dirs"].div_(torch.linalg.vector_norm(g["dirs"], dim=2, keepdim=True))
        if self.rng_mode == "batched" and g["inputs_args"] is None:          # fixed number of reset-state draws per episode (see _draw_inputs)
            snap = base._reset_snapshot()
            for r in range(g["RM"]):
                base.reset()
                g["reset_states"][r].copy_(base.state)
            base._reset_restore(snap)
        check(lib.sgbpo_rollout_fwd(base.env_id, dtype_code(dt), C.byref(self._consts()), C.byref(self.policy_pack.desc),
                                    C.byref(g["r"]), stream_ptr()), "sgbpo_rollout_fwd")

    def _fwd_tail(self, g):
        base, algo, H = self.base, self.algo, g["H"]
        N, o = base.num_envs, base.obs_dim
        dt = base.state.dtype
        lib = _lib.lib()
        if self.vf_native:
            check(lib.sgbpo_mlp_rows(dtype_code(dt), H * N, C.byref(self.vf_pack.desc), C.c_void_p(g["obs"][1].data_ptr()),
                                     C.c_void_p(g["values"][1].data_ptr()), stream_ptr()), "sgbpo_mlp_rows")
        else:
            with torch.no_grad():
                g["values"][1:H + 1] = algo.target_value_function(g["obs"][1:H + 1].reshape(H * N, o)).view(H, N)
        # reward sum, intervention count, any non-finite / infeasible step: one launch (was ~10 library launches)
        guarded = self.wrapper is not None
        check(lib.sgbpo_episode_stats(dtype_code(dt), g["rewards"].numel(), C.c_void_p(g["rewards"].data_ptr()),
                                      g["flags"].numel() if guarded else 0, C.c_void_p(g["flags"].data_ptr()),
                                      _lib.FL_INTERVENED, _lib.FL_NONFINITE | _lib.FL_INFEASIBLE,
                                      C.c_void_p(g["scal"].data_ptr()), C.c_void_p(g["n_interv"].data_ptr()),
                                      C.c_void_p(g["bad"].data_ptr()), stream_ptr()), "sgbpo_episode_stats")
        # the episode's loss VALUE is complete here (shac.py:131-142 reads rewards, values and the schedule only); computing it
        # at the end of the forward graph lets it leave with the reward scalar, so update_policy's float(loss) does not wait
        # for the reverse pass and the host prepares the next episode while the reverse pass runs
        if not self.forward_only:
            check(lib.sgbpo_shac_loss(dtype_code(dt), H + 2, N, C.c_void_p(g["disc_t"].data_ptr()),
                                      C.c_void_p(g["term_int"].data_ptr()), C.c_void_p(g["values"].data_ptr()),
                                      C.c_void_p(g["rewards"].data_ptr()), C.c_void_p(g["norm_t"].data_ptr()),
                                      C.c_void_p(g["loss"].data_ptr()), stream_ptr()), "sgbpo_shac_loss")

    def _bwd_body(self, g):
        """Loss, terminal-value gradient, fused reverse pass, weight-gradient GEMM (capturable)."""
        base, algo, H = self.base, self.algo, g["H"]
        N = base.num_envs
        dt = base.state.dtype
        W = self.policy_pack.width
        o = base.obs_dim
        g["flat"].zero_()
        if self.engine_bwd:            # step Jacobians of all H x N steps in one parallel launch (first: no shared memory, so in the
            # merged episode graph it shares the SMs with the critic kernel of the other branch)
            check(_lib.lib().sgbpo_rollout_jac(base.env_id, dtype_code(dt), C.byref(self._consts()), C.byref(g["r"]), stream_ptr()),
                  "sgbpo_rollout_jac")
        # critic input-gradient on the (at most KM) terminal rows; unused slots carry weight 0
        x = g["obs"].index_select(0, g["term_rows"]).detach()
        if self.vf_vjp_native:         # one launch: forward + transposed pass of the target critic (csrc mlp_rows_vjp_kernel)
            check(_lib.lib().sgbpo_mlp_rows_vjp(dtype_code(dt), g["KM"] * N, C.byref(self.vf_pack.desc), C.c_void_p(x.data_ptr()),
                                                C.c_void_p(g["roww"].data_ptr()), C.c_void_p(g["gobs_term"].data_ptr()), stream_ptr()),
                  "sgbpo_mlp_rows_vjp")
        else:
            x = x.clone().requires_grad_(True)
            v = (algo.target_value_function(x.view(-1, o)).view(g["KM"], N) * g["term_w"].unsqueeze(1)).sum()
            g["gobs_term"].copy_(torch.autograd.grad(v, x)[0])
        check(_lib.lib().sgbpo_rollout_bwd(base.env_id, dtype_code(dt), C.byref(self._consts()),
                                           C.byref(self.policy_pack.desc), C.byref(g["r"]), C.byref(g["gr"]), stream_ptr()),
              "sgbpo_rollout_bwd")
        if not self.engine_bwd:
            g["dw_hid"].copy_(_tall_skinny_gram(g["dz2"].view(H * N, W), g["h1"].view(H * N, W)))

    def _capture(self, body, g):
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up on a side stream, as torch.cuda.graph requires
            body(g)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body(g)
        return graph

    def _fill_schedule(self, g, H: int, reset_at, terminal_rows):
        """Host side of the per-episode schedule: reset / aux indices, terminal rows, discounts (shac.py:131-139)."""
        N = self.base.num_envs
        disc, norm = self._weights(terminal_rows, H, N)
        reset_idx, aux_src, cur = [-1] * H, [-1] * H, -1
        for t in range(H):
            if t in reset_at:
                cur = reset_at.index(t)
                reset_idx[t] = cur
            aux_src[t] = cur
        rows = [j for j in range(1, H + 1) if terminal_rows[j]]
        term_of_row = [-1] * (H + 2)
        for kk, j in enumerate(rows):
            term_of_row[j] = kk
        KM = g["KM"]
        term_rows = rows + [H] * (KM - len(rows))
        term_w = [-disc[j] / norm for j in rows] + [0.0] * (KM - len(rows))
        grw = [0.0 if terminal_rows[t] else -disc[t] / norm for t in range(H)]
        g["sched_i_host"].copy_(torch.tensor(reset_idx + aux_src + term_of_row + [int(f) for f in terminal_rows] + term_rows,
                                             dtype=torch.int32, device="cpu"))
        g["sched_f_host"].copy_(torch.tensor(disc + grw + [float(norm)] + term_w, dtype=g["sched_f_host"].dtype, device="cpu"))
        g["sched_i"].copy_(g["sched_i_host"], non_blocking=True)
        g["sched_f"].copy_(g["sched_f_host"], non_blocking=True)
        # tensors derived from the schedule (only when it changes: most episodes repeat the previous one)
        g["term"].copy_((g["term_int"] != 0).unsqueeze(1).expand(-1, N))
        g["roww"].copy_(g["term_w"].unsqueeze(1).expand(g["KM"], N))

    def _collect_graph(self, H: int, steps_end: int, pending_end: bool, reset_at, terminal_rows) -> float:
        base, algo = self.base, self.algo
        g = self._g if self._g is not None and self._g["H"] == H else self._alloc_static(H)
        key = (tuple(reset_at), tuple(terminal_rows))
        if g.get("sched_key") != key:            # most episodes repeat the previous schedule: nothing to upload
            self._fill_schedule(g, H, reset_at, terminal_rows)
            g["sched_key"] = key
        buf = algo.buffer
        base.cut_computation_graph()
        g["s_start"].copy_(base.state)
        if not g.get("row0_ready"):
            g["obs"][0].copy_(buf.observations.tensor[0])
            g["values"][0].copy_(buf.values.tensor[0])
        g["row0_ready"] = False
        if g["aux0"] is not None:
            g["aux0"].copy_(base._aux())
        g["rng_ctr_host"][0] = self._episodes
        g["rng_ctr"].copy_(g["rng_ctr_host"], non_blocking=True)
        merged = (self.merged_graph and self.speculative_backward and not self.forward_only and g["bwd_graph"] is not None
                  and g["inputs_args"] is not None)
        if merged:
            want_sync = self._graph_sync()
            if g["ep_graph"] is not None and g.get("ep_sync") != want_sync:
                g["ep_graph"] = None                     # the job became / stopped being data-parallel: capture again
            g["ep_sync"] = want_sync
        if merged and g["ep_graph"] is None:
            g["ep_graph"] = self._capture(self._episode_body, g)
        if merged:
            return self._replay_episode_graph(g, H, steps_end, pending_end, reset_at, terminal_rows)
        if g["fwd_graph"] is None:
            rng_state = torch.cuda.get_rng_state()
            g["fwd_graph"] = self._capture(self._fwd_body, g)
            torch.cuda.set_rng_state(rng_state)          # the warm-up + capture passes must not consume the stream
        g["fwd_graph"].replay()
        # speculative reverse pass: update_policy always follows collect_trajectories in SHAC._learn_episode and the
        # policy does not change in between, so the backward graph is enqueued right behind the forward one and
        # runs while the host waits for the reward scalar; it only writes the static gradient buffers, which
        # nothing reads until update_policy attaches them
        g["ev_fwd"].record()
        g["bwd_enqueued"] = False
        if g["bwd_graph"] is not None and self.speculative_backward:
            g["bwd_graph"].replay()
            g["bwd_enqueued"] = True
        # the episode's host-visible scalars (reward sum, bad-status) leave on a side stream that waits for the
        # forward graph only, so this read-back does not serialise behind the backward graph
        with torch.cuda.stream(g["side"]):
            g["side"].wait_event(g["ev_fwd"])
            g["scal_host"].copy_(g["scal"], non_blocking=True)
            g["loss_host"].copy_(g["loss"], non_blocking=True)
            g["ev_side"].record(g["side"])
        g["ev_side"].synchronize()
        return self._finish_collect(g, H, steps_end, pending_end, reset_at, terminal_rows)

    def _replay_episode_graph(self, g, H, steps_end, pending_end, reset_at, terminal_rows) -> float:
        g["ep_graph"].replay()
        g["bwd_enqueued"] = True
        g["grads_synced"] = bool(g["ep_sync"])
        # the scalars are copied to pinned memory by the graph's branch; the host waits for the LAST of those copies (the episode
        # counter) to land instead of for the whole graph, so it still runs ahead of the reverse pass
        t0, want = time.perf_counter(), self._episodes
        while int(g["done_host"][0]) != want:
            if time.perf_counter() - t0 > 120.0:
                raise _lib.SgbpoError("episode graph: the scalar read-back did not arrive (device fault?)")
        return self._finish_collect(g, H, steps_end, pending_end, reset_at, terminal_rows)

    def _finish_collect(self, g, H, steps_end, pending_end, reset_at, terminal_rows) -> float:
        base, buf = self.base, self.algo.buffer
        buf.observations.tensor, buf.values.tensor, buf.rewards.tensor = g["obs"], g["values"], g["rewards"]
        buf.actions.tensor, buf.terminals.tensor = g["actions"], g["term"]
        if g.get("t_full") is None:
            g["t_full"], g["t_zero"] = torch.full_like(buf.t, H), torch.zeros_like(buf.t)
        buf.t = g["t_full"]
        buf.fast_reset = self._fast_buffer_reset
        base.state = g["states"][H]
        base.steps, base._reset_pending = steps_end, pending_end
        if reset_at:
            base._adopt_reset(g["reset_states"][len(reset_at) - 1].clone())
        if g.get("sr") is None:
            g["sr"] = (torch.zeros_like(base.scheduled_reset), torch.ones_like(base.scheduled_reset))
        base.scheduled_reset = g["sr"][1 if pending_end else 0]
        base.last_flags, base.last_exec_action = g["flags"][H - 1], g["exec_actions"][H - 1]
        if self.wrapper is not None:
            w = self.wrapper
            w.safe_action = g["exec_actions"][H - 1]
            w._interventions = w._interventions + g["n_interv"]
        self.last = dict(graph=True, terminal_rows=terminal_rows, flags=g["flags"], states=g["states"], exec_actions=g["exec_actions"],
                         grads_synced=bool(g.pop("grads_synced", False)))
        if self.wrapper is not None:
            if w.strict:
                if g["scal_host"][1] != 0:
                    raise ValueError("a safeguarded step produced a non-finite action or an infeasible program "
                                     "(the reference raises here: safeguard.py:68-75 / solver error)")
            else:
                w._status = w._status | g["bad"]
        return float(g["scal_host"][0]) / base.num_envs / H

    def static_grad_buffers(self):
        """The two static tensors that hold every policy gradient of the last graph-replayed backward (all
        `policy.*.grad` are views of them), or None when the last episode ran outside the graphs."""
        g = self._g
        if g is None or self.last is None or not self.last.get("graph"):
            return None
        for prm, buf in g["grads"].items():
            if prm.grad is None or prm.grad.data_ptr() != buf.data_ptr():
                return None
        return [g["flat"]] if self.engine_bwd else [g["flat"], g["dw_hid"]]

    def _fast_buffer_reset(self, reset_observation=None, reset_value=None) -> bool:
        """CoupledBuffer.reset() while the buffer's tensors ARE the static graph buffers: carry the last
        observation / value to row 0 in place (coupled_buffer.py:113-131 does the same on fresh storage; the
        other rows are rewritten by the next collect_trajectories).  Returns False when the buffer was
        re-bound by someone else, so the generic reset runs."""
        g, buf = self._g, self.algo.buffer
        if g is None or buf.observations.tensor is not g["obs"] or buf.t is not g["t_full"]:
            buf.fast_reset = None
            return False
        H = g["H"]
        if reset_observation is None:
            g["obs"][0].copy_(g["obs"][H])
            g["values"][0].copy_(g["values"][H])
        else:
            if buf.store_values and reset_value is None:
                return False
            g["obs"][0].copy_(reset_observation.detach())
            g["values"][0].copy_(reset_value.detach())
        buf.t = g["t_zero"]
        g["row0_ready"] = True
        return True

    def _backward_graph(self) -> Tensor:
        g, algo = self._g, self.algo
        for prm, buf in g["grads"].items():
            if prm.grad is None:
                prm.grad = buf
            elif prm.grad.data_ptr() != buf.data_ptr():
                raise _lib.SgbpoError("fused graph path: policy gradients must be zeroed (set_to_none) before the episode")
        if g["bwd_graph"] is None:
            g["bwd_graph"] = self._capture(self._bwd_body, g)
        if not g.get("bwd_enqueued"):
            g["bwd_graph"].replay()
        g["bwd_enqueued"] = False
        return g["loss_host"].clone()        # (host copy made behind the forward graph: reading it does not wait for the reverse pass)

    # ------------------------------------------------------------------------------------------------------
    def collect_trajectories(self) -> float:
        algo, base = self.algo, self.base
        H, N = algo.len_trajectories, base.num_envs
        n, m, o = base.state_dim, base.action_dim, base.obs_dim
        dt, dev = base.state.dtype, base.state.device
        self._episodes += 1
        reset_at, terminal_rows, _, _, pending_end, steps_end = self._schedule(H)
        if self._graphable(reset_at):
            return self._collect_graph(H, steps_end, pending_end, reset_at, terminal_rows)
        if self._g is not None:
            self._g["row0_ready"] = False          # this episode runs outside the static buffers
        base.cut_computation_graph()
        s0 = base.state.detach().contiguous()
        aux0 = base._aux()
        aux0 = None if aux0 is None else aux0.detach().to(dt).contiguous()
        buf = algo.buffer
        obs0 = buf.observations.tensor[0].detach()
        val0 = buf.values.tensor[0].detach()
        eps, noise, dirs, resets, reset_idx, aux_src, terminal_rows, shared = self._draw_inputs(H)
        opts = dict(dtype=dt, device=dev)
        states = torch.empty(H + 1, N, n, **opts)
        obs = torch.zeros(H + 2, N, o, **opts)
        actions = torch.zeros(H + 2, N, m, **opts)
        exec_actions = torch.empty(H, N, m, **opts)
        rewards = torch.zeros(H + 2, N, **opts)
        values = torch.zeros(H + 2, N, **opts)
        flags = torch.empty(H, N, dtype=torch.int32, device=dev)
        states[0], obs[0], values[0] = s0, obs0, val0
        sched = torch.tensor([reset_idx, aux_src], dtype=torch.int32, device=dev)      # one small H2D copy
        reset_idx_t, aux_src_t = sched[0], sched[1]
        r = Rollout()
        r.N, r.H = N, H
        keep = [eps, noise, dirs, aux0, resets, reset_idx_t, aux_src_t, shared, self._reset_aux_t]
        r.reset_aux = None if self._reset_aux_t is None else self._reset_aux_t.data_ptr()
        p = lambda t: None if t is None else t.contiguous().data_ptr()
        eps, noise = eps.contiguous(), noise.contiguous()
        r.eps, r.noise, r.dirs, r.aux0 = p(eps), p(noise), p(dirs), p(aux0)
        r.reset_states, r.reset_idx, r.aux_src, r.shared = p(resets), p(reset_idx_t), p(aux_src_t), p(shared)
        r.states, r.obs, r.actions, r.exec_actions = states.data_ptr(), obs.data_ptr(), actions.data_ptr(), exec_actions.data_ptr()
        r.rewards, r.flags = rewards.data_ptr(), flags.data_ptr()
        W = self.policy_pack.width
        if not self.forward_only and (self._grad_bufs is None or self._grad_bufs["key"] != (H, N, W, dt)):
            mk = lambda *shape: torch.empty(*shape, dtype=dt, device=dev)
            self._grad_bufs = dict(key=(H, N, W, dt), e1=mk(*self._stash_shape(H, N, W)), e2=mk(*self._stash_shape(H, N, W)),
                                   stats=mk(H, N, 4))
            if self.engine_bwd:
                self._grad_bufs.update(jac=mk(H, N, self.jac_stride), pol_jac=mk(H, N, self.pol_jac_stride), dmu=mk(H, N, m))
            else:
                self._grad_bufs.update(h1=mk(H, N, W), dz2=mk(H, N, W))
        gb = self._grad_bufs
        if not self.forward_only:
            r.stash_e1, r.stash_e2, r.stash_stats = gb["e1"].data_ptr(), gb["e2"].data_ptr(), gb["stats"].data_ptr()
        r.engine = self.engine
        if self.forward_only:
            pass
        elif self.engine_bwd:
            r.jac, r.pol_jac, r.dmu = gb["jac"].data_ptr(), gb["pol_jac"].data_ptr(), gb["dmu"].data_ptr()
        elif not self.forward_only:
            r.stash_h1 = gb["h1"].data_ptr()
        lib = _lib.lib()
        self._timed("rollout_fwd", lambda: check(
            lib.sgbpo_rollout_fwd(base.env_id, dtype_code(dt), C.byref(self._consts()), C.byref(self.policy_pack.desc),
                                  C.byref(r), stream_ptr()), "sgbpo_rollout_fwd"))
        if self.vf_native:
            self._timed("mlp_rows", lambda: check(
                lib.sgbpo_mlp_rows(dtype_code(dt), H * N, C.byref(self.vf_pack.desc), C.c_void_p(obs[1].data_ptr()),
                                   C.c_void_p(values[1].data_ptr()), stream_ptr()), "sgbpo_mlp_rows"))
        else:
            with torch.no_grad():
                values[1:H + 1] = algo.target_value_function(obs[1:H + 1].reshape(H * N, o)).view(H, N)
        # hand the results to the reference's objects
        term = torch.zeros(H + 2, N, dtype=torch.bool, device=dev)
        rows = [j for j, f in enumerate(terminal_rows) if f]
        term[rows] = True
        buf.observations.tensor, buf.values.tensor, buf.rewards.tensor = obs, values, rewards
        buf.actions.tensor, buf.terminals.tensor = actions, term
        buf.t = torch.full_like(buf.t, H)
        base.state = states[H]
        base.scheduled_reset = torch.full_like(base.scheduled_reset, bool(base._reset_pending))
        base.last_flags, base.last_exec_action = flags[H - 1], exec_actions[H - 1]
        self.last = dict(r=r, keep=keep + [eps, noise, states, obs, actions, exec_actions, rewards, flags, values],
                         terminal_rows=terminal_rows, obs=obs, values=values, rewards=rewards, flags=flags,
                         exec_actions=exec_actions, states=states)
        if self.wrapper is not None:
            w = self.wrapper
            w.safe_action = exec_actions[H - 1]
            w._interventions = w._interventions + ((flags & _lib.FL_INTERVENED) != 0).sum()
            bad = ((flags & (_lib.FL_NONFINITE | _lib.FL_INFEASIBLE)) != 0).any()
            if w.strict:
                if bool(bad):
                    raise ValueError("a safeguarded step produced a non-finite action or an infeasible program "
                                     "(the reference raises here: safeguard.py:68-75 / solver error)")
            else:
                w._status = w._status | bad.to(torch.int32)
        return float(rewards.sum()) / N / H

    # ------------------------------------------------------------------------------------------------------
    def _weights(self, terminal_rows, H: int, N: int):
        """Per-row discount of shac.py:131-139 (uniform over environments) and the normaliser."""
        gamma = self.algo.GAMMA
        exponent, drop = [], 0
        for r in range(H + 2):
            exponent.append(r - drop)
            if terminal_rows[r]:
                drop += r + 1                      # cumulative, as shipped (see SHAC.discount_weights)
        disc = [gamma ** e for e in exponent]
        norm = N * H + N * sum(terminal_rows)
        return disc, norm

    def backward(self) -> Tensor:
        algo, base, last = self.algo, self.base, self.last
        if last.get("graph"):
            return self._backward_graph()
        H, N = algo.len_trajectories, base.num_envs
        o, m, W = base.obs_dim, base.action_dim, self.policy_pack.width
        dt, dev = base.state.dtype, base.state.device
        terminal_rows = last["terminal_rows"]
        disc, norm = self._weights(terminal_rows, H, N)
        disc_t = torch.tensor(disc, dtype=dt, device=dev)
        term_mask = torch.tensor(terminal_rows, dtype=torch.bool, device=dev)
        picked = torch.where(term_mask.unsqueeze(1), last["values"], last["rewards"])
        loss = -(disc_t.unsqueeze(1) * picked).sum() / norm
        grw = torch.tensor([0.0 if terminal_rows[t] else -disc[t] / norm for t in range(H)], dtype=dt, device=dev)
        # critic input-gradient on the terminal rows (N x o inputs each): plain autograd through the target critic
        rows = [j for j in range(1, H + 1) if terminal_rows[j]]
        term_of_row = [-1] * (H + 2)
        for kk, j in enumerate(rows):
            term_of_row[j] = kk
        if self.vf_vjp_native:
            x = last["obs"][rows].detach().contiguous()                                    # [K, N, o]
            roww = torch.tensor([-disc[j] / norm for j in rows], dtype=dt, device=dev).unsqueeze(1).expand(len(rows), N).contiguous()
            gobs_term = torch.empty(len(rows), N, o, dtype=dt, device=dev)
            check(_lib.lib().sgbpo_mlp_rows_vjp(dtype_code(dt), len(rows) * N, C.byref(self.vf_pack.desc), C.c_void_p(x.data_ptr()),
                                                C.c_void_p(roww.data_ptr()), C.c_void_p(gobs_term.data_ptr()), stream_ptr()),
                  "sgbpo_mlp_rows_vjp")
        else:
            gterm = []
            for kk, j in enumerate(rows):
                x = last["obs"][j].detach().clone().requires_grad_(True)
                v = algo.target_value_function(x).sum() * (-disc[j] / norm)
                gterm.append(torch.autograd.grad(v, x)[0])
            gobs_term = torch.stack(gterm).contiguous()
        term_of_row_t = torch.tensor(term_of_row, dtype=torch.int32, device=dev)
        tc = bool(self.engine_bwd)
        sizes = ([W * W] if tc else []) + [o * W, W * m, 2 * W, 2 * W, 2 * W, m, m]
        flat = torch.zeros(sum(sizes), dtype=dt, device=dev)
        parts = torch.split(flat, sizes)
        dw_part, parts = (parts[0], parts[1:]) if tc else (None, parts)
        g = RolloutGrads()
        g.grw, g.gobs_term, g.term_of_row = grw.data_ptr(), gobs_term.data_ptr(), term_of_row_t.data_ptr()
        (g.g_w_in, g.g_w_out, g.g_b_hid, g.g_gamma, g.g_beta, g.g_b_out, g.g_log_std) = [x.data_ptr() for x in parts[:7]]
        if tc:
            g.g_w_hid = dw_part.data_ptr()
            self._timed("rollout_jac", lambda: check(
                _lib.lib().sgbpo_rollout_jac(base.env_id, dtype_code(dt), C.byref(self._consts()), C.byref(last["r"]), stream_ptr()),
                "sgbpo_rollout_jac"))
        else:
            g.dz2_out = self._grad_bufs["dz2"].data_ptr()
        self._timed("rollout_bwd", lambda: check(
            _lib.lib().sgbpo_rollout_bwd(base.env_id, dtype_code(dt), C.byref(self._consts()),
                                         C.byref(self.policy_pack.desc), C.byref(last["r"]), C.byref(g), stream_ptr()),
            "sgbpo_rollout_bwd"))
        if tc:
            dw_hid = dw_part.view(W, W)
        else:
            dw_hid = _tall_skinny_gram(self._grad_bufs["dz2"].view(H * N, W), self._grad_bufs["h1"].view(H * N, W))   # Linear.weight.grad [out][in]
        g_w_in, g_w_out, g_b_hid, g_gamma, g_beta, g_b_out, g_log_std = parts[:7]
        pack = self.policy_pack
        grads = {
            pack.lin[0].weight: g_w_in.view(W, o), pack.lin[0].bias: g_b_hid.view(2, W)[0],
            pack.norm[0].weight: g_gamma.view(2, W)[0], pack.norm[0].bias: g_beta.view(2, W)[0],
            pack.lin[1].weight: dw_hid, pack.lin[1].bias: g_b_hid.view(2, W)[1],
            pack.norm[1].weight: g_gamma.view(2, W)[1], pack.norm[1].bias: g_beta.view(2, W)[1],
            pack.lin[2].weight: g_w_out.view(m, W), pack.lin[2].bias: g_b_out,
            algo.policy.log_std: g_log_std,
        }
        for prm, val in grads.items():
            prm.grad = val if prm.grad is None else prm.grad + val
        return loss.detach()

