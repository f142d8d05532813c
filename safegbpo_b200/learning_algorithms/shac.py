"""Short Horizon Actor Critic (reference: src/learning_algorithms/shac.py:14-251).

``collect_trajectories`` + the backward half of ``update_policy`` are the hot path.  Two execution modes
with identical results:
  * eager  -- the reference's python loop over the horizon; each step is the policy MLP (torch) plus ONE
              fused CUDA launch for safeguard + simulator (ops.SafeguardedStepFn);
  * fused  -- the whole horizon in csrc/sgbpo_rollout.cu (see ``fused_rollout.py``), used automatically
              when env, safeguard and networks are of the supported fused types.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

from .components.coupled_buffer import CoupledBuffer
from .components.value_function import ValueFunction
from .interfaces.learning_algorithm import LearningAlgorithm


class _LinearSplitK(torch.autograd.Function):
    """y = x W^T + b with the weight gradient computed by a split-K batched GEMM (tall-skinny reduction)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return torch.addmm(bias, x, weight.t())

    @staticmethod
    def backward(ctx, gy):
        from .fused_rollout import _tall_skinny_gram
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gy @ weight if ctx.needs_input_grad[0] else None
        return gx, _tall_skinny_gram(gy, x.contiguous()), gy.sum(dim=0)


class _BiasEluLayerNormRows(torch.autograd.Function):
    """h = LayerNorm(ELU(z + bias)) on (rows, width) training batches as ONE pass forward and ONE backward
    (csrc/sgbpo_critic.cu); the backward also yields the three column sums d gamma, d beta, d bias."""

    @staticmethod
    def forward(ctx, z, bias, gamma, beta):
        import ctypes as C
        from .. import _lib
        z = z.contiguous()
        M, W = z.shape
        h, e = torch.empty_like(z), torch.empty_like(z)
        stats = torch.empty(M, 2, dtype=z.dtype, device=z.device)
        _lib.check(_lib.lib().sgbpo_bias_elu_ln_fwd(_lib.dtype_code(z.dtype), M, W, _lib.ptr(z), _lib.ptr(bias), _lib.ptr(gamma),
                                                   _lib.ptr(beta), _lib.ptr(h), _lib.ptr(e), _lib.ptr(stats), _lib.stream_ptr()),
                   "sgbpo_bias_elu_ln_fwd")
        ctx.save_for_backward(e, stats, gamma)
        return h

    @staticmethod
    def backward(ctx, gy):
        from .. import _lib
        e, stats, gamma = ctx.saved_tensors
        gy = gy.contiguous()
        M, W = gy.shape
        dz = torch.empty_like(gy)
        sums = torch.zeros(3, W, dtype=gy.dtype, device=gy.device)
        _lib.check(_lib.lib().sgbpo_bias_elu_ln_bwd(_lib.dtype_code(gy.dtype), M, W, _lib.ptr(gy), _lib.ptr(e), _lib.ptr(stats),
                                                   _lib.ptr(gamma), _lib.ptr(dz), _lib.ptr(sums[0]), _lib.ptr(sums[1]),
                                                   _lib.ptr(sums[2]), _lib.stream_ptr()), "sgbpo_bias_elu_ln_bwd")
        return dz, sums[2], sums[0], sums[1]


class _MatmulSplitK(torch.autograd.Function):
    """z = x W^T (no bias) with the weight gradient through the split-K GEMM."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.save_for_backward(x, weight)
        return x @ weight.t()

    @staticmethod
    def backward(ctx, gz):
        from .fused_rollout import _tall_skinny_gram
        x, weight = ctx.saved_tensors
        gz = gz.contiguous()
        gx = gz @ weight if ctx.needs_input_grad[0] else None
        return gx, _tall_skinny_gram(gz, x.contiguous())


class _LayerNormRows(torch.autograd.Function):
    """LayerNorm over the last dim for (rows, width) training batches.  Forward and input-gradient are ATen's own
    kernels; the weight / bias gradients (column reductions over ~65 k rows, for which ATen's GammaBetaBackward
    kernel takes 0.5 ms per call) are two plain column sums."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        y, mean, rstd = torch.ops.aten.native_layer_norm(x, [x.shape[-1]], weight, bias, eps)
        ctx.save_for_backward(x, mean, rstd, weight, bias)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, mean, rstd, weight, bias = ctx.saved_tensors
        gy = gy.contiguous()
        gx = torch.ops.aten.native_layer_norm_backward(gy, x, [x.shape[-1]], mean, rstd, weight, bias, [True, False, False])[0]
        xhat = (x - mean) * rstd
        return gx, (gy * xhat).sum(dim=0), gy.sum(dim=0), None


class SHAC(LearningAlgorithm):
    GAMMA = 0.99
    MAX_GRAD_NORM = 1.0

    def __init__(self, env, policy_kwargs: dict, policy_optim_kwargs: dict, vf_kwargs: dict, vf_optim_kwargs: dict,
                 regularisation_coefficient: float, len_trajectories: int, polyak_target: float = 0.995,
                 td_weight: float = 0.95, vf_num_fits: int = 16, vf_fit_num_batches: int = 4, fused: str = "auto",
                 rng_order: str = "reference"):
        super().__init__(env, policy_kwargs, policy_optim_kwargs, vf_kwargs, vf_optim_kwargs,
                         regularisation_coefficient, False)
        self.len_trajectories = len_trajectories
        self.polyak_target, self.td_weight = polyak_target, td_weight
        self.vf_fit_num_batches, self.vf_num_fits = vf_fit_num_batches, vf_num_fits
        self.target_value_function = ValueFunction(self.env.obs_dim, **vf_kwargs)
        # Q3: the reference probes `safe_actions`, an attribute that never exists -> regularisation is dead code
        self.buffer = CoupledBuffer(len_trajectories + 1, self.env.num_envs, self.env.obs_dim, True,
                                    self.env.action_dim, store_safe_actions=hasattr(self.env, "safe_actions"))
        reset_observation, _ = self.env.reset()
        reset_value = self.target_value_function(reset_observation).squeeze(dim=1)
        self.buffer.reset(reset_observation, reset_value)
        self.interactions_per_episode = len_trajectories * self.env.num_envs
        self.fused = fused
        self.rng_order = rng_order      # "reference": draws in the reference's call order; "batched": 3 draws per episode
        self._fused_engine = None
        # data-parallel hook: called between backward and clip/Adam (distributed.allreduce_policy_grads)
        self.grad_sync = None
        # same for the critic: called after every critic batch's backward, before clip + Adam (distributed.allreduce_critic_grads)
        self.vf_grad_sync = None
        self.vf_graphs = True            # critic fit replayed as a CUDA graph when the optimizer allows (plain Adam on CUDA)
        self._vf_graph = None
        self._fused_tail = None          # resolved on first use: fused clip + Adam launch when the optimizer is plain Adam on CUDA

    # ---- episode ---------------------------------------------------------------------------------------
    def _learn_episode(self, eps: int) -> tuple[float, float, float]:
        self.buffer.reset()
        self.policy_optim.zero_grad()
        average_reward = self.collect_trajectories()
        policy_loss = self.update_policy()
        value_loss = self.update_value_function()
        self.update_target_value_function()
        return average_reward, policy_loss, value_loss

    def _engine(self):
        if self.fused == "off":
            return None
        if self._fused_engine is None:
            from .fused_rollout import FusedRollout
            self._fused_engine = FusedRollout.try_create(self)
            if self._fused_engine is not None:
                self._fused_engine.rng_mode = self.rng_order
            if self._fused_engine is None and self.fused == "on":
                raise RuntimeError("fused='on' but this env/safeguard/network combination has no fused rollout")
            if self._fused_engine is None:
                self.fused = "off"
        return self._fused_engine

    def collect_trajectories(self) -> float:
        engine = self._engine()
        if engine is not None:
            return engine.collect_trajectories()
        self.env.cut_computation_graph()
        reward_sum = torch.zeros(())
        for t in range(self.len_trajectories):
            action = self.policy(self.buffer.observations[self.buffer.t])
            observation, reward, terminated, truncated, info = self.env.step(action)
            terminal = terminated | truncated
            value = self.target_value_function(observation).squeeze(dim=1)
            if t == self.len_trajectories - 1:
                terminal = torch.ones_like(terminal)
            self.buffer.add(observation, reward, terminal, value, action, safe_action=None)
            reward_sum = reward_sum + reward.detach().sum()        # one host sync per episode, not per step (Q13)
        return float(reward_sum) / self.env.num_envs / self.len_trajectories

    def discount_weights(self) -> tuple[Tensor, Tensor]:
        """shac.py:131-139: per-row discount with restarts after terminals, and the normaliser."""
        terminals = self.buffer.terminals.tensor
        rows = self.len_trajectories + 2
        exponent = torch.arange(rows).view(-1, 1).repeat(1, self.env.num_envs)
        # shac.py:133-134 subtracts (end + 1) from every later row for EACH terminal row `end`, cumulatively
        # (so with two or more terminals before a row the exponent goes below the restart value -- kept as
        # shipped); an exclusive cumulative sum does the same without the python loop over nonzero()
        drops = torch.cumsum(torch.where(terminals, exponent + 1, torch.zeros_like(exponent)), dim=0)
        exponent = exponent - torch.cat([torch.zeros_like(drops[:1]), drops[:-1]], dim=0)
        discount = self.GAMMA ** exponent.to(torch.get_default_dtype())
        normalisation = self.env.num_envs * self.len_trajectories + terminals.count_nonzero()
        return discount, normalisation

    def policy_loss(self) -> Tensor:
        discount, normalisation = self.discount_weights()
        values = torch.where(self.buffer.terminals.tensor, self.buffer.values.tensor, self.buffer.rewards.tensor)
        loss = -(discount * values).sum() / normalisation
        if self.buffer.store_safe_actions:
            loss = loss + self.regularisation_coefficient * torch.nn.functional.mse_loss(
                self.buffer.safe_actions.tensor, self.buffer.actions.tensor)
        return loss

    def update_policy(self) -> float:
        engine = self._engine()
        if engine is not None:
            loss = engine.backward()
        else:
            loss = self.policy_loss()
            loss.backward()
        if self.grad_sync is not None:
            self.grad_sync(self)
        if self._fused_tail is None:
            from .. import optim as sgoptim
            self._fused_tail = sgoptim.supported(self.policy_optim)
        if self._fused_tail:       # shac.py:148-149 as one launch (csrc/sgbpo_optim.cu)
            from .. import optim as sgoptim
            sgoptim.clip_adam_step(self.policy_optim, self.MAX_GRAD_NORM)
        else:
            torch.nn.utils.clip_grad_norm_(self.policy.parameters(), self.MAX_GRAD_NORM)
            self.policy_optim.step()
        return float(loss)

    # ---- critic (next after the hot path; plain torch, vectorised over environments) -------
    def calculate_estimated_values(self) -> tuple[Tensor, Tensor]:
        """TD-lambda targets, shac.py:187-241.  All environments share the time index, so the reference's
        masked scatter loop becomes a backward scan over rows; output ORDER follows the reference
        (row H-1 first, environments ascending, non-terminal rows only)."""
        H, N = self.len_trajectories, self.env.num_envs
        rewards, values, terminals = self.buffer.rewards.tensor, self.buffer.values.tensor, self.buffer.terminals.tensor
        if rewards.is_cuda and rewards.dtype in (torch.float32, torch.float64):
            return self._estimated_values_kernel(rewards.detach(), values.detach(), terminals)
        td = torch.ones(N)
        avg = torch.zeros(N)
        term_ret = torch.zeros(N)
        est, obs = [], []
        for t in range(H - 1, -1, -1):
            estimate = ~terminals[t]
            init = terminals[t + 1] & estimate
            update = (~terminals[t + 1]) & estimate
            td = torch.where(init, torch.ones_like(td), td)
            avg = torch.where(init, torch.zeros_like(avg), avg)
            term_ret = torch.where(init, rewards[t] + self.GAMMA * values[t + 1], term_ret)
            td_u = td * self.td_weight
            avg_u = avg * self.td_weight * self.GAMMA + self.GAMMA * values[t + 1] \
                + (1 - td_u) / (1 - self.td_weight) * rewards[t]
            td = torch.where(update, td_u, td)
            avg = torch.where(update, avg_u, avg)
            term_ret = torch.where(update, rewards[t] + self.GAMMA * term_ret, term_ret)
            e = (1 - self.td_weight) * avg + td * term_ret
            est.append(e[estimate])
            obs.append(self.buffer.observations.tensor[t][estimate])
        est, obs = torch.cat(est), torch.cat(obs)
        total = H * N
        pad = total - est.shape[0]
        if pad:
            est = torch.cat([est, torch.zeros(pad)])
            obs = torch.cat([obs, torch.zeros(pad, self.env.obs_dim)])
        return est, obs

    def _estimated_values_kernel(self, rewards: Tensor, values: Tensor, terminals: Tensor) -> tuple[Tensor, Tensor]:
        """One launch of csrc/sgbpo_critic.cu (a thread scans one environment) + the reference's output order:
        row H-1 first, environments ascending, non-terminal rows only, zero-padded to H * N."""
        import ctypes as C
        from .. import _lib
        H, N = self.len_trajectories, self.env.num_envs
        rewards, values = rewards.contiguous(), values.contiguous()
        term_u8 = terminals.contiguous().view(torch.uint8)
        est = torch.empty(H, N, dtype=rewards.dtype, device=rewards.device)
        valid = torch.empty(H, N, dtype=torch.uint8, device=rewards.device)
        _lib.check(_lib.lib().sgbpo_td_lambda(_lib.dtype_code(rewards.dtype), N, H, _lib.ptr(rewards), _lib.ptr(values),
                                             _lib.ptr(term_u8), float(self.GAMMA), float(self.td_weight), _lib.ptr(est),
                                             _lib.ptr(valid), _lib.stream_ptr()), "sgbpo_td_lambda")
        obs = self.buffer.observations.tensor[:H].detach()
        if self._uniform_terminals:
            # every environment shares the terminal rows (truncation depends on the shared step counter only,
            # simulator.py:102-106): keep whole rows, no boolean gather, no host sync
            keep = [t for t in range(H - 1, -1, -1) if not self._terminal_rows_host[t]]
            idx = torch.tensor(keep, dtype=torch.int64, device=est.device)
            est_rows, obs_rows = est.index_select(0, idx).reshape(-1), obs.index_select(0, idx).reshape(-1, obs.shape[2])
        else:
            mask = valid.flip(0).bool()
            est_rows, obs_rows = est.flip(0)[mask], obs.flip(0)[mask]
        pad = H * N - est_rows.shape[0]
        if pad:
            est_rows = torch.cat([est_rows, est_rows.new_zeros(pad)])
            obs_rows = torch.cat([obs_rows, obs_rows.new_zeros(pad, obs_rows.shape[1])])
        return est_rows, obs_rows

    @property
    def _uniform_terminals(self) -> bool:
        eng = self._fused_engine
        return eng is not None and eng.last is not None and "terminal_rows" in eng.last and \
            self.buffer.terminals.tensor.shape[0] == len(eng.last["terminal_rows"])

    @property
    def _terminal_rows_host(self):
        return self._fused_engine.last["terminal_rows"]

    def update_value_function(self) -> float:
        with torch.no_grad():
            estimated_values, observations = self.calculate_estimated_values()
        if self._vf_graph_ok(estimated_values):
            return self._fit_value_function_graph(estimated_values, observations)
        batch_size = math.ceil(len(estimated_values) / self.vf_fit_num_batches)
        value_loss = torch.zeros(())
        for _ in range(self.vf_num_fits):
            for b in range(self.vf_fit_num_batches):
                lo, hi = b * batch_size, min((b + 1) * batch_size, len(estimated_values))
                if lo >= len(estimated_values):
                    break
                self.value_function_optim.zero_grad()
                pred = self.value_function(observations[lo:hi]).squeeze(dim=1)
                value_loss = (pred - estimated_values[lo:hi]).square().mean()
                value_loss.backward()
                if self.vf_grad_sync is not None:
                    self.vf_grad_sync(self)
                torch.nn.utils.clip_grad_norm_(self.value_function.parameters(), self.MAX_GRAD_NORM)
                self.value_function_optim.step()
        return float(value_loss)

    # ---- critic fit as one replayed CUDA graph (shac.py:154-184) ---------------------------------------------
    def _vf_graph_ok(self, est: Tensor) -> bool:
        from .. import optim as sgoptim
        # (the learning rate is a launch argument baked into the captured graph: constant schedule only)
        # a collective per batch inside the captured loop needs a capture-safe backend: opt in with vf_graphs = "collective"
        sync_ok = self.vf_grad_sync is None or self.vf_graphs == "collective"
        return (bool(self.vf_graphs) and sync_ok and est.is_cuda and sgoptim.supported(self.value_function_optim)
                and torch.is_grad_enabled() and self.VF_LEARNING_RATE_SCHEDULE == "constant")

    def _fit_value_function_graph(self, est: Tensor, obs: Tensor) -> float:
        """The vf_num_fits x vf_fit_num_batches loop (forward, MSE, backward, clip, Adam per batch) is the same
        ~2 000 launches every episode: captured once on static input buffers and replayed.  Clip + Adam are the
        fused launches of csrc/sgbpo_optim.cu on torch.optim.Adam's own state tensors."""
        from .. import optim as sgoptim
        vf, opt = self.value_function, self.value_function_optim
        sgoptim._ensure_state(opt)
        # the captured graph bakes in the parameter and optimizer-state storage: a load_state_dict() / checkpoint resume that
        # replaces them must trigger a re-capture
        ptrs = tuple((p.data_ptr(), opt.state[p]["exp_avg"].data_ptr(), opt.state[p]["exp_avg_sq"].data_ptr(),
                      opt.state[p]["step"].data_ptr()) for p in vf.parameters())
        key = (tuple(est.shape), tuple(obs.shape), est.dtype, float(opt.param_groups[0]["lr"]), self.vf_num_fits,
               self.vf_fit_num_batches, ptrs)
        g = self._vf_graph
        if g is None or g["key"] != key:
            g = dict(key=key, est=torch.empty_like(est), obs=torch.empty_like(obs), loss=torch.zeros((), dtype=est.dtype, device=est.device),
                     graph=None)
            self._vf_graph = g
        g["est"].copy_(est)
        g["obs"].copy_(obs)
        if g["graph"] is None:
            params = list(vf.parameters())
            sgoptim._ensure_state(opt)
            # warm-up (required before capture) runs the loop for real: snapshot and restore so the first call
            # performs exactly one fit
            snap_p = [p.detach().clone() for p in params]
            snap_s = [{k: v.clone() for k, v in opt.state[p].items()} for p in params]
            for p in params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._fit_body(g)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._fit_body(g)
            with torch.no_grad():
                for p, sp, ss in zip(params, snap_p, snap_s):
                    p.copy_(sp)
                    for k, v in ss.items():
                        opt.state[p][k].copy_(v)
            g["graph"] = graph
        g["graph"].replay()
        return float(g["loss"])

    def _vf_forward_train(self, x: Tensor) -> Tensor:
        """The critic's forward for training batches (tens of thousands of rows x width 128).  Same operations as the
        nn.Sequential, but Linear's weight gradient dY^T X (a 128 x 128 result reduced over ~65 k rows, for which
        cuBLAS launches a handful of CTAs) goes through the split-K GEMM of fused_rollout._tall_skinny_gram."""
        from .fused_rollout import _mlp_layout
        layout = _mlp_layout(self.value_function.model)
        if layout is None:
            return self.value_function(x)
        lin, norm = layout
        h = x
        fusedk = x.is_cuda and x.dtype in (torch.float32, torch.float64)
        for l, ln in zip(lin[:-1], norm):
            if fusedk:      # GEMM, then bias + ELU + LayerNorm in one pass (widths 32 / 64 / 128 by _mlp_layout)
                h = _BiasEluLayerNormRows.apply(_MatmulSplitK.apply(h, l.weight), l.bias, ln.weight, ln.bias)
            else:
                h = _LinearSplitK.apply(h, l.weight, l.bias)
                h = torch.nn.functional.elu(h)
                h = _LayerNormRows.apply(h, ln.weight, ln.bias, ln.eps)
        return _LinearSplitK.apply(h, lin[-1].weight, lin[-1].bias)

    def _fit_body(self, g):
        from .. import optim as sgoptim
        vf, opt = self.value_function, self.value_function_optim
        est, obs = g["est"], g["obs"]
        n = est.shape[0]
        batch_size = math.ceil(n / self.vf_fit_num_batches)
        for _ in range(self.vf_num_fits):
            for b in range(self.vf_fit_num_batches):
                lo, hi = b * batch_size, min((b + 1) * batch_size, n)
                if lo >= n:
                    break
                torch._foreach_zero_([p.grad for p in vf.parameters()])     # (vf_grad_sync may re-bind .grad to flat views)
                pred = self._vf_forward_train(obs[lo:hi]).squeeze(dim=1)
                value_loss = (pred - est[lo:hi]).square().mean()
                value_loss.backward()
                if self.vf_grad_sync is not None:
                    self.vf_grad_sync(self)
                sgoptim.clip_adam_step(opt, self.MAX_GRAD_NORM)
        g["loss"].copy_(value_loss.detach())

    def update_target_value_function(self):
        with torch.no_grad():
            targets, params = list(self.target_value_function.parameters()), list(self.value_function.parameters())
            torch._foreach_mul_(targets, self.polyak_target)                       # shac.py:248-251, two launches
            torch._foreach_add_(targets, params, alpha=1 - self.polyak_target)
