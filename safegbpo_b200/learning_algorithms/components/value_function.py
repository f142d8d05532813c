"""Critic (reference: src/learning_algorithms/components/value_function.py:9-66)."""
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .policy import build_mlp


class ValueFunction(nn.Module):
    def __init__(self, input_dim: int, net_arch: Optional[list] = None, activation_fn: str = "nn.ELU()",
                 layer_norm: bool = True):
        super().__init__()
        self.net_arch = list(net_arch) if net_arch is not None else [64, 64]
        self.activation_fn, self.layer_norm = activation_fn, layer_norm
        self.model = build_mlp([input_dim] + self.net_arch, 1, activation_fn, layer_norm)

    def forward(self, x: Tensor, a: Optional[Tensor] = None) -> Tensor:
        if a is not None:
            x = torch.cat([x, a], dim=1)
        if not torch.is_grad_enabled() and x.is_cuda and x.dim() == 2:
            out = self._rows_kernel(x)
            if out is not None:
                return out
        return self.model(x)

    def _rows_kernel(self, x: Tensor) -> Optional[Tensor]:
        """Inference (no autograd) on CUDA: one launch of the batched MLP kernel (csrc/sgbpo_mlp_tc.cu on the tensor
        cores for fp32 width-128 nets, the FMA tiles of csrc/sgbpo_rollout_impl.cuh otherwise) instead of torch's
        ~3 launches per layer.  Returns None for layouts the kernels do not cover (torch evaluates those)."""
        import ctypes as C
        from ... import _lib
        from ..fused_rollout import PackedMlp, _mlp_layout, mlp_rows_fits
        pack = self.__dict__.get("_pack")
        if pack is None:
            if _mlp_layout(self.model) is None:
                self.__dict__["_pack"] = False
                return None
            pack = PackedMlp(self.model)
            self.__dict__["_pack"] = pack
        if pack is False or x.dtype not in (torch.float32, torch.float64) or x.shape[1] != pack.in_dim:
            return None
        if next(self.model.parameters()).dtype != x.dtype:
            return None
        if not mlp_rows_fits(x.dtype, pack.in_dim, pack.width, pack.n_hidden):
            return self._rows_layered(x, pack)
        pack.refresh()
        x = x.contiguous()
        out = torch.empty(x.shape[0], 1, dtype=x.dtype, device=x.device)
        _lib.check(_lib.lib().sgbpo_mlp_rows(_lib.dtype_code(x.dtype), x.shape[0], C.byref(pack.desc), C.c_void_p(x.data_ptr()),
                                             C.c_void_p(out.data_ptr()), _lib.stream_ptr()), "sgbpo_mlp_rows")
        return out

    def _rows_layered(self, x: Tensor, pack) -> Tensor:
        """Networks whose parameters do not fit one SM's shared memory (float64 [128]x3): library GEMM per layer plus
        ONE fused pass for bias + ELU + LayerNorm (csrc/sgbpo_critic.cu) instead of ATen's separate passes (its
        float64 LayerNorm statistics kernel alone takes 0.8 ms per call on 262 k rows)."""
        from ... import _lib
        h = x.contiguous()
        for lin, norm in zip(pack.lin[:-1], pack.norm):
            z = h @ lin.weight.t()
            h, e = torch.empty_like(z), torch.empty_like(z)
            stats = torch.empty(z.shape[0], 2, dtype=z.dtype, device=z.device)
            _lib.check(_lib.lib().sgbpo_bias_elu_ln_fwd(_lib.dtype_code(z.dtype), z.shape[0], z.shape[1], _lib.ptr(z), _lib.ptr(lin.bias),
                                                       _lib.ptr(norm.weight), _lib.ptr(norm.bias), _lib.ptr(h), _lib.ptr(e),
                                                       _lib.ptr(stats), _lib.stream_ptr()), "sgbpo_bias_elu_ln_fwd")
        return torch.addmm(pack.lin[-1].bias, h, pack.lin[-1].weight.t())
