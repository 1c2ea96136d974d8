"""Shared plumbing of the drop-in KAN convolution modules (not part of the reference API)."""
from __future__ import annotations

from inspect import signature
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from .. import functional as KF

_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
_DROPOUT = {1: nn.Dropout1d, 2: nn.Dropout2d, 3: nn.Dropout3d}


def pair(v, ndim: int, fill: int = 1) -> Tuple[int, int]:
    """Conv hyper-parameter -> (h, w) pair.  A 1-D layer is run as a 2-D layer of height 1: ``fill`` is the value of the
    dummy height axis - 1 for kernel size / stride / dilation, 0 for padding."""
    if isinstance(v, (tuple, list)):
        v = tuple(int(i) for i in v)
        if len(v) == 1:
            v = (v[0],) * ndim
    else:
        v = (int(v),) * ndim
    if ndim == 1:
        return (fill, v[0]) if len(v) == 1 else (fill, v[-1])
    return (v[0], v[1])


def act_kind(module: nn.Module) -> int:
    if isinstance(module, nn.Identity):
        return L.ACT_IDENTITY
    if isinstance(module, nn.GELU):
        if getattr(module, "approximate", "none") != "none":
            raise NotImplementedError("only the exact (erf) GELU base activation is implemented")
        return L.ACT_GELU
    if isinstance(module, nn.SiLU):
        return L.ACT_SILU
    raise NotImplementedError(f"base_activation {type(module).__name__} is not implemented in the CUDA kernels "
                              "(supported: nn.GELU, nn.SiLU, None)")


def check_groups(groups: int, input_dim: int, output_dim: int) -> None:
    # same messages as the reference (kan_layers.py:148-153)
    if groups <= 0:
        raise ValueError('groups must be a positive integer')
    if input_dim % groups != 0:
        raise ValueError('input_dim must be divisible by groups')
    if output_dim % groups != 0:
        raise ValueError('output_dim must be divisible by groups')


def filter_norm_kwargs(norm_class, norm_kwargs: dict) -> dict:
    valid = signature(norm_class).parameters
    return {k: v for k, v in norm_kwargs.items() if k in valid}


class KANConvBase(nn.Module):
    """Common forward plumbing: 1-D/2-D reshaping, norm-module introspection, BatchNorm running statistics."""

    ndim: int = 2
    precision: Optional[str] = None      # None -> global kanconv_b200.get_precision()

    def _to4d(self, x: torch.Tensor) -> torch.Tensor:
        if self.ndim == 2:
            if x.dim() != 4:
                raise ValueError(f"expected a 4-D input [N, C, H, W], got {tuple(x.shape)}")
            return x
        if self.ndim == 1:
            if x.dim() != 3:
                raise ValueError(f"expected a 3-D input [N, C, L], got {tuple(x.shape)}")
            return x.unsqueeze(2)
        raise NotImplementedError("3-D KAN convolutions are not implemented in the CUDA kernels")

    def _from4d(self, y: torch.Tensor) -> torch.Tensor:
        return y.squeeze(2) if self.ndim == 1 else y

    def _w4d(self, w: torch.Tensor) -> torch.Tensor:
        return w.unsqueeze(2) if self.ndim == 1 else w

    @staticmethod
    def _norm_spec(norms: Sequence[nn.Module], out_act: int, training: bool):
        """-> (NormSpec, gammas, betas, given_mean, given_rstd) for a ModuleList of per-group norm modules."""
        m0 = norms[0]
        groups = len(norms)
        gammas: List[torch.Tensor] = []
        betas: List[torch.Tensor] = []
        given_mean = given_rstd = None
        if isinstance(m0, nn.modules.instancenorm._InstanceNorm):
            if m0.track_running_stats:
                raise NotImplementedError("InstanceNorm with track_running_stats=True is not implemented")
            kind, affine, eps, use_batch = L.NORM_INSTANCE, bool(m0.affine), float(m0.eps), True
        elif isinstance(m0, nn.modules.batchnorm._BatchNorm):
            kind, affine, eps = L.NORM_BATCH, bool(m0.affine), float(m0.eps)
            use_batch = training or not m0.track_running_stats
            if not use_batch:
                given_mean = torch.cat([m.running_mean for m in norms])
                given_rstd = torch.rsqrt(torch.cat([m.running_var for m in norms]) + eps)
        elif isinstance(m0, nn.Identity):
            kind, affine, eps, use_batch = L.NORM_NONE, False, 0.0, True
        else:
            raise NotImplementedError(f"norm layer {type(m0).__name__} is not implemented in the CUDA kernels "
                                      "(supported: InstanceNorm, BatchNorm, Identity)")
        if affine:
            gammas = [m.weight for m in norms]
            betas = [m.bias for m in norms]
        return KF.NormSpec(kind, out_act, groups, affine, eps, use_batch), gammas, betas, given_mean, given_rstd

    @staticmethod
    @torch.no_grad()
    def _update_running_stats(norms: Sequence[nn.Module], mean: torch.Tensor, rstd: torch.Tensor, count: int) -> None:
        """BatchNorm bookkeeping (momentum update with the unbiased variance), like nn.BatchNorm2d in train mode."""
        for g, m in enumerate(norms):
            if not isinstance(m, nn.modules.batchnorm._BatchNorm) or not m.track_running_stats:
                continue
            var = (1.0 / (rstd[g] * rstd[g]) - m.eps).clamp_min(0.0)
            if count > 1:
                var = var * (count / (count - 1.0))
            m.num_batches_tracked += 1
            mom = m.momentum if m.momentum is not None else 1.0 / float(m.num_batches_tracked)
            m.running_mean.mul_(1 - mom).add_(mean[g], alpha=mom)
            m.running_var.mul_(1 - mom).add_(var, alpha=mom)

    def _conv_norm_act(self, spec, x4: torch.Tensor, beta, w_base, w_basis, norms, out_act: int, alphas=()):
        """out_act(norm(kan_conv(x))) as one autograd node (functional.kan_layer): same forward kernels as kan_conv + norm_act,
        and a backward in which the norm kernel emits dz directly in the operand layout of the tensor-core dgrad / wgrad."""
        nspec, gammas, betas, gm, gr = self._norm_spec(norms, out_act, self.training)
        y, mean, rstd = KF.kan_layer(spec, nspec, x4, beta, w_base, w_basis, gammas, betas, alphas, gm, gr, self.precision)
        if nspec.norm == L.NORM_BATCH and self.training:
            self._update_running_stats(norms, mean, rstd, y.shape[0] * y.shape[2] * y.shape[3])
        return y

    def _norm_act(self, z: torch.Tensor, norms, out_act: int, alphas=()):
        spec, gammas, betas, gm, gr = self._norm_spec(norms, out_act, self.training)
        y, mean, rstd = KF.norm_act(spec, z, gammas, betas, alphas, gm, gr)
        if spec.norm == L.NORM_BATCH and self.training:
            self._update_running_stats(norms, mean, rstd, z.shape[0] * z.shape[2] * z.shape[3])
        return y


def make_dropout(ndim: int, p: float) -> Optional[nn.Module]:
    return _DROPOUT[ndim](p=p) if p > 0 else None


def conv_class(ndim: int):
    return _CONV[ndim]
