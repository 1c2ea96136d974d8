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
    if ndim == 3:
        return (v[1], v[2])            # (h, w) of a (d, h, w) triple; the depth axis is handled by depth_of()
    return (v[0], v[1])


def depth_of(v, ndim: int) -> int:
    """Depth component of a 3-D conv hyper-parameter (int or (d, h, w))."""
    if isinstance(v, (tuple, list)):
        v = tuple(int(i) for i in v)
        return v[0]
    return int(v)


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
        raise NotImplementedError("3-D inputs go through _kan_conv3d / _norm_act3d")

    def _from4d(self, y: torch.Tensor) -> torch.Tensor:
        return y.squeeze(2) if self.ndim == 1 else y

    # ---- 3-D layers: the volume convolution as kd depth-shifted 2-D convolutions on the CUDA kernels --------------------------
    #   z[:, :, do] = sum_k  conv2d_kan( x[:, :, do * sd + k * dd - pd],  W[:, :, k] )
    # For one depth tap k every output slice that has an input slice is one image of a 2-D batch (depth folded into the batch
    # axis), so a layer costs kd calls of the 2-D op, not D * kd.  Depth padding contributes nothing: the reference pads the
    # EXPANDED tensor (Conv3d's zero padding), not x, so out-of-range slices are skipped rather than evaluated at x = 0.
    def _kan_conv3d(self, spec2, x5: torch.Tensor, u5, beta, w_base5, w_basis5) -> torch.Tensor:
        if x5.dim() != 5:
            raise ValueError(f"expected a 5-D input [N, C, D, H, W], got {tuple(x5.shape)}")
        kd, sd = depth_of(self.kernel_size, 3), depth_of(self.stride, 3)
        pd, dd = depth_of(self.padding, 3), depth_of(self.dilation, 3)
        n, c, d, h, w = x5.shape
        do = (d + 2 * pd - dd * (kd - 1) - 1) // sd + 1
        if do <= 0:
            raise ValueError("3-D KAN convolution: empty output depth")
        z5 = None
        for k in range(kd):
            off = k * dd - pd                                  # input slice of output slice o: o * sd + off
            lo = max(0, -(off // sd)) if off < 0 else 0        # first o with o * sd + off >= 0
            while lo * sd + off < 0:
                lo += 1
            hi = min(do - 1, (d - 1 - off) // sd)              # last o with o * sd + off <= d - 1
            if hi < lo:
                continue
            cnt = hi - lo + 1
            sl = slice(lo * sd + off, (hi * sd + off) + 1, sd)

            def fold(t):                                       # [N, C, cnt, H, W] -> [N * cnt, C, H, W]
                return t[:, :, sl].permute(0, 2, 1, 3, 4).reshape(n * cnt, c, h, w).contiguous()

            xb = fold(x5)
            xs = None if u5 is None else fold(u5)
            wb = [wt[:, :, k].contiguous() for wt in w_base5]
            ws = [wt[:, :, k].contiguous() for wt in w_basis5]
            zk = KF.kan_conv(spec2, xb, xs, beta, wb, ws, self.precision)             # [N * cnt, Cout, Ho, Wo]
            zk = zk.reshape(n, cnt, zk.shape[1], zk.shape[2], zk.shape[3]).permute(0, 2, 1, 3, 4)
            zk = F.pad(zk, (0, 0, 0, 0, lo, do - 1 - hi))
            z5 = zk if z5 is None else z5 + zk
        if z5 is None:
            raise ValueError("3-D KAN convolution: no depth tap touches the input")
        return z5.contiguous()

    def _norm_act3d(self, z5: torch.Tensor, norms, out_act: int, alphas=()) -> torch.Tensor:
        """InstanceNorm3d / BatchNorm3d (+ activation) of [N, C, D, H, W]: the statistics run over D * H * W, which is the plane of
        the 2-D kernels when the volume is viewed as [N, C, D * H, W]."""
        n, c, d, h, w = z5.shape
        return self._norm_act(z5.reshape(n, c, d * h, w), norms, out_act, alphas).reshape(n, c, d, h, w)

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
