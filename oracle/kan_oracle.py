"""CPU oracle for the KAN-convolution hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the arithmetic of the
reference's four convolutional KAN layers.  It is the *checker* for the CUDA path in
``convolutional-kan-for-image-classification_b200/``; it is never imported by the product
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.

Parity status: **pinned** against the live reference.  ``tests/golden/make_golden.py`` imports
the unmodified reference from ``/root/reference`` (possible only in the dev container), runs it in
fp64 and fp32 on seeded inputs and stores inputs / weights / outputs / gradients under
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks every function here against those
fixtures (and against the SURVEY Appendix-E checksums).  The reference itself ships no tests
or golden vectors.

Reference call sites restated here (paths relative to the reference tree):
  * B-spline basis, Cox-de Boor recursion ........ layers/kan_layers.py:203-233
  * KANConvNDLayer.forward_kan / forward ......... layers/kan_layers.py:197-258
  * ChebyKANConvNDLayer.forward_ChebyKAN ......... layers/cheby_kan_layers.py:91-111
  * GRAMKANConvNDLayer.beta/gram_poly/forward_kag  layers/gram_kan_layers.py:150-199
  * FastKANConvNDLayer.forward_fast_kan .......... layers/fast_kan_layers.py:100-120
  * RadialBasisFunction.forward .................. utils/utils.py:19-33
  * KANLayer.forward (fully-connected layer) ..... layers/kan_layers.py:8-114
  * KANConv1DLayer (ndim = 1 binding) ............ layers/kan_layers.py:287-297
  * *KANConv3DLayer (ndim = 3 bindings) .......... layers/kan_layers.py:261-271, cheby_kan_layers.py:114-121,
                                                   gram_kan_layers.py:202-209, fast_kan_layers.py:123-134
  * Hermite / Gegenbauer / Laguerre / Lucas / Fibonacci / Bessel / Taylor / Legendre / Jacobi KANConvNDLayer
    (three-term-recurrence polynomial families) .. layers/<family>_kan_layers.py, lines cited at recurrence_polys()
  * KAN (MLP of KANLayers) ....................... models/kans.py:300-327
  * VGG.make_layers / forward .................... models/kan_vgg.py:40-188

The third-party arithmetic underneath (conv2d, instance/batch norm, PReLU, GELU/SiLU) lives in
PyTorch itself (the reference pins no version; the oracle runs on the torch of this image).
Backward is autograd-derived here exactly as in the reference (it has no hand-written backward).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = [
    "make_knots", "bspline_basis", "cheby_basis", "gram_basis", "rbf_basis",
    "kan_conv2d", "cheby_conv2d", "gram_conv2d", "fastkan_conv2d",
    "OracleKANConv2D", "OracleChebyKANConv2D", "OracleGRAMKANConv2D", "OracleFastKANConv2D",
    "OracleKANConv3D", "OracleChebyKANConv3D", "OracleGRAMKANConv3D", "OracleFastKANConv3D",
    "OraclePolyKANConv", "OracleDegreeMajorPolyKANConv", "recurrence_polys", "polykan_conv", "degree_major_polykan_conv",
    "OracleKANConv1D", "OracleKANLayer", "OracleKAN", "kan_conv1d", "kan_linear",
    "OracleVGG", "VGG_CFGS", "conv_flops",
]

IntOr2 = Union[int, Tuple[int, int]]


# ----------------------------------------------------------------------------------------------
# basis functions
# ----------------------------------------------------------------------------------------------
def make_knots(grid_size: int, spline_order: int, grid_range: Sequence[float]) -> torch.Tensor:
    """fp32 knot vector of kan_layers.py:184-190 (G + 2K + 1 uniformly spaced knots)."""
    lo, hi = float(grid_range[0]), float(grid_range[1])
    h = (hi - lo) / grid_size
    return torch.linspace(lo - h * spline_order, hi + h * spline_order,
                          grid_size + 2 * spline_order + 1, dtype=torch.float32)


def bspline_basis(x: torch.Tensor, knots: torch.Tensor, order: int) -> torch.Tensor:
    """Cox-de Boor recursion of kan_layers.py:209-233.  Returns [..., G+K] (trailing basis axis).

    Order-0 indicators are half-open [t_i, t_{i+1}); the recursion keeps the reference's operation
    order ``(x - t_i) / (t_{i+k} - t_i) * B`` and its zero-denominator guards.
    """
    # The reference keeps the knot vector as an fp32 attribute even when the module is .double()'d, so
    # knot differences are formed in the knots' own dtype and only then promoted against x.
    t = knots.to(device=x.device)
    xe = x.unsqueeze(-1)
    b = ((xe >= t[:-1]) & (xe < t[1:])).to(x.dtype)
    for k in range(1, order + 1):
        d_left = t[k:-1] - t[:-(k + 1)]
        d_right = t[k + 1:] - t[1:-k]
        d_left = torch.where(d_left == 0, torch.ones_like(d_left), d_left)
        d_right = torch.where(d_right == 0, torch.ones_like(d_right), d_right)
        b = (xe - t[:-(k + 1)]) / d_left * b[..., :-1] + (t[k + 1:] - xe) / d_right * b[..., 1:]
    return b


def cheby_basis(x: torch.Tensor, degree: int, eps: float = 1e-7) -> torch.Tensor:
    """cos(d * acos(clamp(tanh x))) for d = 0..degree (cheby_kan_layers.py:93-96).  [..., D+1]."""
    theta = torch.acos(torch.clamp(torch.tanh(x), -1 + eps, 1 - eps))
    d = torch.arange(0, degree + 1, device=x.device)
    return torch.cos(theta.unsqueeze(-1) * d)


def gram_beta(n: int, m: int, beta_weights: torch.Tensor) -> torch.Tensor:
    """gram_kan_layers.py:150-153."""
    return (((m + n) * (m - n) * n ** 2) / (m ** 2 / (4.0 * n ** 2 - 1.0))) * beta_weights[n]


def gram_basis(t: torch.Tensor, degree: int, beta_weights: torch.Tensor) -> List[torch.Tensor]:
    """Gram three-term recurrence of gram_kan_layers.py:155-170; list of D+1 tensors shaped like t."""
    p0 = torch.ones_like(t)
    if degree == 0:
        return [p0]
    p1 = t
    out = [p0, p1]
    for i in range(2, degree + 1):
        p2 = t * p1 - gram_beta(i - 1, i, beta_weights) * p0
        out.append(p2)
        p0, p1 = p1, p2
    return out


def rbf_basis(u: torch.Tensor, grid: torch.Tensor, denominator: float) -> torch.Tensor:
    """exp(-((u - g_j) / den)^2), utils/utils.py:32-33.  [..., G]."""
    return torch.exp(-((u.unsqueeze(-1) - grid.to(u.dtype)) / denominator) ** 2)


def recurrence_polys(family: str, t: torch.Tensor, degree: int, **fam) -> List[torch.Tensor]:
    """Polynomials [p_0(t), ..., p_D(t)] of the reference's three-term-recurrence families, each written as its own file writes it:
      hermite     hermite_kan_layers.py:136-146     H_0 = 1, H_1 = 2t, H_i = 2t H_{i-1} - 2(i-1) H_{i-2}
      gegenbauer  gegenbauer_kan_layers.py:143-153  C_0 = 1, C_1 = 2at, C_{n+1} = (2(n+a) t C_n - (n+2a-1) C_{n-1}) / (n+1)
      laguerre    laguerre_kan_layers.py:151-163    L_0 = 1, L_1 = 1+a-t, L_k = ((2(k-1)+1+a-t) L_{k-1} - (k-1+a) L_{k-2}) / k
      lucas       lucas_kan_layers.py:163-171       L_0 = 2, L_1 = t, L_i = t L_{i-1} + L_{i-2}
      fibonacci   fibonacci_kan_layers.py:152-160   F_0 = 0, F_1 = 1, F_i = t F_{i-1} + F_{i-2}
      bessel      bessel_kan_layers.py:144-153      y_0 = 1, y_1 = t+1, y_i = (2i-1) t y_{i-1} + y_{i-2}
      taylor      taylor_kan_layers.py:143-148      t^0 .. t^(degree-1)   (``degree`` counts the TERMS)
      legendre    legendre_kan_layers.py:108-121    P_0 = 1, P_1 = t, P_{n+1} = ((2n+1) t P_n - n P_{n-1}) / (n+1)
      jacobi      jacobi_kan_layers.py:119-137      P_1 = ((a-b) + (a+b+2) t) / 2, theta recurrences"""
    one = torch.ones_like(t)
    if family == "taylor":
        polys = [one]
        if degree > 1:
            polys.append(t)
            for _ in range(2, degree):
                polys.append(polys[-1] * t)
        return polys
    if family == "fibonacci":
        polys = [torch.zeros_like(t)]
    elif family == "lucas":
        polys = [2 * one]
    else:
        polys = [one]
    if degree < 1:
        return polys
    if family == "hermite":
        polys.append(2 * t)
        for i in range(2, degree + 1):
            polys.append(2 * t * polys[i - 1] - 2 * (i - 1) * polys[i - 2])
    elif family == "gegenbauer":
        a = fam["alpha_param"]
        polys.append(2 * a * t)
        for n in range(1, degree):
            polys.append((2 * (n + a) * t * polys[n] - (n + 2 * a - 1) * polys[n - 1]) / (n + 1))
    elif family == "laguerre":
        a = fam["alpha"]
        polys.append((1 + a) - t)
        for k in range(2, degree + 1):
            polys.append(((2 * (k - 1) + 1 + a - t) * polys[k - 1] - (k - 1 + a) * polys[k - 2]) / k)
    elif family in ("lucas", "fibonacci"):
        polys.append(t if family == "lucas" else one)
        for i in range(2, degree + 1):
            polys.append(t * polys[i - 1] + polys[i - 2])
    elif family == "bessel":
        polys.append(t + 1)
        for i in range(2, degree + 1):
            polys.append((2 * i - 1) * t * polys[i - 1] + polys[i - 2])
    elif family == "legendre":
        polys.append(t)
        for n in range(1, degree):
            polys.append(((2.0 * n + 1.0) * t * polys[-1] - n * polys[-2]) / (n + 1.0))
    elif family == "jacobi":
        a, b = fam["a"], fam["b"]
        polys.append(((a - b) + (a + b + 2) * t) / 2)
        for i in range(2, degree + 1):
            th_k = (2 * i + a + b) * (2 * i + a + b - 1) / (2 * i * (i + a + b))
            th_k1 = (2 * i + a + b - 1) * (a * a - b * b) / (2 * i * (i + a + b) * (2 * i + a + b - 2))
            th_k2 = (i + a - 1) * (i + b - 1) * (2 * i + a + b) / (i * (i + a + b) * (2 * i + a + b - 2))
            polys.append((th_k * t + th_k1) * polys[i - 1] - th_k2 * polys[i - 2])
    else:
        raise ValueError(family)
    return polys


def _act(name: Optional[str]) -> Callable[[torch.Tensor], torch.Tensor]:
    if name is None or name == "identity":
        return lambda v: v
    if name == "gelu":
        return F.gelu            # exact erf GELU == nn.GELU() default
    if name == "silu":
        return F.silu
    raise ValueError(f"unsupported base activation {name!r}")


def _norm(z: torch.Tensor, kind: str, eps: float, weight, bias, running=None, training=True):
    if kind == "instance":
        return F.instance_norm(z, None, None, weight, bias, True, 0.1, eps)
    if kind == "batch":
        rm, rv = running if running is not None else (None, None)
        return F.batch_norm(z, rm, rv, weight, bias, training or rm is None, 0.1, eps)
    if kind == "none":
        return z
    raise ValueError(kind)


def _convnd(x, w, stride, padding, dilation):
    """nn.Conv1d / nn.Conv2d / nn.Conv3d forward (bias-free, groups = 1) chosen by the rank of the filter: the reference binds the
    same N-D layer body to the three conv classes (kan_layers.py:261-297 and the sibling files' *1D/*2D/*3D classes)."""
    return (F.conv1d, F.conv2d, F.conv3d)[w.dim() - 3](x, w, None, stride, padding, dilation)


def _expand(basis: torch.Tensor) -> torch.Tensor:
    """[N, C, H, W, nb] -> [N, C*nb, H, W] with expanded channel c*nb + j (kan_layers.py:236-237)."""
    return basis.movedim(-1, 2).flatten(1, 2)


# ----------------------------------------------------------------------------------------------
# one-group layer functions
# ----------------------------------------------------------------------------------------------
def kan_conv2d(x, w_base, w_spline, prelu_weight, knots, spline_order, act="gelu",
               stride: IntOr2 = 1, padding: IntOr2 = 0, dilation: IntOr2 = 1,
               norm="instance", eps=1e-5, norm_weight=None, norm_bias=None, running=None, training=True):
    """One group of KANConvNDLayer.forward_kan (kan_layers.py:197-247), ndim = 2, no dropout.
    ``running`` = (running_mean, running_var) of a BatchNorm2d norm layer (updated in training, used in eval)."""
    base = _convnd(_act(act)(x), w_base, stride, padding, dilation)
    phi = _expand(bspline_basis(x, knots, spline_order))
    z = base + _convnd(phi, w_spline, stride, padding, dilation)
    return F.prelu(_norm(z, norm, eps, norm_weight, norm_bias, running, training), prelu_weight)


def cheby_conv2d(x, w_poly, degree, stride=1, padding=0, dilation=1,
                 norm="instance", eps=1e-5, norm_weight=None, norm_bias=None):
    """One group of ChebyKANConvNDLayer.forward_ChebyKAN (cheby_kan_layers.py:91-101)."""
    phi = _expand(cheby_basis(x, degree))
    return _norm(_convnd(phi, w_poly, stride, padding, dilation), norm, eps, norm_weight, norm_bias)


def gram_conv2d(x, w_base, w_poly, beta_weights, degree, stride=1, padding=0, dilation=1,
                norm="instance", eps=1e-5, norm_weight=None, norm_bias=None, tanh_scale=None):
    """One group of GRAMKANConvNDLayer.forward_kag (gram_kan_layers.py:172-189); SiLU everywhere.
    ``tanh_scale`` ([N, C, 1, 1], entries 0 or 1/(1-p)) stands for the Dropout2d the reference applies to tanh(x)
    (:178-179) with a mask chosen by the caller."""
    base = _convnd(F.silu(x), w_base, stride, padding, dilation)
    t = torch.tanh(x)
    if tanh_scale is not None:
        t = t * tanh_scale
    polys = gram_basis(t, degree, beta_weights)
    phi = F.silu(torch.cat(polys, dim=1))               # degree-major: channel d*C + c
    z = _convnd(phi, w_poly, stride, padding, dilation) + base
    return F.silu(_norm(z, norm, eps, norm_weight, norm_bias))


def fastkan_conv2d(x, w_base, w_spline, grid, denominator, act="silu", stride=1, padding=0, dilation=1,
                   norm="instance", eps=1e-5, norm_weight=None, norm_bias=None, running=None,
                   training=True):
    """One group of FastKANConvNDLayer.forward_fast_kan (fast_kan_layers.py:100-111).

    NB the normalisation is applied to the *input* of the RBF branch; there is no output norm/act.
    """
    base = _convnd(_act(act)(x), w_base, stride, padding, dilation)
    u = _norm(x, norm, eps, norm_weight, norm_bias, running, training)
    phi = _expand(rbf_basis(u, grid, denominator))
    return base + _convnd(phi, w_spline, stride, padding, dilation)


def polykan_conv(family, x, w_base, w_poly, prelu_weight, degree, act="gelu", stride=1, padding=0, dilation=1,
                 norm="instance", eps=1e-5, norm_weight=None, norm_bias=None, running=None, training=True, **fam):
    """One group of the template-A layers (hermite_kan_layers.py:152-161 and the same method of the gegenbauer / laguerre /
    lucas / fibonacci / bessel / taylor files): PReLU(norm(conv(act(x)) + conv(P(tanh x)))), expanded channel c*nb + j
    (``basis.view(batch, channels * (degree + 1), ...)`` of a [B, C, D+1, ...] tensor)."""
    base = _convnd(_act(act)(x), w_base, stride, padding, dilation)
    phi = torch.stack(recurrence_polys(family, torch.tanh(x), degree, **fam), dim=2).flatten(1, 2)
    z = base + _convnd(phi, w_poly, stride, padding, dilation)
    return F.prelu(_norm(z, norm, eps, norm_weight, norm_bias, running, training), prelu_weight)


def degree_major_polykan_conv(family, x, w_base, w_poly, degree, out_act="silu", stride=1, padding=0, dilation=1,
                              norm="instance", eps=1e-5, norm_weight=None, norm_bias=None, **fam):
    """One group of LegendreKANConvNDLayer.forward_kal (legendre_kan_layers.py:123-152) / JacobiKANConvNDLayer.forward_kaj
    (jacobi_kan_layers.py:139-168): base conv on x itself, polynomials concatenated degree-major (channel j*C + c), output
    activation after the norm.  Legendre normalises x with the min / max of the whole group tensor, Jacobi with tanh."""
    base = _convnd(x, w_base, stride, padding, dilation)
    if family == "legendre":
        t = 2 * (x - x.min()) / (x.max() - x.min()) - 1
    else:
        t = torch.tanh(x)
    phi = torch.cat(recurrence_polys(family, t, degree, **fam), dim=1)
    z = base + _convnd(phi, w_poly, stride, padding, dilation)
    return _act(out_act)(_norm(z, norm, eps, norm_weight, norm_bias))


def kan_conv1d(x, w_base, w_spline, prelu_weight, knots, spline_order, act="gelu", stride=1, padding=0, dilation=1,
               eps=1e-5, norm_weight=None, norm_bias=None):
    """One group of KANConvNDLayer.forward_kan with ndim = 1 (kan_layers.py:197-247 bound by KANConv1DLayer :287-297):
    x [N, C, L]; nn.Conv1d / InstanceNorm1d underneath."""
    base = F.conv1d(_act(act)(x), w_base, None, stride, padding, dilation)
    phi = bspline_basis(x, knots, spline_order).movedim(-1, 2).flatten(1, 2)      # [N, C*nb, L], channel c*nb + j
    z = base + F.conv1d(phi, w_spline, None, stride, padding, dilation)
    return F.prelu(F.instance_norm(z, None, None, norm_weight, norm_bias, True, 0.1, eps), prelu_weight)


def kan_linear(x, base_weight, spline_weight, ln_weight, ln_bias, prelu_weight, knots, spline_order, act="gelu", eps=1e-5):
    """KANLayer.forward (kan_layers.py:46-114): x [B, in]; base_weight [out, in]; spline_weight [out, in, nb];
    LayerNorm over the output features, then PReLU."""
    base = F.linear(_act(act)(x), base_weight)
    bases = bspline_basis(x, knots, spline_order)                                   # [B, in, nb]
    spline = F.linear(bases.reshape(x.shape[0], -1), spline_weight.reshape(spline_weight.shape[0], -1))
    return F.prelu(F.layer_norm(base + spline, (base_weight.shape[0],), ln_weight, ln_bias, eps), prelu_weight)


# ----------------------------------------------------------------------------------------------
# oracle modules (state_dict keys == reference keys)
# ----------------------------------------------------------------------------------------------
def _pair(v: IntOr2) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _act_name(base_activation) -> Optional[str]:
    if base_activation is None:
        return None
    if isinstance(base_activation, str):
        return base_activation.lower()
    n = getattr(base_activation, "__name__", type(base_activation).__name__).lower()
    if n in ("gelu", "silu", "identity"):
        return n
    raise ValueError(f"oracle supports GELU/SiLU/Identity base activations, got {base_activation}")


def _norm_kind(norm_layer) -> str:
    if norm_layer is None:
        return "none"
    n = getattr(norm_layer, "__name__", str(norm_layer))
    if "InstanceNorm" in n:
        return "instance"
    if "BatchNorm" in n:
        return "batch"
    raise ValueError(f"oracle supports InstanceNorm2d/BatchNorm2d, got {norm_layer}")


class _Weight(nn.Module):
    """Holder exposing ``.weight`` so keys read ``base_conv.0.weight`` like nn.Conv2d's."""

    def __init__(self, shape):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape))
        # nn.Conv2d's own reset_parameters() draw, so that a same-seed construction consumes the RNG exactly
        # like the reference ctor does before its explicit re-initialisation (kan_layers.py:159-195).
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))


class _OracleBase(nn.Module):
    _nd = 2                       # spatial rank; the *3D subclasses set 3 (Conv3d / InstanceNorm3d / BatchNorm3d underneath)

    def _kshape(self, kernel_size):
        if self._nd == 2:
            return _pair(kernel_size)
        return (int(kernel_size),) * self._nd if isinstance(kernel_size, int) else tuple(int(k) for k in kernel_size)

    def _check_groups(self, cin, cout, groups):
        if groups <= 0:
            raise ValueError("groups must be a positive integer")
        if cin % groups != 0:
            raise ValueError("input_dim must be divisible by groups")
        if cout % groups != 0:
            raise ValueError("output_dim must be divisible by groups")

    def _make_norm(self, ch, kind, affine, groups):
        mods = []
        for _ in range(groups):
            if kind == "instance":
                mods.append({1: nn.InstanceNorm1d, 2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}[self._nd](ch, affine=affine))
            elif kind == "batch":
                mods.append({1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}[self._nd](ch))
            else:
                mods.append(nn.Identity())
        return nn.ModuleList(mods)

    @staticmethod
    def _nw(norm_mod):
        return getattr(norm_mod, "weight", None), getattr(norm_mod, "bias", None)


class OracleKANConv2D(_OracleBase):
    def __init__(self, input_dim, output_dim, kernel_size, spline_order=3, groups=1, padding=0, stride=1,
                 dilation=1, grid_size=5, base_activation="gelu", grid_range=(-1, 1), norm_layer=nn.InstanceNorm2d,
                 affine=False):
        super().__init__()
        self._check_groups(input_dim, output_dim, groups)
        ks = self._kshape(kernel_size)
        cg, og = input_dim // groups, output_dim // groups
        self.groups, self.cg, self.og = groups, cg, og
        self.spline_order, self.act = spline_order, _act_name(base_activation)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.norm_kind = _norm_kind(norm_layer)
        self.base_conv = nn.ModuleList([_Weight((og, cg) + ks) for _ in range(groups)])
        self.spline_conv = nn.ModuleList([_Weight((og, cg * (grid_size + spline_order)) + ks) for _ in range(groups)])
        self.layer_norm = self._make_norm(og, self.norm_kind, affine, groups)
        self.prelus = nn.ModuleList([nn.PReLU() for _ in range(groups)])
        self.knots = make_knots(grid_size, spline_order, grid_range)   # plain attribute, like the reference
        for m in list(self.base_conv) + list(self.spline_conv):
            nn.init.kaiming_uniform_(m.weight, nonlinearity="linear")

    def forward(self, x):
        outs = []
        for g, xg in enumerate(torch.split(x, self.cg, dim=1)):
            nm = self.layer_norm[g]
            nw, nb = self._nw(nm)
            running = (nm.running_mean, nm.running_var) if isinstance(nm, nn.modules.batchnorm._BatchNorm) else None
            outs.append(kan_conv2d(xg, self.base_conv[g].weight, self.spline_conv[g].weight,
                                   self.prelus[g].weight, self.knots, self.spline_order, self.act,
                                   self.stride, self.padding, self.dilation, self.norm_kind,
                                   1e-5, nw, nb, running, self.training))
        return torch.cat(outs, dim=1)


class OracleChebyKANConv2D(_OracleBase):
    def __init__(self, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                 norm_layer=nn.InstanceNorm2d, affine=False):
        super().__init__()
        self._check_groups(input_dim, output_dim, groups)
        ks = self._kshape(kernel_size)
        cg, og = input_dim // groups, output_dim // groups
        self.groups, self.cg, self.og, self.degree = groups, cg, og, degree
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.norm_kind = _norm_kind(norm_layer)
        self.layer_norm = self._make_norm(og, self.norm_kind, affine, groups)
        self.poly_conv = nn.ModuleList([_Weight((og, (degree + 1) * cg) + ks) for _ in range(groups)])
        self.register_buffer("arange", torch.arange(0, degree + 1, 1).view(1, 1, -1, *([1] * self._nd)))
        for m in self.poly_conv:
            nn.init.normal_(m.weight, mean=0.0, std=1 / (input_dim * (degree + 1) * math.prod(ks)))   # overwritten next line
            nn.init.kaiming_normal_(m.weight, mode="fan_in", nonlinearity="relu")

    def forward(self, x):
        outs = []
        for g, xg in enumerate(torch.split(x, self.cg, dim=1)):
            nw, nb = self._nw(self.layer_norm[g])
            outs.append(cheby_conv2d(xg, self.poly_conv[g].weight, self.degree, self.stride, self.padding,
                                     self.dilation, self.norm_kind, 1e-5, nw, nb))
        return torch.cat(outs, dim=1)


class OracleGRAMKANConv2D(_OracleBase):
    def __init__(self, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                 norm_layer=nn.InstanceNorm2d, affine=False):
        super().__init__()
        self._check_groups(input_dim, output_dim, groups)
        ks = self._kshape(kernel_size)
        cg, og = input_dim // groups, output_dim // groups
        self.groups, self.cg, self.og, self.degree = groups, cg, og, degree
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.norm_kind = _norm_kind(norm_layer)
        self.base_conv = nn.ModuleList([_Weight((og, cg) + ks) for _ in range(groups)])
        self.layer_norm = self._make_norm(og, self.norm_kind, affine, groups)
        self.poly_weights = nn.Parameter(torch.randn(groups, og, cg * (degree + 1), *ks))
        self.beta_weights = nn.Parameter(torch.zeros(degree + 1))
        for m in self.base_conv:
            nn.init.kaiming_uniform_(m.weight, nonlinearity="linear")
        nn.init.kaiming_uniform_(self.poly_weights, nonlinearity="linear")
        nn.init.normal_(self.beta_weights, 0.0, 1.0 / (math.prod(ks) * input_dim * (degree + 1.0)))

    def forward(self, x, tanh_scale=None):
        outs = []
        scales = [None] * self.groups if tanh_scale is None else torch.split(tanh_scale, self.cg, dim=1)
        for g, xg in enumerate(torch.split(x, self.cg, dim=1)):
            nw, nb = self._nw(self.layer_norm[g])
            outs.append(gram_conv2d(xg, self.base_conv[g].weight, self.poly_weights[g], self.beta_weights,
                                    self.degree, self.stride, self.padding, self.dilation, self.norm_kind,
                                    1e-5, nw, nb, scales[g]))
        return torch.cat(outs, dim=1)


class _RBF(nn.Module):
    def __init__(self, lo, hi, n):
        super().__init__()
        self.grid = nn.Parameter(torch.linspace(lo, hi, n), requires_grad=False)
        self.denominator = (hi - lo) / (n - 1)


class OracleFastKANConv2D(_OracleBase):
    def __init__(self, input_dim, output_dim, kernel_size, groups=1, padding=0, stride=1, dilation=1,
                 grid_size=8, base_activation="silu", grid_range=(-2, 2), norm_layer=nn.InstanceNorm2d,
                 affine=False):
        super().__init__()
        self._check_groups(input_dim, output_dim, groups)
        ks = self._kshape(kernel_size)
        cg, og = input_dim // groups, output_dim // groups
        self.groups, self.cg, self.og = groups, cg, og
        self.act = _act_name(base_activation)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.norm_kind = _norm_kind(norm_layer)
        self.base_conv = nn.ModuleList([_Weight((og, cg) + ks) for _ in range(groups)])
        self.spline_conv = nn.ModuleList([_Weight((og, grid_size * cg) + ks) for _ in range(groups)])
        self.layer_norm = self._make_norm(cg, self.norm_kind, affine, groups)   # sized by INPUT channels
        self.rbf = _RBF(float(grid_range[0]), float(grid_range[1]), grid_size)
        for m in list(self.base_conv) + list(self.spline_conv):
            nn.init.kaiming_uniform_(m.weight, nonlinearity="linear")

    def forward(self, x):
        outs = []
        for g, xg in enumerate(torch.split(x, self.cg, dim=1)):
            nm = self.layer_norm[g]
            nw, nb = self._nw(nm)
            running = None
            if isinstance(nm, nn.modules.batchnorm._BatchNorm):
                running = (nm.running_mean, nm.running_var)
            outs.append(fastkan_conv2d(xg, self.base_conv[g].weight, self.spline_conv[g].weight, self.rbf.grid,
                                       self.rbf.denominator, self.act, self.stride, self.padding, self.dilation,
                                       self.norm_kind, 1e-5, nw, nb, running, self.training))
        return torch.cat(outs, dim=1)


class OracleKANConv3D(OracleKANConv2D):
    """KANConv3DLayer (kan_layers.py:261-271): the N-D layer body over nn.Conv3d / nn.InstanceNorm3d; x [N, C, D, H, W]."""
    _nd = 3


class OracleChebyKANConv3D(OracleChebyKANConv2D):
    """ChebyKANConv3DLayer (cheby_kan_layers.py:114-121)."""
    _nd = 3


class OracleGRAMKANConv3D(OracleGRAMKANConv2D):
    """GRAMKANConv3DLayer (gram_kan_layers.py:202-209)."""
    _nd = 3


class OracleFastKANConv3D(OracleFastKANConv2D):
    """FastKANConv3DLayer (fast_kan_layers.py:123-134)."""
    _nd = 3


class OracleKANConv1D(_OracleBase):
    """KANConv1DLayer (kan_layers.py:287-297): the N-D layer bound to nn.Conv1d / nn.InstanceNorm1d."""

    def __init__(self, input_dim, output_dim, kernel_size, spline_order=3, groups=1, padding=0, stride=1, dilation=1,
                 grid_size=5, base_activation="gelu", grid_range=(-1, 1), affine=False):
        super().__init__()
        self._check_groups(input_dim, output_dim, groups)
        k = int(kernel_size)
        cg, og = input_dim // groups, output_dim // groups
        self.groups, self.cg, self.og = groups, cg, og
        self.spline_order, self.act = spline_order, _act_name(base_activation)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.base_conv = nn.ModuleList([_Weight((og, cg, k)) for _ in range(groups)])
        self.spline_conv = nn.ModuleList([_Weight((og, cg * (grid_size + spline_order), k)) for _ in range(groups)])
        self.layer_norm = nn.ModuleList([nn.InstanceNorm1d(og, affine=affine) for _ in range(groups)])
        self.prelus = nn.ModuleList([nn.PReLU() for _ in range(groups)])
        self.knots = make_knots(grid_size, spline_order, grid_range)
        for m in list(self.base_conv) + list(self.spline_conv):
            nn.init.kaiming_uniform_(m.weight, nonlinearity="linear")

    def forward(self, x):
        outs = []
        for g, xg in enumerate(torch.split(x, self.cg, dim=1)):
            nw, nb = self._nw(self.layer_norm[g])
            outs.append(kan_conv1d(xg, self.base_conv[g].weight, self.spline_conv[g].weight, self.prelus[g].weight,
                                   self.knots, self.spline_order, self.act, self.stride, self.padding, self.dilation,
                                   1e-5, nw, nb))
        return torch.cat(outs, dim=1)


class OraclePolyKANConv(_OracleBase):
    """Template-A family layer (``family`` in hermite / gegenbauer / laguerre / lucas / fibonacci / bessel / taylor), any rank:
    module tree base_conv / poly_conv / layer_norm / prelus like <Family>KANConvNDLayer.__init__ (hermite_kan_layers.py:30-125)."""

    def __init__(self, family, nd, input_dim, output_dim, kernel_size, degree, groups=1, padding=0, stride=1, dilation=1,
                 base_activation="gelu", norm_layer=nn.InstanceNorm2d, affine=False, **fam):
        super().__init__()
        self._nd = nd
        self._check_groups(input_dim, output_dim, groups)
        ks = self._kshape(kernel_size) if nd != 1 else (int(kernel_size),)
        cg, og = input_dim // groups, output_dim // groups
        nb = degree if family == "taylor" else degree + 1
        self.family, self.fam, self.degree = family, fam, degree
        self.groups, self.cg, self.og, self.act = groups, cg, og, _act_name(base_activation)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.norm_kind = _norm_kind(norm_layer)
        self.base_conv = nn.ModuleList([_Weight((og, cg) + ks) for _ in range(groups)])
        self.poly_conv = nn.ModuleList([_Weight((og, cg * nb) + ks) for _ in range(groups)])
        self.layer_norm = self._make_norm(og, self.norm_kind, affine, groups)
        self.prelus = nn.ModuleList([nn.PReLU() for _ in range(groups)])
        for m in list(self.base_conv) + list(self.poly_conv):
            nn.init.kaiming_uniform_(m.weight, nonlinearity="linear")

    def forward(self, x):
        outs = []
        for g, xg in enumerate(torch.split(x, self.cg, dim=1)):
            nm = self.layer_norm[g]
            nw, nb = self._nw(nm)
            running = (nm.running_mean, nm.running_var) if isinstance(nm, nn.modules.batchnorm._BatchNorm) else None
            outs.append(polykan_conv(self.family, xg, self.base_conv[g].weight, self.poly_conv[g].weight,
                                     self.prelus[g].weight, self.degree, self.act, self.stride, self.padding, self.dilation,
                                     self.norm_kind, 1e-5, nw, nb, running, self.training, **self.fam))
        return torch.cat(outs, dim=1)


class OracleDegreeMajorPolyKANConv(_OracleBase):
    """LegendreKANConvNDLayer (legendre_kan_layers.py:50-161) / JacobiKANConvNDLayer (jacobi_kan_layers.py:56-178)."""

    def __init__(self, family, nd, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                 base_activation="gelu", norm_layer=nn.InstanceNorm2d, affine=False, **fam):
        super().__init__()
        self._nd = nd
        self._check_groups(input_dim, output_dim, groups)
        ks = self._kshape(kernel_size) if nd != 1 else (int(kernel_size),)
        cg, og = input_dim // groups, output_dim // groups
        if family == "jacobi":
            fam = dict({"a": 1.0, "b": 1.0}, **fam)           # JacobiKANConv2DLayer defaults (jacobi_kan_layers.py:191)
        self.family, self.fam, self.degree = family, fam, degree
        self.groups, self.cg, self.og = groups, cg, og
        self.out_act = "silu" if family == "legendre" else _act_name(base_activation)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.norm_kind = _norm_kind(norm_layer)
        self.base_conv = nn.ModuleList([_Weight((og, cg) + ks) for _ in range(groups)])
        self.layer_norm = self._make_norm(og, self.norm_kind, affine, groups)
        self.poly_weights = nn.Parameter(torch.randn(groups, og, cg * (degree + 1), *ks))
        for m in self.base_conv:
            nn.init.kaiming_uniform_(m.weight, nonlinearity="linear")
        if family == "legendre":
            nn.init.kaiming_uniform_(self.poly_weights, nonlinearity="linear")
        else:
            nn.init.normal_(self.poly_weights, mean=0.0, std=1 / (input_dim * (degree + 1) * math.prod(ks)))

    def forward(self, x):
        outs = []
        for g, xg in enumerate(torch.split(x, self.cg, dim=1)):
            nw, nb = self._nw(self.layer_norm[g])
            outs.append(degree_major_polykan_conv(self.family, xg, self.base_conv[g].weight, self.poly_weights[g], self.degree,
                                                  self.out_act, self.stride, self.padding, self.dilation, self.norm_kind,
                                                  1e-5, nw, nb, **self.fam))
        return torch.cat(outs, dim=1)


class OracleKANLayer(nn.Module):
    """KANLayer (kan_layers.py:8-114); state_dict keys base_weight, spline_weight, layer_norm.{weight,bias}, prelu.weight."""

    def __init__(self, input_features, output_features, grid_size=5, spline_order=3, base_activation="gelu",
                 grid_range=(-1, 1)):
        super().__init__()
        self.spline_order, self.act = spline_order, _act_name(base_activation)
        self.base_weight = nn.Parameter(torch.randn(output_features, input_features))
        self.spline_weight = nn.Parameter(torch.randn(output_features, input_features, grid_size + spline_order))
        self.layer_norm = nn.LayerNorm(output_features)
        self.prelu = nn.PReLU()
        self.knots = make_knots(grid_size, spline_order, grid_range)
        nn.init.kaiming_uniform_(self.base_weight, nonlinearity="linear")
        nn.init.kaiming_uniform_(self.spline_weight, nonlinearity="linear")

    def forward(self, x):
        return kan_linear(x, self.base_weight, self.spline_weight, self.layer_norm.weight, self.layer_norm.bias,
                          self.prelu.weight, self.knots, self.spline_order, self.act, self.layer_norm.eps)


class OracleKAN(nn.Module):
    """KAN MLP (models/kans.py:300-327): KANLayers with optional Dropout in between (keys layers.{i}.*)."""

    def __init__(self, layers_hidden, dropout=0.0, grid_size=5, spline_order=3, base_activation="gelu", grid_range=(-1, 1),
                 first_dropout=True):
        super().__init__()
        self.layers = nn.ModuleList([])
        if dropout > 0 and first_dropout:
            self.layers.append(nn.Dropout(p=dropout))
        n = len(layers_hidden) - 1
        for i, (fin, fout) in enumerate(zip(layers_hidden[:-1], layers_hidden[1:])):
            self.layers.append(OracleKANLayer(fin, fout, grid_size, spline_order, base_activation, grid_range))
            if dropout > 0 and i != n - 1:
                self.layers.append(nn.Dropout(p=dropout))

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


# ----------------------------------------------------------------------------------------------
# KAN-VGG (models/kan_vgg.py:29-188, 'kanconv' + Linear head) for the CPU baseline
# ----------------------------------------------------------------------------------------------
VGG_CFGS: Dict[str, List[Union[str, int]]] = {
    # reference cfgs (models/kan_vgg.py:20-26) ...
    "VGG16_small": [16, 16, "M", 32, 32, "M", 64, 64, 64, "M", 128, 128, 128, "M", 128, 128, 128],
    "VGG16_kansmall": [8, 8, "M", 16, 16, "M", 32, 32, 32, "M", 64, 64, 64, "M", 64, 64, 64],
    "VGG19_small": [16, 16, "M", 32, 32, "M", 64, 64, 64, 64, "M", 128, 128, 128, 128, "M", 128, 128, 128, 128],
    "VGG16": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512],
    "VGG19": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512],
    # ... plus the VGG11 cfg BASELINE config 3 needs (absent upstream)
    "VGG11": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512],
}


class OracleVGG(nn.Module):
    """KAN-VGG with B-spline KAN convs (k=3, pad=1, SiLU base activation, InstanceNorm) + Linear head."""

    def __init__(self, input_channels=3, num_classes=10, arch="VGG16", spline_order=3, grid_size=5,
                 grid_range=(-1, 1), expected_feature_shape=(1, 1), dropout_linear=0.5, width_scale=1):
        super().__init__()
        layers: List[nn.Module] = []
        c = input_channels
        for v in VGG_CFGS[arch]:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                oc = int(v) * width_scale
                layers.append(OracleKANConv2D(c, oc, 3, spline_order=spline_order, padding=1, grid_size=grid_size,
                                              base_activation="silu", grid_range=grid_range))
                c = oc
        self.features = nn.ModuleList(layers)
        self.avgpool = nn.AdaptiveAvgPool2d(expected_feature_shape)
        self.classifier = nn.Sequential(nn.Dropout(p=dropout_linear),
                                        nn.Linear(c * expected_feature_shape[0] * expected_feature_shape[1], num_classes))

    def forward(self, x):
        for m in self.features:
            x = m(x)
        return self.classifier(torch.flatten(self.avgpool(x), 1))


def conv_flops(n, cin, cout, ho, wo, kh, kw, width, groups=1) -> float:
    """SURVEY 8(d) shared formula: dense-equivalent FLOPs of ONE direction of one KAN conv layer."""
    return 2.0 * n * ho * wo * cout * (cin // groups) * width * kh * kw
