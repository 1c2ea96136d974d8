"""Polynomial KAN convolution layers whose basis obeys a three-term recurrence - drop-ins for nine of the reference's
"template" families.  One CUDA basis functor (``KC_BASIS_RECUR`` in ``include/kanconv.h``) evaluates

    p_0 = c0,   p_1 = a1 t + b1,   p_i = (A_i t + B_i) p_{i-1} + C_i p_{i-2},        t = tanh(x)

and its derivative inside the same fused kernels as the B-spline / Chebyshev layers; a family is a coefficient table.

Template A - ``base_conv`` / ``poly_conv`` / ``layer_norm`` / ``prelus``, expanded channel ``c*(D+1) + j``,
``PReLU(norm(conv(act(x)) + conv(P(tanh x))))``, Dropout on the output:
  HermiteKANConvNDLayer     layers/hermite_kan_layers.py:30-165      H_1 = 2t, H_n = 2t H_{n-1} - 2(n-1) H_{n-2}
  GegenbauerKANConvNDLayer  layers/gegenbauer_kan_layers.py:34-183   (n+1) C_{n+1} = 2(n+a) t C_n - (n+2a-1) C_{n-1}
  LaguerreKANConvNDLayer    layers/laguerre_kan_layers.py:38-184     k L_k = (2k-1+a-t) L_{k-1} - (k-1+a) L_{k-2}
  LucasKANConvNDLayer       layers/lucas_kan_layers.py:40-200        L_0 = 2, L_n = t L_{n-1} + L_{n-2}
  FibonacciKANConvNDLayer   layers/fibonacci_kan_layers.py:41-203    F_0 = 0, F_1 = 1, F_n = t F_{n-1} + F_{n-2}
  BesselKANConvNDLayer      layers/bessel_kan_layers.py:38-172       y_1 = t+1, y_n = (2n-1) t y_{n-1} + y_{n-2}
  TaylorKANConvNDLayer      layers/taylor_kan_layers.py:40-177       t^0 .. t^(degree-1)  (``degree`` terms)
Template B - ``base_conv`` / ``layer_norm`` / ``poly_weights``, degree-major expanded channel ``j*C + c``
(``torch.concatenate(polys, dim=1)``), base branch WITHOUT activation, ``act(norm(...))`` on the output:
  LegendreKANConvNDLayer    layers/legendre_kan_layers.py:50-161     on x min-max normalised over the whole group tensor
  JacobiKANConvNDLayer      layers/jacobi_kan_layers.py:56-178       on tanh(x), parameters a, b

Same constructor signatures, module trees, state_dict keys and RNG consumption as upstream."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib as L
from .. import functional as KF
from ._base import KANConvBase, act_kind, check_groups, filter_norm_kwargs, make_dropout, pair

Coef = Tuple[float, float, float, List[Tuple[float, float, float]]]      # c0, a1, b1, [(A_i, B_i, C_i) for i = 2 ..]


# ---- coefficient tables -------------------------------------------------------------------------------------------------------
def hermite_coef(nb: int) -> Coef:
    return 1.0, 2.0, 0.0, [(2.0, 0.0, -2.0 * (i - 1)) for i in range(2, nb)]


def gegenbauer_coef(nb: int, alpha: float) -> Coef:
    # i = n + 1:  C_i = (2 (i - 1 + alpha) t C_{i-1} - (i + 2 alpha - 2) C_{i-2}) / i
    return 1.0, 2.0 * alpha, 0.0, [(2.0 * (i - 1 + alpha) / i, 0.0, -(i + 2.0 * alpha - 2.0) / i) for i in range(2, nb)]


def laguerre_coef(nb: int, alpha: float) -> Coef:
    return 1.0, -1.0, 1.0 + alpha, [(-1.0 / k, (2.0 * k - 1.0 + alpha) / k, -(k - 1.0 + alpha) / k) for k in range(2, nb)]


def lucas_coef(nb: int) -> Coef:
    return 2.0, 1.0, 0.0, [(1.0, 0.0, 1.0)] * max(nb - 2, 0)


def fibonacci_coef(nb: int) -> Coef:
    return 0.0, 0.0, 1.0, [(1.0, 0.0, 1.0)] * max(nb - 2, 0)


def bessel_coef(nb: int) -> Coef:
    return 1.0, 1.0, 1.0, [(2.0 * i - 1.0, 0.0, 1.0) for i in range(2, nb)]


def taylor_coef(nb: int) -> Coef:
    return 1.0, 1.0, 0.0, [(1.0, 0.0, 0.0)] * max(nb - 2, 0)


def legendre_coef(nb: int) -> Coef:
    # i = n + 1:  P_i = ((2 i - 1) t P_{i-1} - (i - 1) P_{i-2}) / i
    return 1.0, 1.0, 0.0, [((2.0 * i - 1.0) / i, 0.0, -(i - 1.0) / i) for i in range(2, nb)]


def jacobi_coef(nb: int, a: float, b: float) -> Coef:
    rows = []
    for i in range(2, nb):
        th_k = (2 * i + a + b) * (2 * i + a + b - 1) / (2 * i * (i + a + b))
        th_k1 = (2 * i + a + b - 1) * (a * a - b * b) / (2 * i * (i + a + b) * (2 * i + a + b - 2))
        th_k2 = (i + a - 1) * (i + b - 1) * (2 * i + a + b) / (i * (i + a + b) * (2 * i + a + b - 2))
        rows.append((th_k, th_k1, -th_k2))
    return 1.0, (a + b + 2.0) / 2.0, (a - b) / 2.0, rows


def recur_params(coef: Coef, nb: int, pre_squashed: bool) -> Tuple[float, ...]:
    """kc_desc.params of KC_BASIS_RECUR*: (pre, c0, a1, b1, A_2, B_2, C_2, ...), 4 + 3 * max(nb - 2, 0) floats."""
    c0, a1, b1, rows = coef
    rows = list(rows)[:max(nb - 2, 0)]
    out = [1.0 if pre_squashed else 0.0, float(c0), float(a1), float(b1)]
    for r in rows:
        out += [float(v) for v in r]
    if len(out) > L.KC_MAX_PARAMS:
        raise NotImplementedError(f"polynomial degree too high for the CUDA kernels ({len(out)} > {L.KC_MAX_PARAMS} coefficients)")
    return tuple(out)


# ---- template A ---------------------------------------------------------------------------------------------------------------
class RecurrenceKANConvNDLayer(KANConvBase):
    """Shared body of the seven template-A families; subclasses provide ``_coef(nb)`` and ``_min_degree``."""
    _min_degree = 0
    _degree_msg = 'degree must be non-negative'

    def _coef(self, nb: int) -> Coef:
        raise NotImplementedError

    def _nb(self) -> int:
        return self.degree + 1

    def _init_template(self, conv_class, norm_class, input_dim, output_dim, kernel_size, degree, groups, padding, stride,
                       dilation, ndim, base_activation, dropout, norm_kwargs):
        check_groups(groups, input_dim, output_dim)
        if degree < self._min_degree:
            raise ValueError(self._degree_msg)
        self.input_dim, self.output_dim, self.kernel_size, self.degree = input_dim, output_dim, kernel_size, degree
        self.groups, self.padding, self.stride, self.dilation, self.ndim = groups, padding, stride, dilation, ndim
        self.base_activation = base_activation() if base_activation is not None else nn.Identity()
        self.norm_kwargs = norm_kwargs
        self.input_dim_group, self.output_dim_group = input_dim // groups, output_dim // groups
        nb = self._nb()
        self.poly_input_dim_group = self.input_dim_group * nb
        # parameter holders: the reference's own module tree (RNG consumption and state_dict keys match); never called
        self.base_conv = nn.ModuleList([conv_class(self.input_dim_group, self.output_dim_group, kernel_size, stride, padding,
                                                   dilation, groups=1, bias=False) for _ in range(groups)])
        self.poly_conv = nn.ModuleList([conv_class(self.poly_input_dim_group, self.output_dim_group, kernel_size, stride,
                                                   padding, dilation, groups=1, bias=False) for _ in range(groups)])
        self.layer_norm = nn.ModuleList([norm_class(self.output_dim_group, **filter_norm_kwargs(norm_class, norm_kwargs))
                                         for _ in range(groups)])
        self.prelus = nn.ModuleList([nn.PReLU() for _ in range(groups)])
        self.dropout = make_dropout(ndim, dropout)
        for m in list(self.base_conv) + list(self.poly_conv):
            nn.init.kaiming_uniform_(m.weight, nonlinearity='linear')
        self._spec = KF.ConvSpec(basis=L.BASIS_RECUR, act=act_kind(self.base_activation), nb=nb, order=nb - 1,
                                 params=recur_params(self._coef(nb), nb, False),
                                 kernel=pair(kernel_size, ndim), stride=pair(stride, ndim), padding=pair(padding, ndim, fill=0),
                                 dilation=pair(dilation, ndim), groups=groups)

    def __init__(self, conv_class, norm_class, input_dim, output_dim, kernel_size, degree, groups=1, padding=0, stride=1,
                 dilation=1, ndim: int = 2, base_activation=nn.GELU, dropout: float = 0.0, **norm_kwargs):
        super().__init__()
        self._init_template(conv_class, norm_class, input_dim, output_dim, kernel_size, degree, groups, padding, stride,
                            dilation, ndim, base_activation, dropout, norm_kwargs)

    def forward(self, x):
        w_base, w_poly = [m.weight for m in self.base_conv], [m.weight for m in self.poly_conv]
        alphas = [m.weight for m in self.prelus]
        if self.ndim == 3:
            z = self._kan_conv3d(self._spec, x, None, None, w_base, w_poly)
            y = self._norm_act3d(z, self.layer_norm, L.OUT_PRELU, alphas)
        else:
            y = self._from4d(self._conv_norm_act(self._spec, self._to4d(x), None, [self._w4d(w) for w in w_base],
                                                 [self._w4d(w) for w in w_poly], self.layer_norm, L.OUT_PRELU, alphas))
        return y if self.dropout is None else self.dropout(y)


_CONVS = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
_INORMS = {1: nn.InstanceNorm1d, 2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}


def _bind(base, ndim: int, name: str):
    """The reference's ``<Family>KANConv{1,2,3}DLayer`` classes: the N-D layer bound to nn.Conv{n}d / nn.InstanceNorm{n}d."""
    def __init__(self, input_dim, output_dim, kernel_size, degree, groups=1, padding=0, stride=1, dilation=1,
                 base_activation=nn.GELU, dropout=0.0, norm_layer=_INORMS[ndim], **norm_kwargs):
        base.__init__(self, conv_class=_CONVS[ndim], norm_class=norm_layer, input_dim=input_dim, output_dim=output_dim,
                      kernel_size=kernel_size, degree=degree, groups=groups, padding=padding, stride=stride,
                      dilation=dilation, ndim=ndim, base_activation=base_activation, dropout=dropout, **norm_kwargs)
    return type(name, (base,), {"__init__": __init__, "__module__": base.__module__,
                                "__doc__": f"{base.__name__} bound to nn.Conv{ndim}d / nn.InstanceNorm{ndim}d."})


def _bind_gegenbauer(base, ndim: int, name: str):
    """Same, with Gegenbauer's positional ``alpha_param`` after ``degree`` (gegenbauer_kan_layers.py:185-246)."""
    def __init__(self, input_dim, output_dim, kernel_size, degree, alpha_param, groups=1, padding=0, stride=1, dilation=1,
                 base_activation=nn.GELU, dropout=0.0, norm_layer=_INORMS[ndim], **norm_kwargs):
        base.__init__(self, conv_class=_CONVS[ndim], norm_class=norm_layer, input_dim=input_dim, output_dim=output_dim,
                      kernel_size=kernel_size, degree=degree, alpha_param=alpha_param, groups=groups, padding=padding,
                      stride=stride, dilation=dilation, ndim=ndim, base_activation=base_activation, dropout=dropout,
                      **norm_kwargs)
    return type(name, (base,), {"__init__": __init__, "__module__": base.__module__,
                                "__doc__": f"{base.__name__} bound to nn.Conv{ndim}d / nn.InstanceNorm{ndim}d."})


def _bind_laguerre(base, ndim: int, name: str):
    """Same, with Laguerre's positional ``alpha`` after ``degree`` (laguerre_kan_layers.py:186-212)."""
    def __init__(self, input_dim, output_dim, kernel_size, degree, alpha, groups=1, padding=0, stride=1, dilation=1,
                 base_activation=nn.GELU, dropout=0.0, norm_layer=_INORMS[ndim], **norm_kwargs):
        base.__init__(self, conv_class=_CONVS[ndim], norm_class=norm_layer, input_dim=input_dim, output_dim=output_dim,
                      kernel_size=kernel_size, degree=degree, alpha=alpha, groups=groups, padding=padding, stride=stride,
                      dilation=dilation, ndim=ndim, base_activation=base_activation, dropout=dropout, **norm_kwargs)
    return type(name, (base,), {"__init__": __init__, "__module__": base.__module__,
                                "__doc__": f"{base.__name__} bound to nn.Conv{ndim}d / nn.InstanceNorm{ndim}d."})


class HermiteKANConvNDLayer(RecurrenceKANConvNDLayer):
    def _coef(self, nb):
        return hermite_coef(nb)


class LucasKANConvNDLayer(RecurrenceKANConvNDLayer):
    def _coef(self, nb):
        return lucas_coef(nb)


class FibonacciKANConvNDLayer(RecurrenceKANConvNDLayer):
    _min_degree = 1
    _degree_msg = 'degree must be at least 1'

    def _coef(self, nb):
        return fibonacci_coef(nb)


class BesselKANConvNDLayer(RecurrenceKANConvNDLayer):
    def _coef(self, nb):
        return bessel_coef(nb)


class TaylorKANConvNDLayer(RecurrenceKANConvNDLayer):
    """``degree`` = number of Taylor terms t^0 .. t^(degree-1) (taylor_kan_layers.py:47,81)."""
    _min_degree = 1
    _degree_msg = 'degree must be at least 1'

    def _nb(self):
        return self.degree

    def _coef(self, nb):
        return taylor_coef(nb)


class GegenbauerKANConvNDLayer(RecurrenceKANConvNDLayer):
    def __init__(self, conv_class, norm_class, input_dim, output_dim, kernel_size, degree, alpha_param, groups=1, padding=0,
                 stride=1, dilation=1, ndim: int = 2, base_activation=nn.GELU, dropout: float = 0.0, **norm_kwargs):
        KANConvBase.__init__(self)
        check_groups(groups, input_dim, output_dim)
        if degree < 0:
            raise ValueError('degree must be non-negative')
        if alpha_param <= -0.5:
            raise ValueError('alpha_param must be greater than -0.5')
        self.alpha_param = alpha_param
        self._init_template(conv_class, norm_class, input_dim, output_dim, kernel_size, degree, groups, padding, stride,
                            dilation, ndim, base_activation, dropout, norm_kwargs)

    def _coef(self, nb):
        return gegenbauer_coef(nb, float(self.alpha_param))


class LaguerreKANConvNDLayer(RecurrenceKANConvNDLayer):
    def __init__(self, conv_class, norm_class, input_dim, output_dim, kernel_size, degree, alpha, groups=1, padding=0,
                 stride=1, dilation=1, ndim: int = 2, base_activation=nn.GELU, dropout: float = 0.0, **norm_kwargs):
        KANConvBase.__init__(self)
        check_groups(groups, input_dim, output_dim)
        if degree < 0:
            raise ValueError('degree must be non-negative')
        if alpha <= -1.0:
            raise ValueError('alpha must be greater than -1 for Laguerre polynomials')
        self.alpha = alpha
        self._init_template(conv_class, norm_class, input_dim, output_dim, kernel_size, degree, groups, padding, stride,
                            dilation, ndim, base_activation, dropout, norm_kwargs)

    def _coef(self, nb):
        return laguerre_coef(nb, float(self.alpha))


HermiteKANConv1DLayer, HermiteKANConv2DLayer, HermiteKANConv3DLayer = (
    _bind(HermiteKANConvNDLayer, n, f"HermiteKANConv{n}DLayer") for n in (1, 2, 3))
LucasKANConv1DLayer, LucasKANConv2DLayer, LucasKANConv3DLayer = (
    _bind(LucasKANConvNDLayer, n, f"LucasKANConv{n}DLayer") for n in (1, 2, 3))
FibonacciKANConv1DLayer, FibonacciKANConv2DLayer, FibonacciKANConv3DLayer = (
    _bind(FibonacciKANConvNDLayer, n, f"FibonacciKANConv{n}DLayer") for n in (1, 2, 3))
BesselKANConv1DLayer, BesselKANConv2DLayer, BesselKANConv3DLayer = (
    _bind(BesselKANConvNDLayer, n, f"BesselKANConv{n}DLayer") for n in (1, 2, 3))
TaylorKANConv1DLayer, TaylorKANConv2DLayer, TaylorKANConv3DLayer = (
    _bind(TaylorKANConvNDLayer, n, f"TaylorKANConv{n}DLayer") for n in (1, 2, 3))
GegenbauerKANConv1DLayer, GegenbauerKANConv2DLayer, GegenbauerKANConv3DLayer = (
    _bind_gegenbauer(GegenbauerKANConvNDLayer, n, f"GegenbauerKANConv{n}DLayer") for n in (1, 2, 3))
LaguerreKANConv1DLayer, LaguerreKANConv2DLayer, LaguerreKANConv3DLayer = (
    _bind_laguerre(LaguerreKANConvNDLayer, n, f"LaguerreKANConv{n}DLayer") for n in (1, 2, 3))


# ---- template B ---------------------------------------------------------------------------------------------------------------
class _DegreeMajorPolyLayer(KANConvBase):
    """Shared body of Legendre / Jacobi: ``poly_weights [groups, Cout/g, (Cin/g)(D+1), k..]`` with inner index ``j*C + c``."""

    def _init_template(self, conv_class, norm_class, conv_w_fun, input_dim, output_dim, degree, kernel_size, base_activation,
                       groups, padding, stride, dilation, dropout, ndim, norm_kwargs):
        ndim = int(ndim)
        self.input_dim, self.output_dim, self.degree, self.kernel_size = input_dim, output_dim, degree, kernel_size
        self.padding, self.stride, self.dilation, self.groups = padding, stride, dilation, groups
        self.base_activation = base_activation
        self.conv_w_fun, self.ndim, self.norm_kwargs = conv_w_fun, ndim, norm_kwargs
        self.dropout = make_dropout(ndim, dropout)
        check_groups(groups, input_dim, output_dim)
        self.base_conv = nn.ModuleList([conv_class(input_dim // groups, output_dim // groups, kernel_size, stride, padding,
                                                   dilation, groups=1, bias=False) for _ in range(groups)])
        self.layer_norm = nn.ModuleList([norm_class(output_dim // groups, **filter_norm_kwargs(norm_class, norm_kwargs))
                                         for _ in range(groups)])
        poly_shape = (groups, output_dim // groups, (input_dim // groups) * (degree + 1)) + tuple(
            kernel_size for _ in range(ndim))
        self.poly_weights = nn.Parameter(torch.randn(*poly_shape))
        for m in self.base_conv:
            nn.init.kaiming_uniform_(m.weight, nonlinearity='linear')

    def _make_spec(self, coef: Coef, pre_squashed: bool):
        nb, ndim = self.degree + 1, self.ndim
        return KF.ConvSpec(basis=L.BASIS_RECUR_DM, act=L.ACT_IDENTITY, nb=nb, order=self.degree,
                           params=recur_params(coef, nb, pre_squashed), kernel=pair(self.kernel_size, ndim),
                           stride=pair(self.stride, ndim), padding=pair(self.padding, ndim, fill=0),
                           dilation=pair(self.dilation, ndim), groups=self.groups)

    def _out_act(self):
        """(kernel-side output activation, torch-side remainder) of ``base_activation(norm(z))``."""
        a = self.base_activation
        if isinstance(a, nn.SiLU):
            return L.OUT_SILU, None
        if isinstance(a, nn.Identity):
            return L.OUT_NONE, None
        if isinstance(a, nn.GELU) and a.approximate == 'none':
            return L.OUT_NONE, F.gelu          # elementwise epilogue outside the norm kernel (differentiated by autograd)
        raise NotImplementedError(f"{type(self).__name__}: output activation {type(a).__name__} is not implemented "
                                  "(supported: SiLU, GELU, Identity)")

    def _forward(self, x, x_basis):
        w_base = [m.weight for m in self.base_conv]
        w_poly = [self.poly_weights[g] for g in range(self.groups)]
        out_act, tail = self._out_act()
        if self.ndim == 3:
            z = self._kan_conv3d(self._spec, x, x_basis, None, w_base, w_poly)
            y = self._norm_act3d(z, self.layer_norm, out_act)
        elif x_basis is None:
            y = self._from4d(self._conv_norm_act(self._spec, self._to4d(x), None, [self._w4d(w) for w in w_base],
                                                 [self._w4d(w) for w in w_poly], self.layer_norm, out_act))
        else:
            z = KF.kan_conv(self._spec, self._to4d(x), self._to4d(x_basis), None, [self._w4d(w) for w in w_base],
                            [self._w4d(w) for w in w_poly], self.precision)
            y = self._from4d(self._norm_act(z, self.layer_norm, out_act))
        return y if tail is None else tail(y)


class LegendreKANConvNDLayer(_DegreeMajorPolyLayer):
    def __init__(self, conv_class, norm_class, conv_w_fun, input_dim, output_dim, degree, kernel_size,
                 groups=1, padding=0, stride=1, dilation=1, dropout: float = 0.0, ndim: int = 2, **norm_kwargs):
        super().__init__()
        self._init_template(conv_class, norm_class, conv_w_fun, input_dim, output_dim, degree, kernel_size, nn.SiLU(),
                            groups, padding, stride, dilation, dropout, ndim, norm_kwargs)
        nn.init.kaiming_uniform_(self.poly_weights, nonlinearity='linear')
        self._spec = self._make_spec(legendre_coef(degree + 1), True)

    def forward(self, x):
        # legendre_kan_layers.py:127-133: every group's slice is mapped to [-1, 1] with the min / max of the WHOLE slice (batch
        # included) - two full reductions and an elementwise map in torch (autograd sends the gradient through min / max like
        # upstream); Dropout acts on the normalised input.  The kernels take it as the basis input "pre-squashed".
        cg = self.input_dim // self.groups
        parts = []
        for xg in torch.split(x, cg, dim=1):
            xn = 2 * (xg - xg.min()) / (xg.max() - xg.min()) - 1 if xg.shape[0] > 0 else xg
            parts.append(xn if self.dropout is None else self.dropout(xn))
        return self._forward(x, parts[0] if len(parts) == 1 else torch.cat(parts, dim=1))


class JacobiKANConvNDLayer(_DegreeMajorPolyLayer):
    def __init__(self, conv_class, norm_class, conv_w_fun, input_dim, output_dim, degree, kernel_size,
                 base_activation=nn.SiLU, a: float = 1.0, b: float = 1.0,
                 groups=1, padding=0, stride=1, dilation=1, dropout: float = 0.0, ndim: int = 2, **norm_kwargs):
        super().__init__()
        self.a, self.b = a, b
        self._init_template(conv_class, norm_class, conv_w_fun, input_dim, output_dim, degree, kernel_size,
                            base_activation() if base_activation is not None else nn.Identity(),
                            groups, padding, stride, dilation, dropout, ndim, norm_kwargs)
        nn.init.normal_(self.poly_weights, mean=0.0, std=1 / (input_dim * (degree + 1) * kernel_size ** int(ndim)))
        self._spec = self._make_spec(jacobi_coef(degree + 1, float(a), float(b)), False)

    def forward(self, x):
        if self.dropout is not None and self.training:
            # jacobi_kan_layers.py:147-148 drops whole channels of the EXPANDED tensor, which never exists here
            raise NotImplementedError("JacobiKANConv: dropout > 0 in training mode acts on the expanded basis tensor and is "
                                      "not implemented in the fused kernels")
        return self._forward(x, None)


def _bind_b(base, ndim: int, name: str, jacobi: bool):
    conv_fun = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[ndim]
    if jacobi:
        def __init__(self, input_dim, output_dim, kernel_size, degree=3, base_activation=nn.GELU, a=1.0, b=1.0, groups=1,
                     padding=0, stride=1, dilation=1, dropout: float = 0.0, norm_layer=_INORMS[ndim], **norm_kwargs):
            base.__init__(self, conv_class=_CONVS[ndim], norm_class=norm_layer, conv_w_fun=conv_fun, input_dim=input_dim,
                          output_dim=output_dim, degree=degree, kernel_size=kernel_size, base_activation=base_activation,
                          a=a, b=b, groups=groups, padding=padding, stride=stride, dilation=dilation, ndim=ndim,
                          dropout=dropout, **norm_kwargs)
    else:
        def __init__(self, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                     dropout: float = 0.0, norm_layer=_INORMS[ndim], **norm_kwargs):
            base.__init__(self, _CONVS[ndim], norm_layer, conv_fun, input_dim, output_dim, degree, kernel_size,
                          groups=groups, padding=padding, stride=stride, dilation=dilation, ndim=ndim, dropout=dropout,
                          **norm_kwargs)
    return type(name, (base,), {"__init__": __init__, "__module__": base.__module__,
                                "__doc__": f"{base.__name__} bound to nn.Conv{ndim}d / nn.InstanceNorm{ndim}d."})


LegendreKANConv1DLayer, LegendreKANConv2DLayer, LegendreKANConv3DLayer = (
    _bind_b(LegendreKANConvNDLayer, n, f"LegendreKANConv{n}DLayer", False) for n in (1, 2, 3))
JacobiKANConv1DLayer, JacobiKANConv2DLayer, JacobiKANConv3DLayer = (
    _bind_b(JacobiKANConvNDLayer, n, f"JacobiKANConv{n}DLayer", True) for n in (1, 2, 3))
