"""B-spline KAN convolution layers - drop-in for the reference's ``layers/kan_layers.py`` (KANConvNDLayer :116-258,
KANConv{1,2}DLayer :274-297).  Same constructor signatures, module tree and state_dict keys; the arithmetic runs in
hand-written CUDA (kanconv_b200.functional)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib as L
from .. import functional as KF
from ._base import KANConvBase, act_kind, check_groups, filter_norm_kwargs, make_dropout, pair


class KANConvNDLayer(KANConvBase):
    def __init__(self, conv_class, norm_class, input_dim, output_dim, spline_order, kernel_size,
                 groups=1, padding=0, stride=1, dilation=1,
                 ndim: int = 2, grid_size=5, base_activation=nn.GELU, grid_range=[-1, 1], dropout=0.0,
                 **norm_kwargs):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.spline_order, self.kernel_size = spline_order, kernel_size
        self.padding, self.stride, self.dilation = padding, stride, dilation
        self.groups, self.ndim, self.grid_size = groups, ndim, grid_size
        self.base_activation = base_activation() if base_activation is not None else nn.Identity()
        self.grid_range = grid_range
        self.norm_kwargs = norm_kwargs
        self.dropout = make_dropout(ndim, dropout)
        check_groups(groups, input_dim, output_dim)
        self.input_dim_group = input_dim // groups
        self.output_dim_group = output_dim // groups
        nb = grid_size + spline_order
        # parameter holders: the very nn.Conv / norm / PReLU modules of the reference, so that construction consumes
        # the RNG identically and state_dict keys match.  Their forward() is never called.
        self.base_conv = nn.ModuleList([conv_class(self.input_dim_group, self.output_dim_group, kernel_size, stride,
                                                   padding, dilation, groups=1, bias=False) for _ in range(groups)])
        self.spline_conv = nn.ModuleList([conv_class(nb * self.input_dim_group, self.output_dim_group, kernel_size,
                                                     stride, padding, dilation, groups=1, bias=False)
                                          for _ in range(groups)])
        self.layer_norm = nn.ModuleList([norm_class(self.output_dim_group, **filter_norm_kwargs(norm_class, norm_kwargs))
                                         for _ in range(groups)])
        self.prelus = nn.ModuleList([nn.PReLU() for _ in range(groups)])
        h = (grid_range[1] - grid_range[0]) / grid_size
        self.grid = torch.linspace(grid_range[0] - h * spline_order, grid_range[1] + h * spline_order,
                                   grid_size + 2 * spline_order + 1, dtype=torch.float32)   # plain attribute, as upstream
        for m in self.base_conv:
            nn.init.kaiming_uniform_(m.weight, nonlinearity='linear')
        for m in self.spline_conv:
            nn.init.kaiming_uniform_(m.weight, nonlinearity='linear')
        self._spec = KF.ConvSpec(basis=L.BASIS_BSPLINE, act=act_kind(self.base_activation), nb=nb, order=spline_order,
                                 params=tuple(float(v) for v in self.grid.tolist()),
                                 kernel=pair(kernel_size, ndim), stride=pair(stride, ndim), padding=pair(padding, ndim, fill=0),
                                 dilation=pair(dilation, ndim), groups=groups)

    def forward(self, x):
        if self.ndim == 3:
            z = self._kan_conv3d(self._spec, x, None, None, [m.weight for m in self.base_conv], [m.weight for m in self.spline_conv])
            y = self._norm_act3d(z, self.layer_norm, L.OUT_PRELU, [m.weight for m in self.prelus])
            return y if self.dropout is None else self.dropout(y)
        x4 = self._to4d(x)
        y = self._conv_norm_act(self._spec, x4, None, [self._w4d(m.weight) for m in self.base_conv],
                                [self._w4d(m.weight) for m in self.spline_conv], self.layer_norm, L.OUT_PRELU,
                                [m.weight for m in self.prelus])
        y = self._from4d(y)
        if self.dropout is not None:
            y = self.dropout(y)
        return y


class KANConv3DLayer(KANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, spline_order=3, groups=1, padding=0, stride=1, dilation=1,
                 grid_size=5, base_activation=nn.GELU, grid_range=[-1, 1], dropout=0.0, norm_layer=nn.InstanceNorm3d,
                 **norm_kwargs):
        super().__init__(nn.Conv3d, norm_layer, input_dim, output_dim, spline_order, kernel_size, groups=groups,
                         padding=padding, stride=stride, dilation=dilation, ndim=3, grid_size=grid_size,
                         base_activation=base_activation, grid_range=grid_range, dropout=dropout, **norm_kwargs)


class KANConv2DLayer(KANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, spline_order=3, groups=1, padding=0, stride=1, dilation=1,
                 grid_size=5, base_activation=nn.GELU, grid_range=[-1, 1], dropout=0.0, norm_layer=nn.InstanceNorm2d,
                 **norm_kwargs):
        super().__init__(nn.Conv2d, norm_layer, input_dim, output_dim, spline_order, kernel_size, groups=groups,
                         padding=padding, stride=stride, dilation=dilation, ndim=2, grid_size=grid_size,
                         base_activation=base_activation, grid_range=grid_range, dropout=dropout, **norm_kwargs)


class KANConv1DLayer(KANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, spline_order=3, groups=1, padding=0, stride=1, dilation=1,
                 grid_size=5, base_activation=nn.GELU, grid_range=[-1, 1], dropout=0.0, norm_layer=nn.InstanceNorm1d,
                 **norm_kwargs):
        super().__init__(nn.Conv1d, norm_layer, input_dim, output_dim, spline_order, kernel_size, groups=groups,
                         padding=padding, stride=stride, dilation=dilation, ndim=1, grid_size=grid_size,
                         base_activation=base_activation, grid_range=grid_range, dropout=dropout, **norm_kwargs)


class KANLayer(nn.Module):
    """B-spline KAN fully-connected layer - drop-in for the reference's ``KANLayer`` (kan_layers.py:8-114).

    SURVEY 8(f) rank 2: the basis expansion + contraction ``[B, in*nb] x [in*nb, out]`` (+ the base branch) runs through the
    same CUDA op as the convolution (a 1x1 convolution over a 1x1 map, ``spline_weight [out, in, nb]`` viewed as
    ``[out, in*nb, 1, 1]``), and the LayerNorm + PReLU tail (kan_layers.py:110-112) through ``kc_layernorm_act_fwd/bwd``."""

    def __init__(self, input_features, output_features, grid_size=5, spline_order=3, base_activation=nn.GELU,
                 grid_range=[-1, 1]):
        super().__init__()
        self.input_features, self.output_features = input_features, output_features
        self.grid_size, self.spline_order, self.grid_range = grid_size, spline_order, grid_range
        self.base_activation = base_activation() if base_activation is not None else nn.Identity()
        self.base_weight = nn.Parameter(torch.randn(output_features, input_features))
        self.spline_weight = nn.Parameter(torch.randn(output_features, input_features, grid_size + spline_order))
        self.layer_norm = nn.LayerNorm(output_features)
        self.prelu = nn.PReLU()
        h = (grid_range[1] - grid_range[0]) / grid_size
        knots = torch.linspace(grid_range[0] - h * spline_order, grid_range[1] + h * spline_order,
                               grid_size + 2 * spline_order + 1, dtype=torch.float32)
        self.grid = knots.expand(input_features, -1).contiguous()
        nn.init.kaiming_uniform_(self.base_weight, nonlinearity='linear')
        nn.init.kaiming_uniform_(self.spline_weight, nonlinearity='linear')
        self._spec = KF.ConvSpec(basis=L.BASIS_BSPLINE, act=act_kind(self.base_activation), nb=grid_size + spline_order,
                                 order=spline_order, params=tuple(float(v) for v in knots.tolist()), kernel=(1, 1),
                                 stride=(1, 1), padding=(0, 0), dilation=(1, 1), groups=1)
        self.precision = None

    def forward(self, x):
        lead = x.shape[:-1]
        x4 = x.reshape(-1, self.input_features, 1, 1)
        z = KF.kan_conv(self._spec, x4, None, None, [self.base_weight[:, :, None, None]],
                        [self.spline_weight.reshape(self.output_features, -1, 1, 1)], self.precision)
        z = z.reshape(*lead, self.output_features)
        return KF.layer_norm_act(z, self.layer_norm.weight, self.layer_norm.bias, self.prelu.weight, self.layer_norm.eps)
