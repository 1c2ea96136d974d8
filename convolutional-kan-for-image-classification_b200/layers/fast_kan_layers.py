"""FastKAN (Gaussian RBF) convolution layers - drop-in for the reference's ``layers/fast_kan_layers.py`` (:34-162)."""
from __future__ import annotations

import torch.nn as nn

from .. import _lib as L
from .. import functional as KF
from ..utils.utils import RadialBasisFunction
from ._base import KANConvBase, act_kind, check_groups, filter_norm_kwargs, make_dropout, pair


class FastKANConvNDLayer(KANConvBase):
    def __init__(self, conv_class, norm_class, input_dim, output_dim, kernel_size,
                 groups=1, padding=0, stride=1, dilation=1,
                 ndim: int = 2, grid_size=8, base_activation=nn.SiLU, grid_range=[-2, 2], dropout=0.0, **norm_kwargs):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.kernel_size, self.padding, self.stride, self.dilation = kernel_size, padding, stride, dilation
        self.groups, self.ndim, self.grid_size = groups, ndim, grid_size
        self.base_activation = base_activation() if base_activation is not None else nn.Identity()
        self.grid_range = grid_range
        self.norm_kwargs = norm_kwargs
        check_groups(groups, input_dim, output_dim)
        self.base_conv = nn.ModuleList([conv_class(input_dim // groups, output_dim // groups, kernel_size, stride, padding,
                                                   dilation, groups=1, bias=False) for _ in range(groups)])
        self.spline_conv = nn.ModuleList([conv_class(grid_size * input_dim // groups, output_dim // groups, kernel_size,
                                                     stride, padding, dilation, groups=1, bias=False)
                                          for _ in range(groups)])
        # NB: the FastKAN norm acts on the INPUT of the RBF branch and is sized by the input channels (:80)
        self.layer_norm = nn.ModuleList([norm_class(input_dim // groups, **filter_norm_kwargs(norm_class, norm_kwargs))
                                         for _ in range(groups)])
        self.rbf = RadialBasisFunction(grid_range[0], grid_range[1], grid_size)
        self.dropout = make_dropout(ndim, dropout)
        for m in self.base_conv:
            nn.init.kaiming_uniform_(m.weight, nonlinearity='linear')
        for m in self.spline_conv:
            nn.init.kaiming_uniform_(m.weight, nonlinearity='linear')
        self._act = act_kind(self.base_activation)
        self._geom = dict(kernel=pair(kernel_size, ndim), stride=pair(stride, ndim), padding=pair(padding, ndim, fill=0),
                          dilation=pair(dilation, ndim), groups=groups)

    def forward(self, x):
        if self.ndim == 3:
            xd = x if self.dropout is None else self.dropout(x)
            u = self._norm_act3d(xd, self.layer_norm, L.OUT_NONE)
            spec = KF.ConvSpec(basis=L.BASIS_RBF, act=self._act, nb=self.grid_size, order=0, params=self.rbf.host_params(),
                               **self._geom)
            return self._kan_conv3d(spec, x, u, None, [m.weight for m in self.base_conv], [m.weight for m in self.spline_conv])
        x4 = self._to4d(x)
        xd = x4 if self.dropout is None else self._to4d(self.dropout(x))
        u = self._norm_act(xd, self.layer_norm, L.OUT_NONE)
        spec = KF.ConvSpec(basis=L.BASIS_RBF, act=self._act, nb=self.grid_size, order=0, params=self.rbf.host_params(),
                           **self._geom)
        z = KF.kan_conv(spec, x4, u, None, [self._w4d(m.weight) for m in self.base_conv],
                        [self._w4d(m.weight) for m in self.spline_conv], self.precision)
        return self._from4d(z)


class FastKANConv3DLayer(FastKANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, groups=1, padding=0, stride=1, dilation=1,
                 grid_size=8, base_activation=nn.SiLU, grid_range=[-2, 2], dropout=0.0,
                 norm_layer=nn.InstanceNorm3d, **norm_kwargs):
        super().__init__(nn.Conv3d, norm_layer, input_dim, output_dim, kernel_size, groups=groups, padding=padding,
                         stride=stride, dilation=dilation, ndim=3, grid_size=grid_size, base_activation=base_activation,
                         grid_range=grid_range, dropout=dropout, **norm_kwargs)


class FastKANConv2DLayer(FastKANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, groups=1, padding=0, stride=1, dilation=1,
                 grid_size=8, base_activation=nn.SiLU, grid_range=[-2, 2], dropout=0.0,
                 norm_layer=nn.InstanceNorm2d, **norm_kwargs):
        super().__init__(nn.Conv2d, norm_layer, input_dim, output_dim, kernel_size, groups=groups, padding=padding,
                         stride=stride, dilation=dilation, ndim=2, grid_size=grid_size, base_activation=base_activation,
                         grid_range=grid_range, dropout=dropout, **norm_kwargs)


class FastKANConv1DLayer(FastKANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, groups=1, padding=0, stride=1, dilation=1,
                 grid_size=8, base_activation=nn.SiLU, grid_range=[-2, 2], dropout=0.0,
                 norm_layer=nn.InstanceNorm1d, **norm_kwargs):
        super().__init__(nn.Conv1d, norm_layer, input_dim, output_dim, kernel_size, groups=groups, padding=padding,
                         stride=stride, dilation=dilation, ndim=1, grid_size=grid_size, base_activation=base_activation,
                         grid_range=grid_range, dropout=dropout, **norm_kwargs)
