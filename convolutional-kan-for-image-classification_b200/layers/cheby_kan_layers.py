"""Chebyshev KAN convolution layers - drop-in for the reference's ``layers/cheby_kan_layers.py`` (:39-141)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib as L
from .. import functional as KF
from ._base import KANConvBase, check_groups, filter_norm_kwargs, make_dropout, pair


class ChebyKANConvNDLayer(KANConvBase):
    def __init__(self, conv_class, norm_layer, input_dim, output_dim, degree, kernel_size,
                 groups=1, padding=0, stride=1, dilation=1, ndim: int = 2, dropout=0.0, **norm_kwargs):
        super().__init__()
        self.input_dim, self.output_dim, self.degree = input_dim, output_dim, degree
        self.kernel_size, self.padding, self.stride, self.dilation = kernel_size, padding, stride, dilation
        self.groups, self.ndim = groups, ndim
        self.norm_kwargs = norm_kwargs
        self.epsilon = 1e-7
        self.dropout = make_dropout(ndim, dropout)
        check_groups(groups, input_dim, output_dim)
        self.layer_norm = nn.ModuleList([norm_layer(output_dim // groups, **filter_norm_kwargs(norm_layer, norm_kwargs))
                                         for _ in range(groups)])
        self.poly_conv = nn.ModuleList([conv_class((degree + 1) * input_dim // groups, output_dim // groups, kernel_size,
                                                   stride, padding, dilation, groups=1, bias=False)
                                        for _ in range(groups)])
        self.register_buffer("arange", torch.arange(0, degree + 1, 1).view(1, 1, -1, *([1] * ndim)))
        for m in self.poly_conv:
            # the reference draws normal_ and then overwrites it with kaiming_normal_ (cheby_kan_layers.py:88-90)
            nn.init.normal_(m.weight, mean=0.0, std=1 / (input_dim * (degree + 1) * kernel_size ** ndim))
            nn.init.kaiming_normal_(m.weight, mode='fan_in', nonlinearity='relu')
        self._spec = KF.ConvSpec(basis=L.BASIS_CHEBY, act=L.ACT_NONE, nb=degree + 1, order=degree, params=(),
                                 kernel=pair(kernel_size, ndim), stride=pair(stride, ndim), padding=pair(padding, ndim, fill=0),
                                 dilation=pair(dilation, ndim), groups=groups)

    def forward(self, x):
        if self.ndim == 3:
            z = self._kan_conv3d(self._spec, x, None, None, [], [m.weight for m in self.poly_conv])
            y = self._norm_act3d(z, self.layer_norm, L.OUT_NONE)
            return y if self.dropout is None else self.dropout(y)
        x4 = self._to4d(x)
        y = self._from4d(self._conv_norm_act(self._spec, x4, None, [], [self._w4d(m.weight) for m in self.poly_conv],
                                             self.layer_norm, L.OUT_NONE))
        if self.dropout is not None:
            y = self.dropout(y)
        return y


class ChebyKANConv3DLayer(ChebyKANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                 dropout=0.0, norm_layer=nn.InstanceNorm3d, **norm_kwargs):
        super().__init__(nn.Conv3d, norm_layer, input_dim, output_dim, degree, kernel_size, groups=groups, padding=padding,
                         stride=stride, dilation=dilation, ndim=3, dropout=dropout, **norm_kwargs)


class ChebyKANConv2DLayer(ChebyKANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                 dropout=0.0, norm_layer=nn.InstanceNorm2d, **norm_kwargs):
        super().__init__(nn.Conv2d, norm_layer, input_dim, output_dim, degree, kernel_size, groups=groups, padding=padding,
                         stride=stride, dilation=dilation, ndim=2, dropout=dropout, **norm_kwargs)


class ChebyKANConv1DLayer(ChebyKANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                 dropout=0.0, norm_layer=nn.InstanceNorm1d, **norm_kwargs):
        super().__init__(nn.Conv1d, norm_layer, input_dim, output_dim, degree, kernel_size, groups=groups, padding=padding,
                         stride=stride, dilation=dilation, ndim=1, dropout=dropout, **norm_kwargs)
