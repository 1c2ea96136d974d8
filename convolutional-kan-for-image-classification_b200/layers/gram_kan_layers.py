"""Gram-polynomial KAN convolution layers - drop-in for the reference's ``layers/gram_kan_layers.py`` (:85-229).

Differences by design: no ``lru_cache`` on the basis (upstream's never hits and pins 128 activations)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib as L
from .. import functional as KF
from ._base import KANConvBase, check_groups, filter_norm_kwargs, make_dropout, pair


class GRAMKANConvNDLayer(KANConvBase):
    def __init__(self, conv_class, norm_class, conv_w_fun, input_dim, output_dim, degree, kernel_size,
                 base_activation=nn.SiLU, groups=1, padding=0, stride=1, dilation=1, dropout: float = 0.0,
                 ndim: int = 2, **norm_kwargs):
        super().__init__()
        ndim = int(ndim)
        self.input_dim, self.output_dim, self.degree = input_dim, output_dim, degree
        self.kernel_size, self.padding, self.stride, self.dilation = kernel_size, padding, stride, dilation
        self.groups, self.ndim = groups, ndim
        self.base_activation = base_activation() if base_activation is not None else nn.Identity()
        self.conv_w_fun = conv_w_fun
        self.norm_kwargs = norm_kwargs
        self.p_dropout = dropout
        self.dropout = make_dropout(ndim, dropout)
        check_groups(groups, input_dim, output_dim)
        if not isinstance(self.base_activation, nn.SiLU):
            raise NotImplementedError("GRAM KAN convolution: only the SiLU activation of the reference 1D/2D/3D classes "
                                      "is implemented")
        self.base_conv = nn.ModuleList([conv_class(input_dim // groups, output_dim // groups, kernel_size, stride, padding,
                                                   dilation, groups=1, bias=False) for _ in range(groups)])
        self.layer_norm = nn.ModuleList([norm_class(output_dim // groups, **filter_norm_kwargs(norm_class, norm_kwargs))
                                         for _ in range(groups)])
        poly_shape = (groups, output_dim // groups, (input_dim // groups) * (degree + 1)) + tuple(
            kernel_size for _ in range(ndim))
        self.poly_weights = nn.Parameter(torch.randn(*poly_shape))
        self.beta_weights = nn.Parameter(torch.zeros(degree + 1, dtype=torch.float32))
        for m in self.base_conv:
            nn.init.kaiming_uniform_(m.weight, nonlinearity='linear')
        nn.init.kaiming_uniform_(self.poly_weights, nonlinearity='linear')
        nn.init.normal_(self.beta_weights, mean=0.0,
                        std=1.0 / ((kernel_size ** ndim) * input_dim * (degree + 1.0)))
        self._spec = KF.ConvSpec(basis=L.BASIS_GRAM, act=L.ACT_SILU, nb=degree + 1, order=degree, params=(),
                                 kernel=pair(kernel_size, ndim), stride=pair(stride, ndim), padding=pair(padding, ndim, fill=0),
                                 dilation=pair(dilation, ndim), groups=groups)
        self._spec_presquashed = KF.ConvSpec(**dict(self._spec.__dict__, params=(1.0,)))

    def forward(self, x):
        if self.ndim == 3:
            t, spec = None, self._spec
            if self.dropout is not None and self.training:        # Dropout3d on tanh(x), as in the 2-D branch below
                t, spec = self.dropout(torch.tanh(x)), self._spec_presquashed
            z = self._kan_conv3d(spec, x, t, self.beta_weights, [m.weight for m in self.base_conv],
                                 [self.poly_weights[g] for g in range(self.groups)])
            return self._norm_act3d(z, self.layer_norm, L.OUT_SILU)
        x4 = self._to4d(x)
        x_basis, spec = None, self._spec
        if self.dropout is not None and self.training:
            # gram_kan_layers.py:176-179: Dropout acts on t = tanh(x), spline branch only.  The squashing and the mask are
            # elementwise torch ops (autograd differentiates them); the kernels then take t as the basis input and skip
            # their own tanh (ConvSpec.params = (1.0,) = "pre-squashed").
            x_basis = self._to4d(self.dropout(torch.tanh(x)))
            spec = self._spec_presquashed
        w_base = [self._w4d(m.weight) for m in self.base_conv]
        w_poly = [self._w4d(self.poly_weights[g]) for g in range(self.groups)]
        if x_basis is None:
            return self._from4d(self._conv_norm_act(spec, x4, self.beta_weights, w_base, w_poly, self.layer_norm, L.OUT_SILU))
        z = KF.kan_conv(spec, x4, x_basis, self.beta_weights, w_base, w_poly, self.precision)
        return self._from4d(self._norm_act(z, self.layer_norm, L.OUT_SILU))


class GRAMKANConv3DLayer(GRAMKANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                 dropout: float = 0.0, norm_layer=nn.InstanceNorm3d, **norm_kwargs):
        super().__init__(nn.Conv3d, norm_layer, None, input_dim, output_dim, degree, kernel_size, groups=groups,
                         padding=padding, stride=stride, dilation=dilation, ndim=3, dropout=dropout, **norm_kwargs)


class GRAMKANConv2DLayer(GRAMKANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                 dropout: float = 0.0, norm_layer=nn.InstanceNorm2d, **norm_kwargs):
        super().__init__(nn.Conv2d, norm_layer, None, input_dim, output_dim, degree, kernel_size, groups=groups,
                         padding=padding, stride=stride, dilation=dilation, ndim=2, dropout=dropout, **norm_kwargs)


class GRAMKANConv1DLayer(GRAMKANConvNDLayer):
    def __init__(self, input_dim, output_dim, kernel_size, degree=3, groups=1, padding=0, stride=1, dilation=1,
                 dropout: float = 0.0, norm_layer=nn.InstanceNorm1d, **norm_kwargs):
        super().__init__(nn.Conv1d, norm_layer, None, input_dim, output_dim, degree, kernel_size, groups=groups,
                         padding=padding, stride=stride, dilation=dilation, ndim=1, dropout=dropout, **norm_kwargs)
