"""String-keyed layer factory - drop-in for the reference's ``layers/kan_conv.py`` (``_calculate_same_padding`` :12-25,
``kan_conv`` :27-69, ``gramkan_conv`` :158-194, ``chebykan_conv`` :197-232, ``fastkan_conv`` :235-276,
``CONV_KAN_FACTORY`` :726-745).  Builders keep upstream's signatures; ``padding=None`` means "same" padding; a layer is
wrapped in ``L1`` when ``l1_decay > 0``.  The 13 other basis families of the reference are outside the hot path
(SURVEY 8(f)); their keys are present but raise ``NotImplementedError`` instead of silently building something else."""
from typing import Callable, List, Optional, Tuple, Union

import torch.nn as nn

from .cheby_kan_layers import ChebyKANConv2DLayer
from .fast_kan_layers import FastKANConv2DLayer
from .gram_kan_layers import GRAMKANConv2DLayer
from .kan_layers import KANConv2DLayer
from ..utils.regularization import L1, L2  # noqa: F401

IntOr2 = Union[int, Tuple[int, int]]


def _calculate_same_padding(kernel_size: IntOr2, dilation: IntOr2) -> IntOr2:
    """'same' padding for stride 1: (d * (k - 1)) // 2 per axis; an int when both axes agree on a square kernel."""
    kh, kw = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
    dh, dw = (dilation, dilation) if isinstance(dilation, int) else dilation
    ph, pw = (dh * (kh - 1)) // 2, (dw * (kw - 1)) // 2
    return ph if (ph == pw and kh == kw) else (ph, pw)


def _finish(layer: nn.Module, l1_decay: float) -> nn.Module:
    return L1(layer, l1_decay) if l1_decay > 0 else layer


def kan_conv(in_planes: int, out_planes: int, kernel_size: IntOr2, spline_order: int = 3, groups: int = 1,
             stride: IntOr2 = 1, dilation: IntOr2 = 1, padding: Optional[IntOr2] = None, grid_size: int = 5,
             base_activation: Optional[Callable[..., nn.Module]] = nn.GELU, grid_range: List = [-1, 1],
             l1_decay: float = 0.0, dropout: float = 0.0,
             norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d, **norm_kwargs) -> KANConv2DLayer:
    if padding is None:
        padding = _calculate_same_padding(kernel_size, dilation)
    layer = KANConv2DLayer(input_dim=in_planes, output_dim=out_planes, kernel_size=kernel_size, spline_order=spline_order,
                           stride=stride, padding=padding, dilation=dilation, groups=groups, grid_size=grid_size,
                           base_activation=base_activation, grid_range=grid_range, dropout=dropout, norm_layer=norm_layer,
                           **norm_kwargs)
    return _finish(layer, l1_decay)


def gramkan_conv(in_planes: int, out_planes: int, kernel_size: IntOr2, degree: int = 3, groups: int = 1,
                 stride: IntOr2 = 1, dilation: IntOr2 = 1, padding: Optional[IntOr2] = None, dropout: float = 0.0,
                 norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d, l1_decay: float = 0.0,
                 **norm_kwargs) -> GRAMKANConv2DLayer:
    if padding is None:
        padding = _calculate_same_padding(kernel_size, dilation)
    layer = GRAMKANConv2DLayer(input_dim=in_planes, output_dim=out_planes, kernel_size=kernel_size, degree=degree,
                               stride=stride, padding=padding, dilation=dilation, groups=groups, dropout=dropout,
                               norm_layer=norm_layer, **norm_kwargs)
    return _finish(layer, l1_decay)


def chebykan_conv(in_planes: int, out_planes: int, kernel_size: IntOr2, degree: int = 3, groups: int = 1,
                  stride: IntOr2 = 1, dilation: IntOr2 = 1, padding: Optional[IntOr2] = None, l1_decay: float = 0.0,
                  dropout: float = 0.0, norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d,
                  **norm_kwargs) -> ChebyKANConv2DLayer:
    if padding is None:
        padding = _calculate_same_padding(kernel_size, dilation)
    layer = ChebyKANConv2DLayer(input_dim=in_planes, output_dim=out_planes, kernel_size=kernel_size, degree=degree,
                                stride=stride, padding=padding, dilation=dilation, groups=groups, dropout=dropout,
                                norm_layer=norm_layer, **norm_kwargs)
    return _finish(layer, l1_decay)


def fastkan_conv(in_planes: int, out_planes: int, kernel_size: IntOr2, groups: int = 1, stride: IntOr2 = 1,
                 dilation: IntOr2 = 1, padding: Optional[IntOr2] = None, grid_size: int = 8,
                 base_activation: Callable[..., nn.Module] = nn.SiLU, grid_range: List = [-2, 2], l1_decay: float = 0.0,
                 dropout: float = 0.0, norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d,
                 **norm_kwargs) -> FastKANConv2DLayer:
    if padding is None:
        padding = _calculate_same_padding(kernel_size, dilation)
    # upstream forwards l1_decay into the layer's **norm_kwargs as well (kan_conv.py:269); it is filtered out there
    layer = FastKANConv2DLayer(input_dim=in_planes, output_dim=out_planes, kernel_size=kernel_size, stride=stride,
                               padding=padding, dilation=dilation, groups=groups, grid_size=grid_size,
                               base_activation=base_activation, grid_range=grid_range, dropout=dropout,
                               l1_decay=l1_decay, norm_layer=norm_layer, **norm_kwargs)
    return _finish(layer, l1_decay)


def conv(in_planes: int, out_planes: int, kernel_size: IntOr2, groups: int = 1, stride: IntOr2 = 1, dilation: IntOr2 = 1,
         padding: Optional[IntOr2] = None, bias: bool = False, **kwargs) -> nn.Conv2d:
    """Plain nn.Conv2d with 'same' padding by default (the factory's non-KAN entry)."""
    if padding is None:
        padding = _calculate_same_padding(kernel_size, dilation)
    return nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding, dilation=dilation,
                     groups=groups, bias=bias)


def _out_of_scope(name: str):
    def builder(*args, **kwargs):
        raise NotImplementedError(f"CONV_KAN_FACTORY[{name!r}]: this basis family is outside the B200 hot path "
                                  "(B-spline 'KAN', 'FastKAN', 'GRAMKAN', 'ChebyKAN' are implemented)")
    builder.__name__ = name.lower() + "_conv"
    return builder


CONV_KAN_FACTORY = {
    "KAN": kan_conv,
    "FastKAN": fastkan_conv,
    "GRAMKAN": gramkan_conv,
    "ChebyKAN": chebykan_conv,
    "conv": conv,
}
for _name in ("LegendreKAN", "WavKAN", "BersnsteinKAN", "BesselKAN", "FibonacciKAN", "FourierKAN", "GegenbauerKAN",
              "HermiteKAN", "JacobiKAN", "LaguerreKAN", "LucasKAN", "ReLUKAN", "TaylorKAN"):
    CONV_KAN_FACTORY[_name] = _out_of_scope(_name)
