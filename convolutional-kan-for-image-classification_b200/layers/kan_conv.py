"""String-keyed layer factory - drop-in for the reference's ``layers/kan_conv.py`` (``_calculate_same_padding`` :12-25,
``kan_conv`` :27-69, ``gramkan_conv`` :158-194, ``chebykan_conv`` :197-232, ``fastkan_conv`` :235-276,
``legendrekan_conv`` :120-157, ``besselkan_conv`` :354-388, ``fibonaccikan_conv`` :391-425, ``gegenbauerkan_conv`` :464-500,
``hermitekan_conv`` :502-536, ``jacobikan_conv`` :538-576, ``laguerrekan_conv`` :578-614, ``lucaskan_conv`` :616-650,
``taylorkan_conv`` :692-724, ``CONV_KAN_FACTORY`` :726-745).  Builders keep upstream's signatures; ``padding=None`` means
"same" padding; a layer is wrapped in ``L1`` when ``l1_decay > 0``.  Four basis families of the reference (wavelet, Bernstein,
Fourier, ReLU) are outside the hot path; their keys are present but raise ``NotImplementedError`` instead of
silently building something else."""
from typing import Callable, List, Optional, Tuple, Union

import torch.nn as nn

from .cheby_kan_layers import ChebyKANConv2DLayer
from .fast_kan_layers import FastKANConv2DLayer
from .gram_kan_layers import GRAMKANConv2DLayer
from .kan_layers import KANConv2DLayer
from .recurrence_kan_layers import (BesselKANConv2DLayer, FibonacciKANConv2DLayer, GegenbauerKANConv2DLayer,
                                    HermiteKANConv2DLayer, JacobiKANConv2DLayer, LaguerreKANConv2DLayer,
                                    LegendreKANConv2DLayer, LucasKANConv2DLayer, TaylorKANConv2DLayer)
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


def legendrekan_conv(in_planes: int, out_planes: int, kernel_size: IntOr2, degree: int = 3, groups: int = 1,
                     stride: IntOr2 = 1, dilation: IntOr2 = 1, padding: Optional[IntOr2] = None, dropout: float = 0.0,
                     norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d, l1_decay: float = 0.0,
                     **norm_kwargs) -> LegendreKANConv2DLayer:
    if padding is None:
        padding = _calculate_same_padding(kernel_size, dilation)
    layer = LegendreKANConv2DLayer(input_dim=in_planes, output_dim=out_planes, kernel_size=kernel_size, degree=degree,
                                   stride=stride, padding=padding, dilation=dilation, groups=groups, dropout=dropout,
                                   norm_layer=norm_layer, **norm_kwargs)
    return _finish(layer, l1_decay)


def _poly_builder(layer_class, name: str, extra: Tuple[Tuple[str, float], ...] = (), pass_l1: bool = True):
    """Builders of the template families (kan_conv.py:354-724 share one body).  Like upstream they do NOT forward ``dilation``
    to the layer (it only enters the 'same' padding) and, except TaylorKAN, hand ``l1_decay`` to the layer as well, where it
    ends up in ``**norm_kwargs`` and is filtered out."""
    defaults = dict(extra)

    def builder(in_planes: int, out_planes: int, kernel_size: IntOr2, groups: int = 1, stride: IntOr2 = 1,
                dilation: IntOr2 = 1, padding: Optional[IntOr2] = None, l1_decay: float = 0.0, dropout: float = 0.0,
                degree: int = 3, base_activation: Optional[Callable[..., nn.Module]] = nn.GELU,
                norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d, **norm_kwargs):
        if padding is None:
            padding = _calculate_same_padding(kernel_size, dilation)
        fam = {k: norm_kwargs.pop(k, v) for k, v in defaults.items()}
        if pass_l1:
            norm_kwargs = dict(norm_kwargs, l1_decay=l1_decay)
        layer = layer_class(input_dim=in_planes, output_dim=out_planes, kernel_size=kernel_size, degree=degree, groups=groups,
                            padding=padding, stride=stride, dropout=dropout, base_activation=base_activation,
                            norm_layer=norm_layer, **fam, **norm_kwargs)
        return _finish(layer, l1_decay)
    builder.__name__ = builder.__qualname__ = name
    return builder


besselkan_conv = _poly_builder(BesselKANConv2DLayer, "besselkan_conv")
fibonaccikan_conv = _poly_builder(FibonacciKANConv2DLayer, "fibonaccikan_conv")
gegenbauerkan_conv = _poly_builder(GegenbauerKANConv2DLayer, "gegenbauerkan_conv", (("alpha_param", 0.0),))
hermitekan_conv = _poly_builder(HermiteKANConv2DLayer, "hermitekan_conv")
jacobikan_conv = _poly_builder(JacobiKANConv2DLayer, "jacobikan_conv", (("a", 1.0), ("b", 1.0)))
laguerrekan_conv = _poly_builder(LaguerreKANConv2DLayer, "laguerrekan_conv", (("alpha", 1.0),))
lucaskan_conv = _poly_builder(LucasKANConv2DLayer, "lucaskan_conv")
taylorkan_conv = _poly_builder(TaylorKANConv2DLayer, "taylorkan_conv", pass_l1=False)


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
                                  "(wavelet, Bernstein, Fourier and ReLU bases are not implemented)")
    builder.__name__ = name.lower() + "_conv"
    return builder


CONV_KAN_FACTORY = {
    "KAN": kan_conv,
    "FastKAN": fastkan_conv,
    "GRAMKAN": gramkan_conv,
    "ChebyKAN": chebykan_conv,
    "LegendreKAN": legendrekan_conv,
    "BesselKAN": besselkan_conv,
    "FibonacciKAN": fibonaccikan_conv,
    "GegenbauerKAN": gegenbauerkan_conv,
    "HermiteKAN": hermitekan_conv,
    "JacobiKAN": jacobikan_conv,
    "LaguerreKAN": laguerrekan_conv,
    "LucasKAN": lucaskan_conv,
    "TaylorKAN": taylorkan_conv,
    "conv": conv,
}
for _name in ("WavKAN", "BersnsteinKAN", "FourierKAN", "ReLUKAN"):
    CONV_KAN_FACTORY[_name] = _out_of_scope(_name)
