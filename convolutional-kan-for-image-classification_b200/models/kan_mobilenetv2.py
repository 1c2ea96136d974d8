"""KAN-MobileNetV2 - drop-in for the reference's ``models/kan_mobilenetv2.py`` (``_make_divisible`` :11-23,
``ConvNormActivation`` :25-76, ``InvertedResidual`` :79-166, ``MobileNetV2KAN`` :169-433, ``mobilenet_v2_kan`` :435-495),
built from this package's CUDA-backed layers through the same string-keyed factories.  BASELINE config 4 is
``mobilenet_v2_kan(num_classes=1000, kan_conv='FastKAN', classifier_type='Linear')``: 35 FastKAN convolutions (34 of them
1x1, BatchNorm on the RBF input, 5 grid points) + 17 depthwise ``Conv2d+BN+ReLU6`` blocks that stay plain torch modules.

Kept quirks: every factory call receives ``activation_layer=...`` / ``norm_layer=...`` / ``**factory_kwargs`` and relies on
the builders swallowing what they do not name; ``_initialize_weights`` re-initialises *every* ``nn.Conv2d`` in the tree -
including the parameter-holder convs inside the KAN layers - with ``kaiming_normal_(fan_out)``."""
from functools import partial
from inspect import signature
from typing import Any, Callable, List, Optional, Tuple, Union

import torch.nn as nn
from torch import Tensor

try:
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass

from ..layers.kan_conv import CONV_KAN_FACTORY, _calculate_same_padding
from .kans import MLP_KAN_FACTORY

DEFAULT_SETTING = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]
KAN_SMALL_SETTING = [[1, 16, 1, 1], [6, 24, 1, 2], [6, 32, 1, 2], [6, 48, 1, 2], [6, 64, 1, 1], [6, 96, 1, 2], [6, 160, 1, 1]]


def _make_divisible(v: float, divisor: int, min_value: Optional[int] = None) -> int:
    """Round a channel count to a multiple of ``divisor`` without dropping more than 10 % (TF-slim rule)."""
    floor = divisor if min_value is None else min_value
    rounded = max(floor, int(v + divisor / 2) // divisor * divisor)
    return rounded + divisor if rounded < 0.9 * v else rounded


class ConvNormActivation(nn.Sequential):
    def __init__(self, in_planes: int, out_planes: int, kernel_size: Union[int, Tuple[int, int]] = 3,
                 stride: Union[int, Tuple[int, int]] = 1, groups: int = 1,
                 norm_layer: Optional[Callable[..., nn.Module]] = nn.BatchNorm2d,
                 activation_layer: Optional[Callable[..., nn.Module]] = nn.ReLU, dilation: int = 1,
                 inplace: Optional[bool] = None, bias: Optional[bool] = None, conv_layer: Callable[..., nn.Module] = nn.Conv2d,
                 padding: Optional[Union[int, Tuple[int, int], str]] = None, affine: bool = True) -> None:
        if padding is None:
            padding = _calculate_same_padding(kernel_size, dilation)
        if bias is None:
            bias = norm_layer is None or not affine
        if inplace is None and activation_layer is not None:
            inplace = activation_layer in (nn.ReLU, nn.ReLU6)
        mods: List[nn.Module] = [conv_layer(in_planes, out_planes, kernel_size, stride=stride, padding=padding,
                                            dilation=dilation, groups=groups, bias=bias)]
        if norm_layer is not None:
            mods.append(norm_layer(out_planes, affine=affine))
        if activation_layer is not None:
            kw = {"inplace": bool(inplace)} if "inplace" in signature(activation_layer).parameters else {}
            mods.append(activation_layer(**kw))
        super().__init__(*mods)
        self.out_channels = out_planes


class InvertedResidual(nn.Module):
    def __init__(self, input_dim: int, output_dim: int, stride: int, expand_ratio: int,
                 norm_layer: Optional[Callable[..., nn.Module]], activation_layer: Callable[..., nn.Module],
                 conv_layer_factory: Callable[..., nn.Module], replace_depthwise: bool = False, **factory_kwargs) -> None:
        super().__init__()
        self.stride = stride
        hidden = int(round(input_dim * expand_ratio))
        self.use_res_connect = stride == 1 and input_dim == output_dim
        common = dict(norm_layer=norm_layer, **factory_kwargs)
        blocks: List[nn.Module] = []
        if expand_ratio != 1:      # 1x1 expansion through the KAN factory
            blocks.append(conv_layer_factory(in_planes=input_dim, out_planes=hidden, kernel_size=1, stride=1,
                                             activation_layer=activation_layer, **common))
        if replace_depthwise:      # depthwise 3x3 as a grouped KAN convolution (groups == channels)
            blocks.append(conv_layer_factory(in_planes=hidden, out_planes=hidden, kernel_size=3, stride=stride, groups=hidden,
                                             activation_layer=activation_layer, **common))
        else:                      # ... or the classic Conv2d + norm + activation
            blocks.append(ConvNormActivation(in_planes=hidden, out_planes=hidden, kernel_size=3, stride=stride, groups=hidden,
                                             norm_layer=norm_layer if norm_layer is not None else nn.Identity,
                                             activation_layer=activation_layer if activation_layer is not None else nn.Identity,
                                             conv_layer=nn.Conv2d, bias=norm_layer is None,
                                             affine=factory_kwargs.get('affine', True)))
        blocks.append(conv_layer_factory(in_planes=hidden, out_planes=output_dim, kernel_size=1, stride=1, activation_layer=None,
                                         **common))          # linear 1x1 projection
        self.conv = nn.Sequential(*blocks)
        self.out_channels = output_dim
        self._is_cn = stride > 1

    def forward(self, x: Tensor) -> Tensor:
        y = self.conv(x)
        return x + y if self.use_res_connect else y


class MobileNetV2KAN(nn.Module, PyTorchModelHubMixin):
    def __init__(self, num_classes: int = 1000, width_mult: float = 1.0,
                 inverted_residual_setting: Optional[List[List[int]]] = None, round_nearest: int = 8, dropout: float = 0.2,
                 input_channels: int = 3, arch: str = "default", conv_type: str = 'kanconv', kan_conv: Optional[str] = "KAN",
                 kan_classifier: Optional[str] = "KAN", classifier_type: str = 'Linear', groups: int = 1, degree: int = 3,
                 spline_order: int = 3, grid_size: int = 5, base_activation: Optional[Callable[..., nn.Module]] = nn.SiLU,
                 grid_range: List = [-1, 1], l1_decay: float = 0.0, affine: bool = True,
                 norm_layer: Optional[Callable[..., nn.Module]] = nn.BatchNorm2d,
                 kan_norm_layer: Optional[Callable[..., nn.Module]] = nn.BatchNorm2d, replace_depthwise: bool = False,
                 classifier_spline_order: Optional[int] = None, classifier_grid_size: Optional[int] = None,
                 classifier_base_activation: Optional[Callable[..., nn.Module]] = None,
                 classifier_grid_range: Optional[List] = None, classifier_l1_decay: Optional[float] = None,
                 classifier_dropout: Optional[float] = None, classifier_degree: Optional[int] = None, **kwargs: Any) -> None:
        super().__init__()
        setting = DEFAULT_SETTING if inverted_residual_setting is None else inverted_residual_setting
        first_stride = 2
        if arch in ("small", "kan_small"):
            first_stride = 1
        if arch == "kan_small":
            setting = KAN_SMALL_SETTING
        if len(setting) == 0 or len(setting[0]) != 4:
            raise ValueError(f"inverted_residual_setting should be non-empty or a 4-element list, got {setting}")
        act = nn.ReLU6
        if kan_norm_layer is None:
            kan_norm_layer = norm_layer
        shared = dict(spline_order=spline_order, grid_size=grid_size, base_activation=base_activation, grid_range=grid_range,
                      l1_decay=l1_decay, dropout=kwargs.get('conv_dropout', 0.0), degree=degree, affine=affine)
        factory_kwargs = dict(shared, **kwargs)

        if conv_type == 'kanconv':
            if kan_conv is None or kan_conv not in CONV_KAN_FACTORY:
                kan_conv = "KAN"
            builder = CONV_KAN_FACTORY[kan_conv]
            bound = dict(shared, groups=groups, norm_layer=kan_norm_layer)
            bound.update({k: v for k, v in kwargs.items() if k in signature(builder).parameters})
            make_conv: Callable[..., nn.Module] = partial(builder, **bound)
        elif conv_type == 'conv':
            def make_conv(in_planes, out_planes, kernel_size, stride=1, padding=None, groups=1, dilation=1,
                          norm_layer=norm_layer, activation_layer=act, affine=affine, **_ignored):
                if padding is None:
                    padding = _calculate_same_padding(kernel_size, dilation if isinstance(kernel_size, int) else (dilation, dilation))
                seq: List[nn.Module] = [nn.Conv2d(in_planes, out_planes, kernel_size, stride=stride, padding=padding,
                                                  dilation=dilation, groups=groups, bias=norm_layer is None or not affine)]
                if norm_layer is not None:
                    seq.append(norm_layer(out_planes, affine=affine))
                if activation_layer is not None:
                    seq.append(activation_layer(inplace=True))
                return nn.Sequential(*seq)
        else:
            raise ValueError(f"unknown conv_type {conv_type!r}")

        cin = _make_divisible(32 * width_mult, round_nearest)
        self.last_channel = _make_divisible(1280 * max(1.0, width_mult), round_nearest)
        feats: List[nn.Module] = [make_conv(in_planes=input_channels, out_planes=cin, kernel_size=3, stride=first_stride,
                                            activation_layer=act, norm_layer=norm_layer, **factory_kwargs)]
        for t, c, n, s in setting:
            cout = _make_divisible(c * width_mult, round_nearest)
            for i in range(n):
                feats.append(InvertedResidual(cin, cout, s if i == 0 else 1, expand_ratio=t, norm_layer=norm_layer,
                                              activation_layer=act, conv_layer_factory=make_conv,
                                              replace_depthwise=replace_depthwise, **factory_kwargs))
                cin = cout
        feats.append(make_conv(in_planes=cin, out_planes=self.last_channel, kernel_size=1, activation_layer=act,
                               norm_layer=norm_layer, **factory_kwargs))
        self.features = nn.Sequential(*feats)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))

        pick = lambda v, default: default if v is None else v   # noqa: E731
        head_dropout = pick(classifier_dropout, dropout)
        if classifier_type == 'KAN':
            head_kind = kan_classifier if kan_classifier in MLP_KAN_FACTORY else 'KAN'
            head_builder = MLP_KAN_FACTORY[head_kind]
            offered = dict(dropout=head_dropout, spline_order=pick(classifier_spline_order, spline_order),
                           grid_size=pick(classifier_grid_size, grid_size),
                           base_activation=pick(classifier_base_activation, base_activation),
                           grid_range=pick(classifier_grid_range, grid_range), l1_decay=pick(classifier_l1_decay, l1_decay),
                           degree=pick(classifier_degree, degree), first_dropout=False)
            named = signature(head_builder).parameters
            head = partial(head_builder, **{k: v for k, v in offered.items() if k in named})
        else:
            head = lambda layers_hidden: nn.Linear(layers_hidden[0], layers_hidden[1])   # noqa: E731
        self.classifier = nn.Sequential()
        self.classifier.add_module("flatten", nn.Flatten(1))
        self.classifier.add_module("head_dropout", nn.Dropout(p=head_dropout))
        self.classifier.add_module("fc", head(layers_hidden=[self.last_channel, num_classes]))
        self._initialize_weights()

        conv_tag = f"_{kan_conv.upper()}" if conv_type == 'kanconv' else "_CONV"
        head_tag = classifier_type + (f"_{(kan_classifier or 'KAN').upper()}" if classifier_type in MLP_KAN_FACTORY else "")
        rdw = "_RDW" if replace_depthwise and conv_type == 'kanconv' else ""
        self.name = f"MobileNetV2KAN_{head_tag}{conv_tag}{rdw}_{arch}"

    def _initialize_weights(self) -> None:
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.BatchNorm2d, nn.GroupNorm)):
                if m.weight is not None:
                    nn.init.ones_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)

    def forward(self, x: Tensor) -> Tensor:
        return self.classifier(self.avgpool(self.features(x)))


def mobilenet_v2_kan(num_classes: int = 1000, width_mult: float = 1.0, input_channels: int = 3, dropout: float = 0.2,
                     conv_type: str = 'kanconv', arch: str = "default", kan_conv: Optional[str] = "KAN",
                     kan_classifier: Optional[str] = "KAN", classifier_type: str = 'Linear', groups: int = 1,
                     spline_order: int = 3, grid_size: int = 5, base_activation: Optional[Callable[..., nn.Module]] = nn.SiLU,
                     grid_range: List = [-1, 1], l1_decay: float = 0.0, affine: bool = True,
                     norm_layer: Optional[Callable[..., nn.Module]] = nn.BatchNorm2d,
                     kan_norm_layer: Optional[Callable[..., nn.Module]] = nn.BatchNorm2d, replace_depthwise: bool = False,
                     classifier_spline_order: Optional[int] = None, classifier_grid_size: Optional[int] = None,
                     classifier_base_activation: Optional[Callable[..., nn.Module]] = None,
                     classifier_grid_range: Optional[List] = None, classifier_l1_decay: Optional[float] = None,
                     classifier_dropout: Optional[float] = None, classifier_degree: Optional[int] = None,
                     degree: Optional[int] = 3, **kwargs: Any) -> MobileNetV2KAN:
    return MobileNetV2KAN(num_classes=num_classes, width_mult=width_mult, input_channels=input_channels, dropout=dropout,
                          conv_type=conv_type, arch=arch, kan_conv=kan_conv, kan_classifier=kan_classifier,
                          classifier_type=classifier_type, groups=groups, spline_order=spline_order, grid_size=grid_size,
                          base_activation=base_activation, grid_range=grid_range, l1_decay=l1_decay, affine=affine,
                          norm_layer=norm_layer, kan_norm_layer=kan_norm_layer, replace_depthwise=replace_depthwise,
                          classifier_spline_order=classifier_spline_order, classifier_grid_size=classifier_grid_size,
                          classifier_base_activation=classifier_base_activation, classifier_grid_range=classifier_grid_range,
                          classifier_l1_decay=classifier_l1_decay, classifier_dropout=classifier_dropout,
                          classifier_degree=classifier_degree, degree=degree, **kwargs)
