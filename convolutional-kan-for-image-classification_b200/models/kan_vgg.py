"""KAN-VGG - drop-in for the reference's ``models/kan_vgg.py`` (``cfgs`` :20-26, ``VGG`` :29-188, ``VGGKAN`` :190-304,
``vggkan`` :307-343), built from this package's CUDA-backed layers through the same string-keyed factories.

Additions: ``cfgs['VGG11']`` (BASELINE config 3; upstream has no such entry).  Quirk kept on purpose:
factory kwargs are filtered by *named* parameters, so ``affine`` / ``degree`` never reach ``kan_conv``."""
from functools import partial
from inspect import signature
from math import prod
from typing import Any, Callable, Dict, List, Optional, Tuple, Union, cast

import torch
import torch.nn as nn

try:  # the reference mixes PyTorchModelHubMixin into VGGKAN; keep it when huggingface_hub is installed
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass

from ..functional import MaxPool2d
from ..layers.kan_conv import CONV_KAN_FACTORY, conv
from .kans import MLP_KAN_FACTORY

cfgs: Dict[str, List[Union[str, int]]] = {
    "VGG16_small": [16, 16, "M", 32, 32, "M", 64, 64, 64, "M", 128, 128, 128, "M", 128, 128, 128],
    "VGG16_kansmall": [8, 8, "M", 16, 16, "M", 32, 32, 32, "M", 64, 64, 64, "M", 64, 64, 64],
    "VGG19_small": [16, 16, "M", 32, 32, "M", 64, 64, 64, 64, "M", 128, 128, 128, 128, "M", 128, 128, 128, 128],
    "VGG16": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512],
    "VGG19": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512],
    "VGG11": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512],
}


class VGG(nn.Module):
    def __init__(self, features: nn.ModuleList, classifier: nn.Module, expected_feature_shape: Tuple = (1, 1)) -> None:
        super().__init__()
        self.features = features
        self.avgpool = nn.AdaptiveAvgPool2d(expected_feature_shape)
        self.classifier = classifier
        self.expected_feature_shape = expected_feature_shape

    @staticmethod
    def make_layers(cfg: List[Union[str, int]], conv_type: str, classifier_type: str, kan_conv: Optional[str] = None,
                    kan_classifier: Optional[Callable[..., nn.Module]] = None, spline_order: int = 3, grid_size: int = 5,
                    base_activation: Optional[Callable[..., nn.Module]] = nn.SiLU, grid_range: List = [-1, 1],
                    dropout: float = 0.0, l1_decay: float = 0.0, groups: int = 1, std_conv_kernel_size: int = 3,
                    std_conv_padding: int = 1, std_conv_bias: bool = True, expected_feature_shape: Tuple = (1, 1),
                    num_input_features: int = 3, num_classes: int = 10, width_scale: int = 1,
                    classifier_dropout: float = 0.5, affine: bool = False,
                    norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d,
                    kan_norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d, degree: int = 3,
                    conv_dropout: float = 0.0, **kwargs: Any):
        if conv_type == 'kanconv':
            if kan_conv is None or kan_conv not in CONV_KAN_FACTORY:
                kan_conv = "KAN"
            builder = CONV_KAN_FACTORY[kan_conv]
            offered = dict(spline_order=spline_order, grid_size=grid_size, base_activation=base_activation,
                           grid_range=grid_range, l1_decay=l1_decay, dropout=conv_dropout, degree=degree, affine=affine,
                           norm_layer=kan_norm_layer, padding=std_conv_padding, groups=groups,
                           kernel_size=std_conv_kernel_size)
            named = signature(builder).parameters
            accepted = {k: v for k, v in offered.items() if k in named}
            accepted.update({k: v for k, v in kwargs.items() if k in named})
            conv_fn = partial(builder, **accepted)
            first_fn = partial(builder, **dict(accepted, dropout=0.0))
        elif conv_type == 'conv':
            bias = std_conv_bias if norm_layer is None else False

            def block(in_c, out_c):
                seq: List[nn.Module] = [conv(in_c, out_c, kernel_size=std_conv_kernel_size, padding=std_conv_padding, bias=bias)]
                if norm_layer is not None:
                    seq.append(norm_layer(out_c, affine=affine) if 'affine' in signature(norm_layer).parameters
                               else norm_layer(out_c))
                seq.append(nn.ReLU(inplace=True))
                return nn.Sequential(*seq)

            conv_fn = first_fn = block
        else:
            raise ValueError(f"unknown conv_type {conv_type!r}")

        layers: List[nn.Module] = []
        in_channels = num_input_features
        for idx, v in enumerate(cfg):
            if v == "M":
                layers.append(MaxPool2d(kernel_size=2, stride=2))
                continue
            out_channels = cast(int, v) * width_scale
            layers.append((first_fn if idx == 0 else conv_fn)(in_channels, out_channels))
            in_channels = out_channels

        feat = in_channels * prod(expected_feature_shape)
        if classifier_type == 'KAN':
            head = nn.Sequential(nn.Dropout(p=classifier_dropout), kan_classifier([feat, num_classes]))
        elif classifier_type == 'Linear':
            head = nn.Sequential(nn.Dropout(p=classifier_dropout), nn.Linear(feat, num_classes))
        elif classifier_type == 'HiddenKAN':
            head = nn.Sequential(kan_classifier([feat, 1024]), nn.Dropout(p=classifier_dropout), nn.Linear(1024, num_classes))
        elif classifier_type == 'VGGKAN':
            head = nn.Sequential(nn.Linear(feat, 1024), nn.ReLU(True), nn.Dropout(p=classifier_dropout),
                                 nn.Linear(1024, 1024), nn.ReLU(True), nn.Dropout(p=classifier_dropout),
                                 kan_classifier([1024, num_classes]))
        elif classifier_type == 'VGG':
            head = nn.Sequential(nn.Linear(feat, 1024), nn.ReLU(True), nn.Dropout(p=classifier_dropout),
                                 nn.Linear(1024, 1024), nn.ReLU(True), nn.Dropout(p=classifier_dropout),
                                 nn.Linear(1024, num_classes))
        else:
            head = nn.Identity()
        return nn.ModuleList(layers), head

    def forward_features(self, x):
        for layer in self.features:
            x = layer(x)
        return x

    def forward(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        x = self.forward_features(x)
        x = self.avgpool(x)
        x = torch.flatten(x, 1)
        return self.classifier(x)


class VGGKAN(VGG, PyTorchModelHubMixin):
    def __init__(self, input_channels: int, num_classes: int, conv_type: str = 'kanconv', kan_conv: Optional[str] = "KAN",
                 kan_classifier: Optional[str] = "KAN", groups: int = 1, spline_order: int = 3, grid_size: int = 5,
                 base_activation: Optional[Callable[..., nn.Module]] = nn.SiLU, grid_range: List = [-1, 1],
                 dropout: float = 0.0, l1_decay: float = 0.0, dropout_linear: float = 0.5, arch: str = 'VGG16',
                 classifier_type: str = 'Linear', expected_feature_shape: Tuple = (1, 1), width_scale: int = 1,
                 affine: bool = False, norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d,
                 kan_norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d, std_conv_kernel_size: int = 3,
                 std_conv_padding: int = 1, std_conv_bias: bool = True, degree: int = 3, conv_dropout: float = 0.0,
                 classifier_spline_order: Optional[int] = None, classifier_grid_size: Optional[int] = None,
                 classifier_base_activation: Optional[Callable[..., nn.Module]] = None,
                 classifier_grid_range: Optional[List] = None, classifier_l1_decay: Optional[float] = None,
                 classifier_dropout: Optional[float] = None, classifier_degree: Optional[int] = None, **kwargs: Any):
        head_factory: Optional[Callable[..., nn.Module]] = None
        head_kind = None
        head_dropout = dropout_linear if classifier_dropout is None else classifier_dropout
        if classifier_type in ('HiddenKAN', 'VGGKAN', 'KAN'):
            head_kind = kan_classifier if kan_classifier else "KAN"
            builder = MLP_KAN_FACTORY[head_kind]
            pick = lambda v, default: default if v is None else v   # noqa: E731
            offered = dict(spline_order=pick(classifier_spline_order, spline_order),
                           grid_size=pick(classifier_grid_size, grid_size),
                           base_activation=pick(classifier_base_activation, nn.SiLU),
                           grid_range=pick(classifier_grid_range, grid_range),
                           l1_decay=pick(classifier_l1_decay, l1_decay), degree=pick(classifier_degree, degree),
                           dropout=0.0, first_dropout=False, bias=False)
            named = signature(builder).parameters
            accepted = {k: v for k, v in offered.items() if k in named}
            accepted.update({k[len('classifier_'):]: v for k, v in kwargs.items()
                             if k.startswith('classifier_') and k[len('classifier_'):] in named})
            head_factory = partial(builder, **accepted)
        conv_tag = f"_{kan_conv.upper()}" if conv_type == 'kanconv' else "_CONV"
        head_tag = classifier_type + (f"_{head_kind.upper()}" if head_factory is not None else "")
        self.name = f"VGGKAN_{head_tag}{conv_tag}_{arch}"
        if arch not in cfgs:
            raise ValueError(f"Unknown arch: {arch}. Available types: {list(cfgs.keys())}")
        extra = {k: v for k, v in kwargs.items() if not k.startswith('classifier_')}
        features, head = self.make_layers(
            cfg=cfgs[arch], conv_type=conv_type, kan_conv=kan_conv, classifier_type=classifier_type,
            kan_classifier=head_factory, spline_order=spline_order, grid_size=grid_size, base_activation=base_activation,
            grid_range=grid_range, dropout=dropout, l1_decay=l1_decay, groups=groups,
            std_conv_kernel_size=std_conv_kernel_size, std_conv_padding=std_conv_padding, std_conv_bias=std_conv_bias,
            expected_feature_shape=expected_feature_shape, num_input_features=input_channels, num_classes=num_classes,
            width_scale=width_scale, classifier_dropout=head_dropout, affine=affine, norm_layer=norm_layer,
            kan_norm_layer=kan_norm_layer, degree=degree, conv_dropout=conv_dropout, **extra)
        super().__init__(features, head, expected_feature_shape)


def vggkan(input_channels: int, num_classes: int, conv_type: str = 'kanconv', kan_conv: Optional[str] = "KAN",
           kan_classifier: Optional[str] = "KAN", classifier_type: str = 'Linear', groups: int = 1, spline_order: int = 3,
           grid_size: int = 5, base_activation: Optional[Callable[..., nn.Module]] = nn.SiLU, grid_range: List = [-1, 1],
           dropout: float = 0.0, l1_decay: float = 0.0, dropout_linear: float = 0.5, arch: str = 'VGG16',
           expected_feature_shape: Tuple[int, int] = (1, 1), width_scale: int = 1, affine: bool = False,
           norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d,
           kan_norm_layer: Optional[Callable[..., nn.Module]] = nn.InstanceNorm2d, std_conv_kernel_size: int = 3,
           std_conv_padding: int = 1, std_conv_bias: bool = True, degree: int = 3, conv_dropout: float = 0.0,
           **kwargs: Any):
    return VGGKAN(input_channels=input_channels, num_classes=num_classes, conv_type=conv_type, kan_conv=kan_conv,
                  kan_classifier=kan_classifier, groups=groups, spline_order=spline_order, grid_size=grid_size,
                  base_activation=base_activation, grid_range=grid_range, dropout=dropout, l1_decay=l1_decay,
                  dropout_linear=dropout_linear, arch=arch, classifier_type=classifier_type,
                  expected_feature_shape=expected_feature_shape, width_scale=width_scale, affine=affine,
                  norm_layer=norm_layer, kan_norm_layer=kan_norm_layer, std_conv_kernel_size=std_conv_kernel_size,
                  std_conv_padding=std_conv_padding, std_conv_bias=std_conv_bias, degree=degree,
                  conv_dropout=conv_dropout, **kwargs)
