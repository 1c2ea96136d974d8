"""Public surface of the package (imported by both ``kanconv_b200`` and the hyphen-named source directory)."""
from . import _lib, build as _build_mod, functional  # noqa: F401
from .functional import (ConvSpec, MaxPool2d, NormSpec, get_precision, kan_conv as kan_conv_op, max_pool2d, norm_act,  # noqa: F401
                         set_precision)
from .layers import *  # noqa: F401,F403
from .layers import CONV_KAN_FACTORY  # noqa: F401


def build(force: bool = False):
    """Compile csrc/*.cu for sm_100a into csrc/libkanconv.so (nvcc cross-compiles without a GPU)."""
    return _build_mod.build(force=force)
