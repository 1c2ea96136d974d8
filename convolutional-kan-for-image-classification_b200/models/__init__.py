from .kan_vgg import VGG, VGGKAN, cfgs, vggkan  # noqa: F401
from .kan_mobilenetv2 import ConvNormActivation, InvertedResidual, MobileNetV2KAN, mobilenet_v2_kan  # noqa: F401
from .kans import KAN, MLP_KAN_FACTORY, mlp_kan  # noqa: F401
