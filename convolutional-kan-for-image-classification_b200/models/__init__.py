from .kan_vgg import VGG, VGGKAN, cfgs, vggkan  # noqa: F401
from .kans import KAN, MLP_KAN_FACTORY, mlp_kan  # noqa: F401
