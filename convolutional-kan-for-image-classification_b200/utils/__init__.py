from .utils import RadialBasisFunction  # noqa: F401
from .regularization import L1, L2, WeightDecay  # noqa: F401
