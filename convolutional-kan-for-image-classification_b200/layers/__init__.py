from .kan_layers import KANConvNDLayer, KANConv1DLayer, KANConv2DLayer, KANConv3DLayer, KANLayer  # noqa: F401
from .cheby_kan_layers import (ChebyKANConvNDLayer, ChebyKANConv1DLayer, ChebyKANConv2DLayer,  # noqa: F401
                               ChebyKANConv3DLayer)
from .gram_kan_layers import GRAMKANConvNDLayer, GRAMKANConv1DLayer, GRAMKANConv2DLayer, GRAMKANConv3DLayer  # noqa: F401
from .fast_kan_layers import FastKANConvNDLayer, FastKANConv1DLayer, FastKANConv2DLayer, FastKANConv3DLayer  # noqa: F401
from .kan_conv import (CONV_KAN_FACTORY, _calculate_same_padding, chebykan_conv, conv, fastkan_conv,  # noqa: F401
                       gramkan_conv, kan_conv)
