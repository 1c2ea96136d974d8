from .kan_layers import KANConvNDLayer, KANConv1DLayer, KANConv2DLayer, KANConv3DLayer, KANLayer  # noqa: F401
from .cheby_kan_layers import (ChebyKANConvNDLayer, ChebyKANConv1DLayer, ChebyKANConv2DLayer,  # noqa: F401
                               ChebyKANConv3DLayer)
from .gram_kan_layers import GRAMKANConvNDLayer, GRAMKANConv1DLayer, GRAMKANConv2DLayer, GRAMKANConv3DLayer  # noqa: F401
from .fast_kan_layers import FastKANConvNDLayer, FastKANConv1DLayer, FastKANConv2DLayer, FastKANConv3DLayer  # noqa: F401
from .recurrence_kan_layers import (  # noqa: F401
    HermiteKANConvNDLayer, HermiteKANConv1DLayer, HermiteKANConv2DLayer, HermiteKANConv3DLayer,
    GegenbauerKANConvNDLayer, GegenbauerKANConv1DLayer, GegenbauerKANConv2DLayer, GegenbauerKANConv3DLayer,
    LaguerreKANConvNDLayer, LaguerreKANConv1DLayer, LaguerreKANConv2DLayer, LaguerreKANConv3DLayer,
    LucasKANConvNDLayer, LucasKANConv1DLayer, LucasKANConv2DLayer, LucasKANConv3DLayer,
    FibonacciKANConvNDLayer, FibonacciKANConv1DLayer, FibonacciKANConv2DLayer, FibonacciKANConv3DLayer,
    BesselKANConvNDLayer, BesselKANConv1DLayer, BesselKANConv2DLayer, BesselKANConv3DLayer,
    TaylorKANConvNDLayer, TaylorKANConv1DLayer, TaylorKANConv2DLayer, TaylorKANConv3DLayer,
    LegendreKANConvNDLayer, LegendreKANConv1DLayer, LegendreKANConv2DLayer, LegendreKANConv3DLayer,
    JacobiKANConvNDLayer, JacobiKANConv1DLayer, JacobiKANConv2DLayer, JacobiKANConv3DLayer)
from .kan_conv import (CONV_KAN_FACTORY, _calculate_same_padding, besselkan_conv, chebykan_conv, conv,  # noqa: F401
                       fastkan_conv, fibonaccikan_conv, gegenbauerkan_conv, gramkan_conv, hermitekan_conv, jacobikan_conv,
                       kan_conv, laguerrekan_conv, legendrekan_conv, lucaskan_conv, taylorkan_conv)
