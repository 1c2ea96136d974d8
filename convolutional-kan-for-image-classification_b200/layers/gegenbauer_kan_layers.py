"""Gegenbauer KAN convolution layers - drop-in for the reference's ``layers/gegenbauer_kan_layers.py:34-246``; the shared body and the
coefficient table live in ``recurrence_kan_layers.py`` (one CUDA functor for all three-term-recurrence families)."""
from .recurrence_kan_layers import (GegenbauerKANConvNDLayer, GegenbauerKANConv1DLayer,  # noqa: F401
                                    GegenbauerKANConv2DLayer, GegenbauerKANConv3DLayer)
