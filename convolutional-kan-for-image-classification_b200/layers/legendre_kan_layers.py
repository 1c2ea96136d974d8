"""Legendre KAN convolution layers - drop-in for the reference's ``layers/legendre_kan_layers.py:50-192``; the shared body and the
coefficient table live in ``recurrence_kan_layers.py`` (one CUDA functor for all three-term-recurrence families)."""
from .recurrence_kan_layers import (LegendreKANConvNDLayer, LegendreKANConv1DLayer,  # noqa: F401
                                    LegendreKANConv2DLayer, LegendreKANConv3DLayer)
