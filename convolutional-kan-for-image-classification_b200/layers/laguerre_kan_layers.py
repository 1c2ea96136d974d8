"""Laguerre KAN convolution layers - drop-in for the reference's ``layers/laguerre_kan_layers.py:38-212``; the shared body and the
coefficient table live in ``recurrence_kan_layers.py`` (one CUDA functor for all three-term-recurrence families)."""
from .recurrence_kan_layers import (LaguerreKANConvNDLayer, LaguerreKANConv1DLayer,  # noqa: F401
                                    LaguerreKANConv2DLayer, LaguerreKANConv3DLayer)
