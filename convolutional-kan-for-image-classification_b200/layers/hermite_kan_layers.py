"""Hermite KAN convolution layers - drop-in for the reference's ``layers/hermite_kan_layers.py:30-192``; the shared body and the
coefficient table live in ``recurrence_kan_layers.py`` (one CUDA functor for all three-term-recurrence families)."""
from .recurrence_kan_layers import (HermiteKANConvNDLayer, HermiteKANConv1DLayer,  # noqa: F401
                                    HermiteKANConv2DLayer, HermiteKANConv3DLayer)
