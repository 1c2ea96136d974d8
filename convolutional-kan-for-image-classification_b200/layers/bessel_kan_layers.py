"""Bessel KAN convolution layers - drop-in for the reference's ``layers/bessel_kan_layers.py:38-200``; the shared body and the
coefficient table live in ``recurrence_kan_layers.py`` (one CUDA functor for all three-term-recurrence families)."""
from .recurrence_kan_layers import (BesselKANConvNDLayer, BesselKANConv1DLayer,  # noqa: F401
                                    BesselKANConv2DLayer, BesselKANConv3DLayer)
