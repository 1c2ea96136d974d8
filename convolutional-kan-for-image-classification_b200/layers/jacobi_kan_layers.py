"""Jacobi KAN convolution layers - drop-in for the reference's ``layers/jacobi_kan_layers.py:56-208``; the shared body and the
coefficient table live in ``recurrence_kan_layers.py`` (one CUDA functor for all three-term-recurrence families)."""
from .recurrence_kan_layers import (JacobiKANConvNDLayer, JacobiKANConv1DLayer,  # noqa: F401
                                    JacobiKANConv2DLayer, JacobiKANConv3DLayer)
