"""Taylor KAN convolution layers - drop-in for the reference's ``layers/taylor_kan_layers.py:40-205``; the shared body and the
coefficient table live in ``recurrence_kan_layers.py`` (one CUDA functor for all three-term-recurrence families)."""
from .recurrence_kan_layers import (TaylorKANConvNDLayer, TaylorKANConv1DLayer,  # noqa: F401
                                    TaylorKANConv2DLayer, TaylorKANConv3DLayer)
