"""Lucas KAN convolution layers - drop-in for the reference's ``layers/lucas_kan_layers.py:40-228``; the shared body and the
coefficient table live in ``recurrence_kan_layers.py`` (one CUDA functor for all three-term-recurrence families)."""
from .recurrence_kan_layers import (LucasKANConvNDLayer, LucasKANConv1DLayer,  # noqa: F401
                                    LucasKANConv2DLayer, LucasKANConv3DLayer)
