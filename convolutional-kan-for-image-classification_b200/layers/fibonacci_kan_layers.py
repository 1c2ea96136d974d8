"""Fibonacci KAN convolution layers - drop-in for the reference's ``layers/fibonacci_kan_layers.py:41-263``; the shared body and the
coefficient table live in ``recurrence_kan_layers.py`` (one CUDA functor for all three-term-recurrence families)."""
from .recurrence_kan_layers import (FibonacciKANConvNDLayer, FibonacciKANConv1DLayer,  # noqa: F401
                                    FibonacciKANConv2DLayer, FibonacciKANConv3DLayer)
