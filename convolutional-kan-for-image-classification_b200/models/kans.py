"""KAN MLP heads - drop-in surface for the reference's ``models/kans.py`` (``KAN`` :300-327, ``mlp_kan`` :481-485,
``MLP_KAN_FACTORY`` :556-574).  SURVEY 8(f) rank 2: only the B-spline head is built so far; the other keys raise
``NotImplementedError`` (they are not on the BASELINE hot path - its models use ``classifier_type='Linear'``)."""
from typing import List, Type

import torch.nn as nn

from ..layers.kan_layers import KANLayer
from ..utils.regularization import L1


class KAN(nn.Module):
    def __init__(self, layers_hidden, dropout: float = 0.0, grid_size=5, spline_order=3,
                 base_activation: Type[nn.Module] = nn.GELU, grid_range: List = [-1, 1], l1_decay: float = 0.0,
                 first_dropout: bool = True, **kwargs):
        super().__init__()
        self.layers_hidden, self.grid_size, self.spline_order = layers_hidden, grid_size, spline_order
        self.base_activation, self.grid_range = base_activation, grid_range
        self.layers = nn.ModuleList([])
        if dropout > 0 and first_dropout:
            self.layers.append(nn.Dropout(p=dropout))
        self.num_layers = len(layers_hidden[:-1])
        for i, (fin, fout) in enumerate(zip(layers_hidden[:-1], layers_hidden[1:])):
            layer = KANLayer(fin, fout, grid_size=grid_size, spline_order=spline_order, base_activation=base_activation,
                             grid_range=grid_range)
            if l1_decay > 0 and i != self.num_layers - 1:
                layer = L1(layer, l1_decay)
            self.layers.append(layer)
            if dropout > 0 and i != self.num_layers - 1:
                self.layers.append(nn.Dropout(p=dropout))

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


def mlp_kan(layers_hidden: List[int], dropout: float = 0.0, grid_size: int = 5, spline_order: int = 3,
            base_activation: Type[nn.Module] = nn.GELU, grid_range: List = [-1, 1], l1_decay: float = 0.0,
            first_dropout: bool = True) -> KAN:
    return KAN(layers_hidden, dropout=dropout, grid_size=grid_size, spline_order=spline_order,
               base_activation=base_activation, grid_range=grid_range, l1_decay=l1_decay, first_dropout=first_dropout)


def _out_of_scope(name: str):
    def builder(*args, **kwargs):
        raise NotImplementedError(f"MLP_KAN_FACTORY[{name!r}] is outside the B200 hot path (only 'KAN' is implemented)")
    builder.__name__ = "mlp_" + name.lower()
    return builder


MLP_KAN_FACTORY = {"KAN": mlp_kan}
for _name in ("FastKAN", "LegendreKAN", "BersnsteinKAN", "BesselKAN", "ChebyKAN", "FibonacciKAN", "FourierKAN",
              "GegenbauerKAN", "GRAMKAN", "HermiteKAN", "JacobiKAN", "LaguerreKAN", "LucasKAN", "ReLUKAN", "TaylorKAN",
              "WavKAN"):
    MLP_KAN_FACTORY[_name] = _out_of_scope(_name)
