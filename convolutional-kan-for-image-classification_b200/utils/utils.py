"""Parameter holder for the Gaussian RBF grid of FastKAN (reference ``utils/utils.py:19-33``).  The RBF itself is
evaluated inside the convolution kernels; this module only keeps the frozen ``grid`` parameter in the state_dict."""
import torch
import torch.nn as nn


class RadialBasisFunction(nn.Module):
    def __init__(self, grid_min: float = -2., grid_max: float = 2., num_grids: int = 8, denominator: float = None):
        super().__init__()
        self.grid = nn.Parameter(torch.linspace(grid_min, grid_max, num_grids), requires_grad=False)
        self.denominator = denominator or (grid_max - grid_min) / (num_grids - 1)
        self._host_params = None

    def host_params(self):
        """(grid values..., denominator) as Python floats, cached so the hot path never syncs the device."""
        if self._host_params is None:
            self._host_params = tuple(float(v) for v in self.grid.detach().cpu().tolist()) + (float(self.denominator),)
        return self._host_params

    def _load_from_state_dict(self, *args, **kwargs):
        self._host_params = None
        return super()._load_from_state_dict(*args, **kwargs)
