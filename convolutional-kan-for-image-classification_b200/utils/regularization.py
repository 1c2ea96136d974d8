"""Backward-hook weight-decay wrappers, compatible with the reference's ``utils/regularization.py:57-159`` (the factory
wraps a layer in ``L1`` when ``l1_decay > 0``).  The hook fires after the wrapped module's backward and installs
``weight_decay * sign(p)`` (L1) or ``weight_decay * p`` (L2) as the gradient of parameters whose gradient is still
missing or all-zero - exactly the upstream semantics, including that quirk."""
import torch
import torch.nn as nn


class WeightDecay(nn.Module):
    def __init__(self, module, weight_decay, name: str = None):
        if weight_decay < 0.0:
            raise ValueError("Regularization's weight_decay should be greater than 0.0, got {}".format(weight_decay))
        super().__init__()
        self.module = module
        self.weight_decay = weight_decay
        self.name = name
        self.hook = self.module.register_full_backward_hook(self._weight_decay_hook)

    def remove(self):
        self.hook.remove()

    def _selected(self):
        for pname, p in self.module.named_parameters():
            if self.name is None or self.name in pname:
                yield p

    def _weight_decay_hook(self, *_):
        for p in self._selected():
            if p.grad is None or not bool(torch.any(p.grad != 0.0)):
                p.grad = self.regularize(p)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def extra_repr(self) -> str:
        s = "weight_decay={}".format(self.weight_decay)
        return s if self.name is None else s + ", name={}".format(self.name)

    def regularize(self, parameter):
        raise NotImplementedError


class L2(WeightDecay):
    def regularize(self, parameter):
        return self.weight_decay * parameter.data


class L1(WeightDecay):
    def regularize(self, parameter):
        return self.weight_decay * torch.sign(parameter.data)
