"""Parameter constraints with the GPyTorch interface (``gpytorch.constraints``), which is
absent from this image.  pgmuvi picks them at pgmuvi/lightcurve.py:3817-4008; the engine
reads ``lower_bound`` / ``upper_bound`` and the class name from either these objects or the
real GPyTorch ones (duck-typed).
"""
from __future__ import annotations

import math

import torch
from torch import nn


def _inv_softplus(v):
    return torch.where(v > 30, v, v + torch.log(-torch.expm1(-v.clamp_min(1e-300))))


class Interval(nn.Module):
    """value = lower + (upper - lower) * sigmoid(raw)"""

    enforced = True

    def __init__(self, lower_bound, upper_bound, initial_value=None):
        super().__init__()
        lb = torch.as_tensor(lower_bound, dtype=torch.get_default_dtype())
        ub = torch.as_tensor(upper_bound, dtype=torch.get_default_dtype())
        if torch.any(lb >= ub):
            raise RuntimeError("Got parameter bounds with empty intervals.")
        self.register_buffer("lower_bound", lb.clone().detach())
        self.register_buffer("upper_bound", ub.clone().detach())
        self._initial_value = initial_value

    def transform(self, tensor):
        lb, ub = self.lower_bound.to(tensor), self.upper_bound.to(tensor)
        return lb + (ub - lb) * torch.sigmoid(tensor)

    def inverse_transform(self, value):
        lb, ub = self.lower_bound.to(value), self.upper_bound.to(value)
        u = (value - lb) / (ub - lb)
        return torch.log(u) - torch.log1p(-u)

    @property
    def initial_value(self):
        return self._initial_value

    def check(self, value):
        return bool(torch.all(value <= self.upper_bound) and torch.all(value >= self.lower_bound))

    def __repr__(self):
        return f"{type(self).__name__}({float(self.lower_bound):.3E}, {float(self.upper_bound):.3E})"


class GreaterThan(Interval):
    """value = softplus(raw) + lower"""

    def __init__(self, lower_bound, initial_value=None):
        nn.Module.__init__(self)
        lb = torch.as_tensor(lower_bound, dtype=torch.get_default_dtype())
        self.register_buffer("lower_bound", lb.clone().detach())
        self.register_buffer("upper_bound", torch.full_like(lb, math.inf))
        self._initial_value = initial_value

    def transform(self, tensor):
        return torch.nn.functional.softplus(tensor) + self.lower_bound.to(tensor)

    def inverse_transform(self, value):
        return _inv_softplus(value - self.lower_bound.to(value))

    def __repr__(self):
        return f"{type(self).__name__}({float(self.lower_bound):.3E})"


class Positive(GreaterThan):
    def __init__(self, initial_value=None):
        super().__init__(0.0, initial_value=initial_value)

    def __repr__(self):
        return "Positive()"


def describe(constraint):
    """(kind, lb, ub) of a constraint object - ours or GPyTorch's - for the C-ABI table."""
    from ._lib import CON_INTERVAL, CON_NONE, CON_SOFTPLUS
    if constraint is None:
        return CON_NONE, 0.0, 0.0
    names = [c.__name__ for c in type(constraint).__mro__]
    lb = float(torch.as_tensor(constraint.lower_bound).reshape(-1)[0])
    ub = float(torch.as_tensor(constraint.upper_bound).reshape(-1)[0])
    if "GreaterThan" in names:       # includes Positive
        return CON_SOFTPLUS, lb, 0.0
    if "Interval" in names:
        if "LessThan" in names:
            raise NotImplementedError("LessThan constraints are not supported by the engine")
        return CON_INTERVAL, lb, ub
    raise NotImplementedError(f"unsupported constraint type {type(constraint).__name__}")
