"""
glass_b200.grf -- the per-pixel map transformations of ``glass/grf/_transformations.py``
(only ``__call__``: the C_l <-> C(theta) solver side of glass.grf is out of scope,
SURVEY.md section 8a row A7).  ``__call__`` works on NumPy arrays and torch tensors; in
``glass_b200.generate`` these transformations are recognised and fused into the ring-FFT
epilogue of the synthesis kernel instead of being applied as a separate pass.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


def _expm1(x):
    return torch.expm1(x) if isinstance(x, torch.Tensor) else np.expm1(x)


@dataclass
class Normal:
    """glass/grf/_transformations.py:14-27: t(X) = X."""

    def __call__(self, x, _var, /):
        return x

    def _fused(self, var):
        return (_lib.T_NORMAL, 0.0, 1.0)


@dataclass
class Lognormal:
    r"""glass/grf/_transformations.py:60-89: t(X) = lamda [exp(X - var/2) - 1]."""

    lamda: float = 1.0

    def __call__(self, x, var, /):
        x = _expm1(x - var / 2)
        if self.lamda != 1.0:
            x = self.lamda * x
        return x

    def _fused(self, var):
        return (_lib.T_LOGNORMAL, float(var) / 2, float(self.lamda))


@dataclass
class SquaredNormal:
    r"""glass/grf/_transformations.py:141-176: t(X) = lamda [(X - a)^2 - 1]."""

    a: float
    lamda: float = 1.0

    def __call__(self, x, _var, /):
        x = (x - self.a) ** 2 - 1
        if self.lamda != 1.0:
            x = self.lamda * x
        return x

    def _fused(self, var):
        return (_lib.T_SQUARED_NORMAL, float(self.a), float(self.lamda))


def fused_descriptor(t, var):
    """(kind, p0, p1) if ``t`` is one of the known transformations (ours or the
    reference's own dataclasses of the same name), else None."""
    if hasattr(t, "_fused"):
        return t._fused(var)
    name = type(t).__name__
    if name == "Normal" and not hasattr(t, "lamda"):
        return (_lib.T_NORMAL, 0.0, 1.0)
    if name == "Lognormal" and hasattr(t, "lamda"):
        return (_lib.T_LOGNORMAL, float(var) / 2, float(t.lamda))
    if name == "SquaredNormal" and hasattr(t, "lamda") and hasattr(t, "a"):
        return (_lib.T_SQUARED_NORMAL, float(t.a), float(t.lamda))
    return None
