"""glass_b200.harmonics -- mirror of ``glass/harmonics.py``."""

from __future__ import annotations

import numpy as np
import torch


def multalm(alm, bl):
    """
    Multiply alm by bl (glass/harmonics.py:17-47).  alm in GLASS (l-major) order:
    entry ``l(l+1)/2 + m`` is scaled by ``bl[l]``.
    """
    n = bl.shape[0]
    if isinstance(alm, torch.Tensor):
        bl = torch.as_tensor(bl, device=alm.device)
        return alm * torch.repeat_interleave(bl, torch.arange(1, n + 1, device=alm.device))
    return alm * np.repeat(bl, np.arange(n) + 1)
