"""glass_b200.shells -- only the ``RadialWindow`` container of ``glass/shells.py:184`` (the
shape consumed by ``MultiPlaneConvergence.add_window`` and ``redshifts``); window
construction stays with upstream GLASS (out of scope)."""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any


@dataclass(frozen=True)
class RadialWindow:
    """glass/shells.py:184-260: redshift abscissae ``za``, weights ``wa``, effective ``zeff``."""

    za: Any
    wa: Any
    zeff: float = math.nan
