"""
Observations: the visibility map of glass/observations.py.  The n(z) models and
tomographic binning helpers of that module are plain array arithmetic outside the hot path and stay
with upstream GLASS.
"""

from __future__ import annotations

import torch

from . import healpix as hp


def vmap_galactic_ecliptic(nside: int, galactic=(30, 90), ecliptic=(20, 80), *, xp=None):
    """
    Visibility map masking galactic and ecliptic plane (glass/observations.py:51-101): a strip in
    galactic coordinates is blocked out, the map rotated to equatorial coordinates, a strip blocked
    out there, and the map rotated to ecliptic coordinates -- the reference's four calls, each one
    kernel here (``glb_query_strip``, ``glb_rotate_map_pixel``); the maps stay in HBM in between.
    NumPy array out, or a CUDA tensor with ``xp=torch``.
    """
    if len(galactic) != 2:
        msg = "galactic stripe must be a pair of numbers"
        raise TypeError(msg)
    if len(ecliptic) != 2:
        msg = "ecliptic stripe must be a pair of numbers"
        raise TypeError(msg)
    m = 1 - hp.query_strip(nside, galactic, dtype=torch.float64, xp=torch)  # ones * (1 - strip)
    m = hp.Rotator(coord="GC").rotate_map_pixel(m)
    m *= 1 - hp.query_strip(nside, ecliptic, dtype=torch.float64, xp=torch)
    m = hp.Rotator(coord="CE").rotate_map_pixel(m)
    return m if xp is torch else m.cpu().numpy()
