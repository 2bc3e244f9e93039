"""
glass_b200.rng -- what ``rng=`` may be (mirror of ``glass/rng.py``).

The reference wraps a sequential NumPy PCG64 stream (glass/rng.py:58-247); ``rng=None``
means a fresh seed-42 generator per call site (glass/rng.py:94-104).  A sequential stream
cannot be reproduced by a parallel sampler, so here every draw is a pure function of
``(seed, stream id, element index)`` through Philox4x32-10 on the device, and ``rng`` only
supplies the 64-bit seed:

* ``None``                   -> ``SEED`` (42), i.e. deterministic per call site like the reference
* ``int``                    -> that seed
* ``numpy.random.Generator`` -> one 63-bit integer drawn from it (advances the generator,
                                so successive calls get different, reproducible seeds)
* :class:`Deviates`          -> parity mode: explicit deviates supplied by the caller
                                (what the tests use for bit-exact comparisons)
"""

from __future__ import annotations

import numpy as np

SEED = 42

# Philox stream ids of the per-galaxy samplers.  A draw is a pure function of (seed, stream,
# element index): the stream says WHICH sampler, never how many calls were made before, so a
# given seed gives the same galaxies on any rank, in any order of calls (with ``rng=None`` the
# reference, too, returns the same draw on every call; pass a ``numpy.random.Generator`` to get
# a fresh seed per call).
STREAM_REDSHIFTS = 0x7A5F6E7A  # redshifts_from_nz
STREAM_PHZ = 0x70687A5F  # gaussian_phz
STREAM_ELLIPTICITY = 0x65707331  # ellipticity_gaussian / ellipticity_intnorm


class Deviates:
    """Explicit deviates for parity tests.

    ``normal_alm``: list (one per shell) of complex arrays in GLASS order, exactly the
    ``z`` of glass/fields.py:407.  ``poisson``: list of int64 count maps (one per
    population).  ``uv``: callable n -> (u, v) or tuple of arrays for in-pixel offsets.
    """

    def __init__(self, normal_alm=None, poisson=None, uv=None, uniform=None, normal=None):
        self.normal_alm = list(normal_alm) if normal_alm is not None else None
        self.poisson = list(poisson) if poisson is not None else None
        self.uv = uv
        self.uniform = uniform
        self.normal = normal
        self._i_alm = 0
        self._i_poisson = 0

    def next_normal_alm(self):
        z = self.normal_alm[self._i_alm]
        self._i_alm += 1
        return z

    def next_poisson(self):
        n = self.poisson[self._i_poisson]
        self._i_poisson += 1
        return n


def default_rng(seed: int = SEED):
    """glass/rng.py:21-55 (NumPy backend)."""
    return np.random.default_rng(seed)


def seed_from(rng) -> int:
    if rng is None:
        return SEED
    if isinstance(rng, (int, np.integer)):
        return int(rng) & 0xFFFFFFFFFFFFFFFF
    if isinstance(rng, np.random.Generator):
        return int(rng.integers(0, 2**63 - 1))
    if isinstance(rng, Deviates):
        return SEED
    raise TypeError(f"unsupported rng: {type(rng).__name__}")
