"""
glass_b200.shapes -- B200-native mirror of ``ellipticity_intnorm`` and
``ellipticity_gaussian`` (glass/shapes.py:288-362, 223-285): Philox normals and the
ellipticity formula in one kernel per population.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _arrays as A
from . import _lib
from . import healpix as hp
from . import rng as _rng



def _run(mode, count, sigma_fn, sigma, rng, xp):
    deviates = rng if isinstance(rng, _rng.Deviates) else None
    seed = _rng.seed_from(rng)
    on_device = xp is torch or A.is_cuda(count) or A.is_cuda(sigma)
    device = torch.device("cuda", hp._device_index())
    cnt, sig = np.broadcast_arrays(A.to_np(count), A.to_np(sigma))
    total = int(np.sum(cnt))
    out = torch.empty(total, dtype=torch.complex128, device=device)
    lib = _lib.load()
    pos = 0
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        for k in np.ndindex(*cnt.shape):
            n_k = int(cnt[k])
            if n_k == 0:
                continue
            normals = None
            if mode == 0 and deviates is not None and deviates.normal is not None:
                nn = deviates.normal(n_k) if callable(deviates.normal) else deviates.normal[pos : pos + n_k]
                normals = A.to_dev(nn, device, torch.complex128)
            _lib.check(
                lib.glb_ellipticity(
                    mode, float(sigma_fn(float(sig[k]))), None if normals is None else normals.data_ptr(), n_k,
                    C.c_uint64(seed), C.c_uint32(_rng.STREAM_ELLIPTICITY), C.c_uint64(pos), out[pos:].data_ptr(), st,
                ),
                "glb_ellipticity",
            )
            pos += n_k
    return out if on_device else out.cpu().numpy()


def ellipticity_gaussian(count, sigma, *, rng=None, xp=None):
    """Sample Gaussian galaxy ellipticities, re-drawing |e| > 1 (glass/shapes.py:223-285)."""
    return _run(1, count, lambda s: s, sigma, rng, xp)


def ellipticity_intnorm(count, sigma, *, rng=None, xp=None):
    """Sample galaxy ellipticities with intrinsic normal distribution
    (glass/shapes.py:288-362)."""
    sig = A.to_np(sigma)
    if not np.all((sig >= 0) & (sig < 0.5**0.5)):
        msg = "sigma must be between 0 and sqrt(0.5)"
        raise ValueError(msg)
    return _run(0, count, lambda s: s * ((8 + 5 * s**2) / (2 - 4 * s**2)) ** 0.5, sigma, rng, xp)
