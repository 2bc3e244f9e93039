"""
glass_b200.galaxies -- B200-native mirror of the hot-path part of ``glass/galaxies.py``:
``galaxy_shear``, ``redshifts`` / ``redshifts_from_nz``.

``galaxy_shear`` replaces the Python loop over 10 000-galaxy chunks (glass/galaxies.py:330-335)
by one kernel doing the pixel lookup, the three map gathers and the reduced-shear formula.
"""

from __future__ import annotations

import ctypes as C
import warnings

import numpy as np
import torch

from . import _arrays as A
from . import _lib
from . import healpix as hp
from . import rng as _rng

_CDF_TABLES: dict = {}  # (device, key of z, key of nz) -> (cdf, z) on the device


def _array_key(a):
    """Identity of an array's CONTENT without a device round trip: CUDA tensors by storage and
    version counter, host arrays by their bytes (n(z) tables are a few hundred numbers)."""
    if isinstance(a, torch.Tensor) and a.is_cuda:
        return ("cuda", a.data_ptr(), a._version, tuple(a.shape), tuple(a.stride()))
    h = np.ascontiguousarray(A.to_np(a), dtype=np.float64)
    return ("host", h.shape, h.tobytes())


def _cdf_table(z_k, nz_k, device):
    """Normalised cumulative trapezoid of n(z) and its grid as device arrays
    (glass/galaxies.py:77-89), cached: the sampler is called once per batch of galaxies with the
    same window, and building the table on the host costs a stream synchronisation (CUDA inputs)
    and two blocking pageable copies each time."""
    key = (str(device), _array_key(z_k), _array_key(nz_k))
    hit = _CDF_TABLES.get(key)
    if hit is None:
        zh, nzh = np.asarray(A.to_np(z_k), dtype=np.float64), np.asarray(A.to_np(nz_k), dtype=np.float64)
        zh, nzh = np.broadcast_arrays(zh, nzh)
        cdf = _cumulative_trapezoid(nzh, zh)
        cdf /= cdf[-1]
        if len(_CDF_TABLES) >= 256:
            _CDF_TABLES.clear()
        # keep the key's CUDA tensors alive with the entry so that data_ptr cannot be recycled
        hit = _CDF_TABLES[key] = (A.to_dev(cdf, device), A.to_dev(zh, device), int(cdf.shape[0]), (z_k, nz_k))
    return hit[:3]


def _cumulative_trapezoid(f, x):
    """glass/arraytools.py:197-226."""
    f = np.asarray(f, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    return np.concatenate([np.zeros(f.shape[:-1] + (1,)), np.cumsum((f[..., 1:] + f[..., :-1]) * 0.5 * np.diff(x), axis=-1)], axis=-1)


def redshifts(n, w, *, rng=None):
    """Sample redshifts from a radial window function (glass/galaxies.py:92-119)."""
    return redshifts_from_nz(n, w.za, w.wa, rng=rng, warn=False)


@A.nvtx("glass.redshifts_from_nz")
def redshifts_from_nz(count, z, nz, *, rng=None, warn: bool = True):
    """
    Generate galaxy redshifts from a source distribution (glass/galaxies.py:188-268):
    inverse-CDF sampling ``interp(U, cdf, z)`` per population (glass/galaxies.py:77-89).
    Returns a CUDA tensor when ``z``/``nz`` are CUDA tensors, else a NumPy array.
    """
    if warn:
        warnings.warn(
            "when sampling galaxies, redshifts_from_nz() is often not the function you"
            " want. Try redshifts() instead. Use warn=False to suppress this warning.",
            stacklevel=2,
        )
    device, on_device = A.pick_device(z, nz, count)
    deviates = rng if isinstance(rng, _rng.Deviates) else None
    seed = _rng.seed_from(rng)
    ch = A.to_np(count)
    z, nz = (a if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64) for a in (z, nz))
    zs, nzs = tuple(z.shape), tuple(nz.shape)
    dims = np.broadcast_shapes(ch.shape, zs[:-1], nzs[:-1])
    count_out = np.broadcast_to(ch, dims)
    total = int(np.sum(count_out))
    out = torch.empty(total, dtype=torch.float64, device=device)
    lib = _lib.load()
    pos = 0
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        for k in np.ndindex(*dims):
            n_k = int(count_out[k])
            if n_k == 0:
                continue
            d_cdf, d_z, ncdf = _cdf_table(A.take_leading(z, zs[:-1], dims, k), A.take_leading(nz, nzs[:-1], dims, k), device)
            u = None
            if deviates is not None and deviates.uniform is not None:
                u = A.to_dev(deviates.uniform(n_k) if callable(deviates.uniform) else deviates.uniform[pos : pos + n_k], device)
            _lib.check(
                lib.glb_redshifts_from_cdf(
                    d_cdf.data_ptr(), d_z.data_ptr(), ncdf, None if u is None else u.data_ptr(), n_k,
                    C.c_uint64(seed), C.c_uint32(_rng.STREAM_REDSHIFTS), C.c_uint64(pos), out[pos:].data_ptr(), st,
                ),
                "glb_redshifts_from_cdf",
            )
            pos += n_k
    return out if on_device else out.cpu().numpy()


def redshifts_from_bins(bins, z, nz_dict, *, rng=None):
    """
    Redshifts for galaxies with tomographic bin labels (glass/galaxies.py:122-185): every
    galaxy draws from the n(z) of its bin.  The labels are tallied (sorted unique labels and their
    counts), one run of inverse-CDF draws is made per label with the redshift kernel
    (:func:`redshifts_from_nz`), and the runs are scattered back to the galaxies' positions with
    one stable argsort -- the same order of draws as the reference, so supplied uniform deviates
    reproduce it bit for bit.  CUDA ``bins`` keep everything on the device.  N-D label arrays are
    handled in flattened (C) order; the reference's argsort pair works along the last axis only and
    is meaningful for 1-D labels.
    """
    keys, values = list(nz_dict.keys()), list(nz_dict.values())
    on_device = A.is_cuda(bins) or A.is_cuda(z)
    b = bins if isinstance(bins, torch.Tensor) else torch.as_tensor(np.asarray(bins))
    labels, inverse, counts = torch.unique(b.reshape(-1), sorted=True, return_inverse=True, return_counts=True)
    labels_h, counts_h = labels.cpu().numpy(), counts.cpu().numpy()
    runs = []
    for x, k in zip(labels_h, counts_h):
        idx = next(i for i, key in enumerate(keys) if np.all(A.to_np(key) == x))  # labels need not be hashable
        runs.append(redshifts_from_nz(int(k), z, values[idx], rng=rng, warn=False))
    if on_device:
        dev = next(t.device for t in (bins, z) if A.is_cuda(t))
        red = torch.cat([torch.as_tensor(r, device=dev) for r in runs]) if runs else torch.empty(0, dtype=torch.float64, device=dev)
        order = torch.argsort(inverse.to(dev), stable=True)  # positions of the galaxies, bin by bin
        out = torch.empty_like(red)
        out[order] = red
        return out.reshape(b.shape)
    red = np.concatenate([A.to_np(r) for r in runs]) if runs else np.empty(0)
    order = np.argsort(inverse.cpu().numpy(), kind="stable")
    out = np.empty_like(red)
    out[order] = red
    return out.reshape(tuple(b.shape))


@A.nvtx("glass.galaxy_shear")
def galaxy_shear(lon, lat, eps, kappa, gamma1, gamma2, *, reduced_shear: bool = True, ipix=None):
    """
    Observed galaxy shears from weak lensing (glass/galaxies.py:271-347).

    ``ipix`` (extension, optional): ring pixel index of every galaxy if already known
    (e.g. from the position sampler); skips the ang2pix lookup.
    """
    device, on_device = A.pick_device(lon, lat, eps, kappa, gamma1, gamma2)
    k = A.to_dev(kappa, device)
    g1 = A.to_dev(gamma1, device)
    g2 = A.to_dev(gamma2, device)
    npix = max(k.shape[-1], g1.shape[-1], g2.shape[-1])
    nside = hp.npix2nside(npix)
    k, g1, g2 = (t.expand(npix).contiguous() if t.numel() != npix else t.reshape(npix) for t in (k, g1, g2))
    lon_d, lat_d = A.to_dev(lon, device), A.to_dev(lat, device)
    eps_d = A.to_dev(eps, device, torch.complex128)
    lon_d, lat_d, eps_d = torch.broadcast_tensors(lon_d, lat_d, eps_d)
    lon_d, lat_d, eps_d = lon_d.contiguous().reshape(-1), lat_d.contiguous().reshape(-1), eps_d.contiguous().reshape(-1)
    n = eps_d.numel()
    out = torch.empty(n, dtype=torch.complex128, device=device)
    ip = None if ipix is None else A.to_dev(ipix, device, torch.int64)
    lib = _lib.load()
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        _lib.check(
            lib.glb_galaxy_shear(
                nside, lon_d.data_ptr(), lat_d.data_ptr(), None if ip is None else ip.data_ptr(), eps_d.data_ptr(), n,
                k.data_ptr(), g1.data_ptr(), g2.data_ptr(), int(bool(reduced_shear)), out.data_ptr(), st,
            ),
            "glb_galaxy_shear",
        )
    return out if on_device else out.cpu().numpy()


def gaussian_phz(z, sigma_0, *, lower=None, upper=None, rng=None, xp=None):
    r"""
    Photometric redshifts assuming a Gaussian error :math:`\sigma(z) = (1 + z) \sigma_0`
    (glass/galaxies.py:350-455), with plain rejection sampling outside ``[lower, upper]``.

    Returns an array of the broadcast shape of ``z`` and ``sigma_0`` (CUDA tensor if an input is
    one or ``xp is torch``, else NumPy).  Parity mode: ``rng=Deviates(normal=[n_0, n_1, ...])``
    supplies the standard normals of the reference's successive rounds (full-size arrays).
    """
    device, on_device = A.pick_device(z, sigma_0, lower, upper)
    on_device = on_device or (xp is torch)
    deviates = rng if isinstance(rng, _rng.Deviates) else None
    seed = _rng.seed_from(rng)

    def dev(a):  # keeps 0-d inputs 0-d (A.to_dev goes through ascontiguousarray, which makes them 1-d)
        if isinstance(a, torch.Tensor):
            return a.to(device=device, dtype=torch.float64)
        return torch.as_tensor(np.asarray(a, dtype=np.float64), device=device)

    z_d, s_d = dev(z), dev(sigma_0)
    dims = tuple(torch.broadcast_shapes(z_d.shape, s_d.shape))
    lo_d = dev(0.0 if lower is None else lower)
    hi_d = dev(float("inf") if upper is None else upper)
    if lower is None and upper is not None:
        lo_d = torch.zeros_like(hi_d)
    if upper is None and lower is not None:
        hi_d = torch.full_like(lo_d, float("inf"))
    if (lo_d.ndim == hi_d.ndim != 0) and not (tuple(lo_d.shape) == tuple(hi_d.shape) == dims):
        msg = "lower and upper must best scalars or have the same shape as z"
        raise ValueError(msg)
    if not bool(torch.all(lo_d < hi_d).item()):
        msg = "requires lower < upper"
        raise ValueError(msg)
    n = int(np.prod(dims)) if dims else 1
    z_f = z_d.expand(dims).contiguous().reshape(-1) if dims else z_d.reshape(1)

    def per_galaxy(t):  # (device array or None, scalar)
        if t.ndim == 0:
            return None, float(t.item())
        return t.expand(dims).contiguous().reshape(-1), 0.0

    s_arr, s_val = per_galaxy(s_d)
    lo_arr, lo_val = per_galaxy(lo_d)
    hi_arr, hi_val = per_galaxy(hi_d)
    out = torch.empty(n, dtype=torch.float64, device=device)
    lib = _lib.load()
    ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream

        def launch(normals, redraw_only, nbad):
            _lib.check(
                lib.glb_gaussian_phz(
                    z_f.data_ptr(), ptr(s_arr), s_val, ptr(lo_arr), lo_val, ptr(hi_arr), hi_val, ptr(normals), redraw_only, n,
                    C.c_uint64(seed), C.c_uint32(_rng.STREAM_PHZ), out.data_ptr(), ptr(nbad), st,
                ),
                "glb_gaussian_phz",
            )

        if deviates is not None and deviates.normal is not None:
            rounds = iter(deviates.normal)
            first = True
            while True:
                nbad = torch.zeros(1, dtype=torch.int64, device=device)
                launch(dev(next(rounds)).expand(dims).contiguous().reshape(-1) if dims else dev(next(rounds)).reshape(1), 0 if first else 1, nbad)
                first = False
                if int(nbad.item()) == 0:
                    break
        else:
            launch(None, 0, None)
    res = out.reshape(dims) if dims else out.reshape(())
    return res if on_device else res.cpu().numpy()
