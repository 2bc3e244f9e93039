"""
glass_b200.healpix -- B200-native mirror of the seam ``glass/healpix.py``.

Same names and argument meaning as the reference wrappers, but instead of
round-tripping through host NumPy into healpy / healpix (``@numpy_fallback``,
glass/_array_api_utils.py:569) the work runs in the sm_100a kernels of
libglassb200.so.  Array rule: NumPy in -> NumPy out (host buffers, copies inside
the call); torch CUDA tensors in -> torch CUDA tensors out (no host traffic).
"""

from __future__ import annotations

import ctypes as C
import math
import os
from typing import Sequence

import numpy as np
import torch

from . import _lib

_PLANS: dict[tuple[int, int, int, int], "Plan"] = {}


def nside2npix(nside: int) -> int:
    """glass/healpix.py:296."""
    return 12 * int(nside) * int(nside)


def npix2nside(npix: int) -> int:
    """glass/healpix.py:279 (raises ValueError for an invalid pixel count)."""
    nside = math.isqrt(int(npix) // 12)
    if nside < 1 or 12 * nside * nside != int(npix):
        raise ValueError(f"invalid npix: {npix}")
    return nside


def get_nside(m) -> int:
    """glass/healpix.py:215."""
    return npix2nside(m.shape[-1])


def alm_getlmax(size: int) -> int:
    lmax = (math.isqrt(8 * int(size) + 1) - 3) // 2
    if (lmax + 1) * (lmax + 2) // 2 != int(size):
        raise ValueError(f"invalid alm size: {size}")
    return lmax


def _device_index(device=None) -> int:
    if not torch.cuda.is_available():
        raise _lib.GlassB200Error("glass_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.cuda.current_device()
    return torch.device(device).index or 0


LEGENDRE_MODES = {"auto": 0, "fp64": 1, "int8": 2}


class Plan:
    """Owner of a ``glb_plan`` (ring tables, twiddles, chirp spectra, workspace)."""

    def __init__(self, nside: int, lmax: int, max_batch: int = 1, device=None):
        self.lib = _lib.load()
        self.nside, self.lmax, self.max_batch = int(nside), int(lmax), int(max_batch)
        self.device = _device_index(device)
        self.npix = nside2npix(nside)
        self.nalm = (lmax + 1) * (lmax + 2) // 2
        self.nring = 4 * self.nside - 1
        h = C.c_void_p()
        _lib.check(self.lib.glb_plan_create(C.byref(h), self.nside, self.lmax, self.max_batch, self.device), "glb_plan_create")
        self.handle = h
        mode = os.environ.get("GLB_LEGENDRE", "auto").lower()
        if mode not in LEGENDRE_MODES:
            raise ValueError(f"GLB_LEGENDRE must be one of {sorted(LEGENDRE_MODES)}")
        self.set_legendre_mode(mode)

    def set_legendre_mode(self, mode: str) -> None:
        """Where the contraction over l of the scalar synthesis runs: "auto" (groups of eight maps on the
        INT8 tensor cores at nside >= 1024, the FP64 pipe otherwise), "fp64", or "int8" (groups of four and
        eight maps at any nside).  See ``glb_plan_set_legendre_mode``."""
        _lib.check(self.lib.glb_plan_set_legendre_mode(self.handle, LEGENDRE_MODES[mode]), "glb_plan_set_legendre_mode")
        self.legendre_mode = mode

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.glb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def torch_device(self):
        return torch.device("cuda", self.device)

    def stream_ptr(self) -> int:
        return torch.cuda.current_stream(self.torch_device).cuda_stream


def get_plan(nside: int, lmax: int, max_batch: int = 1, device=None) -> Plan:
    dev = _device_index(device)
    key = (int(nside), int(lmax), dev)
    for (ns, lm, d, mb), pl in _PLANS.items():
        if (ns, lm, d) == key and mb >= max_batch:
            return pl
    pl = Plan(nside, lmax, max_batch, dev)
    _PLANS[(int(nside), int(lmax), dev, int(max_batch))] = pl
    return pl


def clear_plans() -> None:
    _PLANS.clear()


def release_scratch() -> None:
    """Give back what the cached plans allocate on first use and rebuild on demand (the tile blocks of the
    INT8 Legendre path): for a caller that is done generating fields and needs the memory for maps."""
    for pl in _PLANS.values():
        _lib.check(pl.lib.glb_plan_release_scratch(pl.handle), "glb_plan_release_scratch")


def _as_cuda_c128(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.complex128).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.complex128)).to(device)


def _transform_args(transforms):
    """[(kind, p0, p1), ...] -> ctypes arrays (or NULLs)."""
    if transforms is None:
        return None, None, None
    n = len(transforms)
    kinds = (C.c_int * n)(*[int(t[0]) for t in transforms])
    params = (C.c_double * (2 * n))(*[float(v) for t in transforms for v in (t[1], t[2])])
    return kinds, params, (kinds, params)


def alm2map_batch(alms: torch.Tensor, nside: int, lmax: int | None = None, transforms=None, out=None) -> torch.Tensor:
    """
    Batched scalar synthesis on device: alms [nmaps, nalm] complex128 CUDA ->
    maps [nmaps, npix] float64 CUDA, optional fused per-map pixel transform
    ``transforms = [(kind, p0, p1), ...]`` (see include/glass_b200.h).
    """
    if alms.dim() != 2 or not alms.is_cuda or alms.dtype != torch.complex128:
        raise ValueError("alms must be a [nmaps, nalm] complex128 CUDA tensor")
    alms = alms.contiguous()
    nmaps = alms.shape[0]
    lmax = alm_getlmax(alms.shape[1]) if lmax is None else int(lmax)
    if (lmax + 1) * (lmax + 2) // 2 != alms.shape[1]:
        raise ValueError("alm size does not match lmax (mmax == lmax is required)")
    pl = get_plan(nside, lmax, max_batch=min(4, max(1, nmaps)), device=alms.device)
    if out is None:
        out = torch.empty((nmaps, pl.npix), dtype=torch.float64, device=alms.device)
    kinds, params, _keep = _transform_args(transforms)
    with torch.cuda.device(alms.device):
        rc = pl.lib.glb_alm2map(pl.handle, alms.data_ptr(), nmaps, out.data_ptr(), kinds, params, pl.stream_ptr())
    _lib.check(rc, "glb_alm2map")
    return out


def alm2map(
    alms,
    nside: int,
    *,
    inplace: bool = False,
    lmax: int | None = None,
    pixwin: bool = False,
    pol: bool = True,
):
    """
    Computes a HEALPix map given the alm (glass/healpix.py:38-78).

    Scalar transforms only: a single alm array, or a sequence of alm arrays with
    ``pol=False`` (each transformed independently, as healpy does).
    """
    single = not isinstance(alms, (list, tuple)) and getattr(alms, "ndim", 1) == 1
    if pixwin:  # healpy smooths the alm with the pixel window of nside before the synthesis
        size = (alms if single else alms[0]).shape[-1]
        lm = alm_getlmax(size) if lmax is None else lmax
        pw = globals()["pixwin"](nside, lmax=min(lm, 4 * nside))
        pw = np.concatenate([pw, np.zeros(lm + 1 - pw.size)])
        alms = almxfl(alms, pw) if single else [almxfl(a, pw) for a in alms]
    if not single and pol:
        raise NotImplementedError("polarised (TEB) alm2map is outside the GLASS hot path; pass pol=False")
    host = not isinstance(alms if single else alms[0], torch.Tensor)
    dev = torch.device("cuda", _device_index()) if host else (alms if single else alms[0]).device
    seq = [alms] if single else list(alms)
    stack = torch.stack([_as_cuda_c128(a, dev) for a in seq])
    maps = alm2map_batch(stack, nside, lmax)
    if host:
        res = maps.cpu().numpy()
        return res[0] if single else [res[i] for i in range(len(seq))]
    return maps[0] if single else [maps[i] for i in range(len(seq))]


# --------------------------------------------------------------------------------------
# pixel <-> angle, alm scaling
# --------------------------------------------------------------------------------------


def _dev_and_kind(*arrays):
    for a in arrays:
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.device, True
    return torch.device("cuda", _device_index()), False


def _to(x, device, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x)).to(device=device, dtype=dtype).contiguous()


def _out(t, on_device):
    return t if on_device else t.cpu().numpy()


def pixwin(nside: int, *, lmax: int | None = None, pol: bool = False, xp=None):
    """
    Return the pixel window function for the given nside (glass/healpix.py:313-356).

    healpy reads it from data files; here it is generated from its definition on the device
    (:mod:`glass_b200.pixwin`, cached per nside).  ``pol=True`` returns ``(w^T, w^P)``.  NumPy
    arrays, or CUDA tensors with ``xp=torch``.
    """
    from . import pixwin as _pw

    out = _pw.pixwin(nside, lmax=lmax, pol=pol)
    if xp is torch:
        dev = torch.device("cuda", _device_index())
        conv = lambda a: torch.as_tensor(a, device=dev)  # noqa: E731
        return tuple(conv(a) for a in out) if pol else conv(out)
    return out


def ring2ang_uv(nside: int, ipix, u, v, *, lonlat: bool = False):
    """Position inside ring pixel ``ipix`` at in-pixel offsets (u, v) in [0,1)^2: the
    kernel behind ``healpix.randang`` (glass/healpix.py:426-431), (u, v) = (1/2, 1/2) is
    the pixel centre."""
    dev, on_device = _dev_and_kind(ipix, u, v)
    ip = _to(ipix, dev, torch.int64)
    shape = ip.shape
    ip = ip.reshape(-1)
    uu = _to(u, dev, torch.float64).expand(shape).contiguous().reshape(-1)
    vv = _to(v, dev, torch.float64).expand(shape).contiguous().reshape(-1)
    o1 = torch.empty(ip.numel(), dtype=torch.float64, device=dev)
    o2 = torch.empty_like(o1)
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.glb_ring2ang_uv(int(nside), ip.data_ptr(), uu.data_ptr(), vv.data_ptr(), ip.numel(), int(lonlat), o1.data_ptr(), o2.data_ptr(), st), "glb_ring2ang_uv")
    return _out(o1.reshape(shape), on_device), _out(o2.reshape(shape), on_device)


def randang(nside: int, ipix, *, lonlat: bool = False):
    """Sample random spherical coordinates from the given HEALPix pixels
    (glass/healpix.py:398-432).  Like the reference, every call uses the same seed-42
    stream (glass/healpix.py:430)."""
    dev, on_device = _dev_and_kind(ipix)
    ip = _to(ipix, dev, torch.int64)
    shape = ip.shape
    ip = ip.reshape(-1)
    o1 = torch.empty(ip.numel(), dtype=torch.float64, device=dev)
    o2 = torch.empty_like(o1)
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.glb_randang(int(nside), ip.data_ptr(), ip.numel(), C.c_uint64(42), C.c_uint32(0), int(lonlat), o1.data_ptr(), o2.data_ptr(), st), "glb_randang")
    return _out(o1.reshape(shape), on_device), _out(o2.reshape(shape), on_device)


def ang2pix(nside: int, theta, phi, *, lonlat: bool = False):
    """Angles to RING pixel indices (glass/healpix.py:144-178)."""
    dev, on_device = _dev_and_kind(theta, phi)
    a = _to(theta, dev, torch.float64)
    b = _to(phi, dev, torch.float64)
    a, b = torch.broadcast_tensors(a, b)
    shape = a.shape
    a, b = a.contiguous().reshape(-1), b.contiguous().reshape(-1)
    out = torch.empty(a.numel(), dtype=torch.int64, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.glb_ang2pix(int(nside), a.data_ptr(), b.data_ptr(), a.numel(), int(lonlat), out.data_ptr(), st), "glb_ang2pix")
    return _out(out.reshape(shape), on_device)


def query_strip(nside: int, thetas, *, dtype=None, xp=None):
    """
    Mask of the pixels whose centres lie within the colatitude range ``thetas`` (radians)
    (glass/healpix.py:359-396 -> healpy.query_strip, RING, not inclusive).  ``thetas[0] >= thetas[1]``
    selects the complement, as in HEALPix.  int64 unless ``dtype`` says otherwise; NumPy array, or a
    CUDA tensor with ``xp=torch``.
    """
    theta1, theta2 = (float(t) for t in thetas)
    dev = torch.device("cuda", _device_index())
    out = torch.empty(nside2npix(nside), dtype=torch.float64, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.glb_query_strip(int(nside), theta1, theta2, out.data_ptr(), st), "glb_query_strip")
    if xp is torch:
        tdt = dtype if isinstance(dtype, torch.dtype) else (torch.int64 if dtype is None else getattr(torch, np.dtype(dtype).name))
        return out.to(tdt)
    return out.cpu().numpy().astype(np.int64 if dtype is None else dtype)


_OBLIQUITY_J2000 = (23.452294 - 0.0130125 - 1.63889e-6 + 5.02778e-7) * np.pi / 180.0
# ecliptic -> galactic, the constant matrix of HEALPix / healpy.rotator.get_coordconv_matrix
_E2G = np.array(
    [
        [-0.054882486, -0.993821033, -0.096476249],
        [0.494116468, -0.110993846, 0.862281440],
        [-0.867661702, -0.000346354, 0.497154957],
    ]
)


def _coordconv_matrix(coord) -> np.ndarray:
    """Matrix taking directions from system coord[0] to coord[1]; G galactic, E ecliptic,
    C (or Q) equatorial J2000 -- healpy.rotator.get_coordconv_matrix."""
    if coord is None:
        return np.identity(3)
    if isinstance(coord, str):
        coord = tuple(coord)
    names = [str(c).upper()[:1].replace("Q", "C") for c in coord]
    if len(names) == 1:
        names = names * 2
    if len(names) != 2 or any(c not in "GEC" for c in names):
        msg = "Wrong coord: must be a sequence of one or two of 'G', 'E', 'C'"
        raise TypeError(msg)
    a, b = names
    if a == b:
        return np.identity(3)
    ce, se = np.cos(_OBLIQUITY_J2000), np.sin(_OBLIQUITY_J2000)
    e2q = np.array([[1.0, 0.0, 0.0], [0.0, ce, -se], [0.0, se, ce]])
    g2e, q2e = np.linalg.inv(_E2G), np.linalg.inv(e2q)
    return {"EG": _E2G, "GE": g2e, "EC": e2q, "CE": q2e, "GC": e2q @ g2e, "CG": _E2G @ q2e}[a + b]


class Rotator:
    """Rotation operator between astronomical coordinate systems (glass/healpix.py:435-471)."""

    def __init__(self, *, coord=None) -> None:
        self.coord = coord
        self._matrix = _coordconv_matrix(coord)

    def rotate_map_pixel(self, m):
        """
        Rotate a HEALPix map to the new reference frame in pixel space (glass/healpix.py:457-471 ->
        healpy.Rotator.rotate_map_pixel): every pixel centre of the output is rotated BACK into the
        input frame, where the input map is interpolated bilinearly between the four nearest pixels
        of the two neighbouring rings.  One scalar map (what GLASS passes); NumPy in -> NumPy out,
        CUDA tensor in -> CUDA tensor out.
        """
        dev, on_device = _dev_and_kind(m)
        src = _to(m, dev, torch.float64)
        if src.ndim != 1:
            msg = "rotate_map_pixel takes one scalar map here (polarised triplets are not on the GLASS path)"
            raise NotImplementedError(msg)
        nside = npix2nside(src.numel())
        out = torch.empty_like(src)
        inv = np.ascontiguousarray(np.linalg.inv(self._matrix), dtype=np.float64)
        rot9 = (C.c_double * 9)(*inv.reshape(-1))
        lib = _lib.load()
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.glb_rotate_map_pixel(int(nside), rot9, src.data_ptr(), out.data_ptr(), st), "glb_rotate_map_pixel")
        return _out(out, on_device)


def almxfl(alm, fl, *, inplace: bool = False):
    """Multiply alm by a function of l, zero where not defined (glass/healpix.py:111-140)."""
    dev, on_device = _dev_and_kind(alm)
    a = _to(alm, dev, torch.complex128)
    if not (inplace and on_device and a.data_ptr() == alm.data_ptr()):
        a = a.clone()
    f = _to(fl, dev, torch.float64)
    lmax = alm_getlmax(a.numel())
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.glb_almxfl(lmax, a.data_ptr(), f.data_ptr(), f.numel(), st), "glb_almxfl")
    if on_device:
        return a
    res = a.cpu().numpy()
    if inplace and isinstance(alm, np.ndarray):
        alm[...] = res
        return alm
    return res


def alm2map_spin_batch(alms: torch.Tensor, nside: int, spin: int, lmax: int, out=None):
    """E-only spin synthesis of several map pairs at once (extension; the reference shears one
    convergence plane per ``hp.alm2map_spin([alm, 0], nside, 2, lmax)`` call, glass/lensing.py:428).
    ``alms``: CUDA complex128 [nb, nalm], nb <= 4 per launch group (more are done in turn); returns
    ``(maps1, maps2)`` of shape [nb, npix].  The planes share the two Wigner-d recurrences."""
    alms = alms.contiguous()
    nb, dev = alms.shape[0], alms.device
    if (lmax + 1) * (lmax + 2) // 2 != alms.shape[1]:
        raise ValueError("alm size does not match lmax (mmax == lmax is required)")
    pl = get_plan(nside, lmax, max_batch=4 if nb >= 4 else (2 if nb >= 2 else 1), device=dev)
    m1, m2 = out if out is not None else (torch.empty((nb, pl.npix), dtype=torch.float64, device=dev), torch.empty((nb, pl.npix), dtype=torch.float64, device=dev))
    with torch.cuda.device(dev):
        for a in range(0, nb, 4):
            n = min(4, nb - a)
            rc = pl.lib.glb_alm2map_spin_batch(pl.handle, alms[a:].data_ptr(), n, int(spin), m1[a:].data_ptr(), m2[a:].data_ptr(), pl.stream_ptr())
            _lib.check(rc, "glb_alm2map_spin_batch")
    return m1, m2


def alm2map_spin(alms: Sequence, nside: int, spin: int, lmax: int):
    """
    Computes maps from a set of 2 spinned alm (glass/healpix.py:81-108):
    ``map1 + i map2 = sum -(alm1 + i alm2)_lm  sY_lm``.  ``alms[1]`` may be ``None`` or all
    zeros (what GLASS always passes, glass/lensing.py:334,411) for the E-only fast path.
    Returns a list of the 2 maps in RING scheme.
    """
    a1, a2 = alms[0], alms[1]
    dev, on_device = _dev_and_kind(a1, a2)
    d1 = _to(a1, dev, torch.complex128)
    d2 = None
    if a2 is not None:
        d2 = _to(a2, dev, torch.complex128)
        if not bool(torch.any(d2 != 0)):
            d2 = None
    if (lmax + 1) * (lmax + 2) // 2 != d1.numel():
        raise ValueError("alm size does not match lmax (mmax == lmax is required)")
    pl = get_plan(nside, lmax, max_batch=1, device=dev)
    m1 = torch.empty(pl.npix, dtype=torch.float64, device=dev)
    m2 = torch.empty(pl.npix, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = pl.lib.glb_alm2map_spin(pl.handle, d1.data_ptr(), None if d2 is None else d2.data_ptr(), int(spin), m1.data_ptr(), m2.data_ptr(), pl.stream_ptr())
    _lib.check(rc, "glb_alm2map_spin")
    return [_out(m1, on_device), _out(m2, on_device)]


def map2alm(maps, *, lmax: int | None = None, pol: bool = True, use_pixel_weights: bool = False, niter: int = 3, ring_weights=None):
    """
    Computes the alm of a HEALPix map in RING ordering (glass/healpix.py:232-276; GLASS calls
    it with ``pol=False, use_pixel_weights=True``, glass/lensing.py:306,408).

    healpy's pixel-weight tables are data files that cannot be obtained offline, so
    ``use_pixel_weights`` is accepted for signature compatibility and quadrature weights are
    passed explicitly instead: ``ring_weights`` ([4*nside-1], default uniform) and ``niter``
    Jacobi refinements ``alm += A(map - S(alm))`` (healpy's default ``iter=3``).
    """
    single = not isinstance(maps, (list, tuple)) and getattr(maps, "ndim", 1) == 1
    if not single and pol:
        raise NotImplementedError("polarised (TQU) map2alm is outside the GLASS hot path; pass pol=False")
    seq = [maps] if single else list(maps)
    dev, on_device = _dev_and_kind(*seq)
    ms = [_to(m, dev, torch.float64).reshape(-1) for m in seq]
    nside = npix2nside(ms[0].numel())
    if any(m.numel() != ms[0].numel() for m in ms):
        raise ValueError("all maps must have the same size")
    lmax = 3 * nside - 1 if lmax is None else int(lmax)
    nmaps = len(ms)
    pl = get_plan(nside, lmax, max_batch=4 if nmaps >= 4 else (2 if nmaps >= 2 else 1), device=dev)
    alms = torch.empty((nmaps, pl.nalm), dtype=torch.complex128, device=dev)
    w = None if ring_weights is None else _to(ring_weights, dev, torch.float64)
    if w is not None and w.numel() != 4 * nside - 1:
        raise ValueError("ring_weights must have 4*nside-1 entries")
    with torch.cuda.device(dev):
        done = 0
        while done < nmaps:  # groups of 4, 2, 1: the refinement syntheses of a group share one recurrence
            g = 4 if nmaps - done >= 4 else (2 if nmaps - done >= 2 else 1)
            grp = ms[done] if g == 1 else torch.stack(ms[done : done + g])
            grp = grp.contiguous()
            rc = pl.lib.glb_map2alm_batch(pl.handle, grp.data_ptr(), g, None if w is None else w.data_ptr(), int(niter),
                                          alms[done:].data_ptr(), pl.stream_ptr())
            _lib.check(rc, "glb_map2alm_batch")
            done += g
    if single:
        return _out(alms[0], on_device)
    return [_out(alms[b], on_device) for b in range(nmaps)]
