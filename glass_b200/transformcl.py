"""
glass_b200.transformcl -- C_l <-> C(theta) on the GPU, with the interface of the third-party
``transformcl`` package that GLASS calls (``cltocorr``, ``corrtocl``, ``cltovar``, ``theta``:
glass/grf/_solver.py:11,100-130, glass/grf/_core.py:179, glass/fields.py:890).

transformcl and its backend ``flt`` are un-vendored dependencies of the reference and are not
installed here, so this restates their published definition: the correlation function lives on
the n open nodes ``theta_k = pi (k + 1/2) / n`` and

    cltocorr:  C(theta_k) = sum_l (2l+1)/(4 pi) C_l P_l(cos theta_k)          (flt.idlt)
    corrtocl:  the exact inverse on those nodes                                (flt.dlt)

flt evaluates the pair one spectrum at a time with DCTs and a Chebyshev-Legendre recurrence.  Here
both directions are dense FP64 matrices built once per length on the device,

    P[k, l]    = P_l(cos theta_k)                        three-term recurrence along l
    Pinv       = L @ D,   D = DCT-II as a matrix (samples -> Chebyshev coefficients),
                          L = Chebyshev -> Legendre connection (closed form, upper triangular)

so that transforming ALL spectra of a simulation (S(S+1)/2 columns) is ONE cuBLAS DGEMM per
direction -- a plain library GEMM on the B200's FP64 units: n = 3 (lmax+1) = 24576 for the padded
solver at lmax 8191 is 2 n^2 S = 2.2e12 flop for 1830 spectra, ~60 ms.  The matrices take
8 n^2 bytes each (4.8 GB at that size) and are cached per (n, device).

Parity: unpinned against transformcl/flt themselves (absent offline); pinned mathematically --
``corrtocl(cltocorr(x)) = x``, Gauss-Legendre quadrature of the projection integrals, the
closed-form pairs used by the reference's own tests.
"""

from __future__ import annotations

import math

import numpy as np
import torch

from . import healpix as hp

_BLOCK = 2048  # rows per block while building the n x n tables (bounds the temporaries)


def _compute_device(*arrays) -> tuple[torch.device, bool]:
    """(device, on_device) -- a CUDA device always; raises when there is none (no CPU fallback)."""
    for a in arrays:
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.device, True
    return torch.device("cuda", hp._device_index()), False


def theta(n: int, *, xp=None):
    """The n nodes ``pi (k + 1/2) / n`` of the transform pair (flt.theta / transformcl.theta)."""
    t = (np.arange(n) + 0.5) * (math.pi / n)
    return torch.as_tensor(t) if xp is torch else t


def _lambda_tables(n: int, device) -> tuple[torch.Tensor, torch.Tensor]:
    """Lambda(z) = Gamma(z + 1/2) / Gamma(z + 1) at z = 0, 1, ... and z = 1/2, 3/2, ... from
    Lambda(z + 1) = Lambda(z) (z + 1/2) / (z + 1): relative error ~ sqrt(n) ulp, where a difference
    of lgamma values would lose 11 digits at z ~ 1e4."""
    k = torch.arange(n + 1, dtype=torch.float64, device=device)
    ints = math.sqrt(math.pi) * torch.cumprod(torch.cat([k.new_ones(1), (k[:-1] + 0.5) / (k[:-1] + 1.0)]), 0)
    half = (2.0 / math.sqrt(math.pi)) * torch.cumprod(torch.cat([k.new_ones(1), (k[:-1] + 1.0) / (k[:-1] + 1.5)]), 0)
    return ints, half  # ints[k] = Lambda(k), half[k] = Lambda(k + 1/2)


class _Tables:
    """P and Pinv for one length on one device."""

    def __init__(self, n: int, device):
        self.n, self.device = n, device
        f64 = torch.float64
        k = torch.arange(n, dtype=torch.int64, device=device)
        # cos(theta_k) with the argument reduced exactly: theta_k = pi (2k+1) / (2n)
        x = torch.cos((2 * k + 1).to(f64) * (math.pi / (2 * n)))
        # ---- P[k, l] by the recurrence (l+1) P_{l+1} = (2l+1) x P_l - l P_{l-1}, stored [l, k] first
        Pt = torch.empty((n, n), dtype=f64, device=device)
        Pt[0] = 1.0
        if n > 1:
            Pt[1] = x
        for l in range(1, n - 1):
            torch.mul(x, Pt[l], out=Pt[l + 1])
            Pt[l + 1].mul_((2 * l + 1) / (l + 1)).sub_(Pt[l - 1], alpha=l / (l + 1))
        self.P = Pt.t().contiguous()
        del Pt
        # ---- Pinv = L @ D in row blocks of L
        lam_int, lam_half = _lambda_tables(n, device)
        Pinv = torch.empty((n, n), dtype=f64, device=device)
        j = k.view(1, n)
        for i0 in range(0, n, _BLOCK):
            i = k[i0 : i0 + _BLOCK].view(-1, 1)
            # L[i, j], j >= i, i + j even:  sqrt(pi) / (2 Lambda(i)) on the diagonal (1 at i = 0),
            # -j (i + 1/2) / ((j + i + 1)(j - i)) Lambda((j - i - 2)/2) Lambda((j + i - 1)/2) above it
            upper = (j > i) & (((i + j) & 1) == 0)
            a = torch.where(upper, (j - i - 2) // 2, torch.zeros_like(j))
            b = torch.where(upper, (j + i - 2) // 2, torch.zeros_like(j))  # (j+i-1)/2 = b + 1/2
            jf, fi = j.to(f64), i.to(f64)
            Lblk = torch.where(upper, -jf * (fi + 0.5) / ((jf + fi + 1.0) * torch.clamp(jf - fi, min=1.0)) * lam_int[a] * lam_half[b], jf.new_zeros(()))
            diag = torch.where(i == 0, fi.new_ones(()), math.sqrt(math.pi) / (2.0 * lam_int[i.clamp(max=n)]))
            Lblk = torch.where(j == i, diag, Lblk)
            # accumulate over column blocks of L = row blocks of D, D[j, k] = (2 - [j=0]) / n cos(j theta_k)
            acc = torch.zeros((Lblk.shape[0], n), dtype=f64, device=device)
            for j0 in range(i0, n, _BLOCK):  # L is upper triangular: columns before i0 are zero
                jj = k[j0 : j0 + _BLOCK].view(-1, 1)
                r = (jj * (2 * k.view(1, n) + 1)) % (4 * n)  # j theta_k = pi r / (2n), exact in int64
                D = torch.cos(r.to(f64) * (math.pi / (2 * n))) * (2.0 / n)
                if j0 == 0:
                    D[0] *= 0.5
                acc.addmm_(Lblk[:, j0 : j0 + _BLOCK], D)
            Pinv[i0 : i0 + _BLOCK] = acc
        self.Pinv = Pinv


_CACHE: dict[tuple[int, str], _Tables] = {}
_CACHE_MAX = 2  # lengths kept (a solver run uses two: n and n + pad)


def _tables(n: int, device) -> _Tables:
    key = (int(n), str(device))
    t = _CACHE.get(key)
    if t is None:
        while len(_CACHE) >= _CACHE_MAX:
            _CACHE.pop(next(iter(_CACHE)))
        t = _CACHE[key] = _Tables(int(n), device)
    return t


def clear_tables() -> None:
    _CACHE.clear()


def _factors(n: int, device) -> torch.Tensor:
    return (2.0 * torch.arange(n, dtype=torch.float64, device=device) + 1.0) / (4.0 * math.pi)


def cltocorr_dev(cl: torch.Tensor) -> torch.Tensor:
    """[n] or [n, S] device tensor of spectra -> correlation functions on ``theta(n)`` (one DGEMM)."""
    n = cl.shape[0]
    if n == 0:
        return cl.clone()
    t = _tables(n, cl.device)
    if cl.ndim == 1:  # as a one-column matrix: the same GEMM, hence the same bits, as a batch of one
        return (t.P @ (cl * _factors(n, cl.device))[:, None])[:, 0]
    return t.P @ (cl * _factors(n, cl.device)[:, None])


def corrtocl_dev(corr: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`cltocorr_dev`."""
    n = corr.shape[0]
    if n == 0:
        return corr.clone()
    t = _tables(n, corr.device)
    f = _factors(n, corr.device)
    if corr.ndim == 1:
        return (t.Pinv @ corr[:, None])[:, 0] / f
    return (t.Pinv @ corr) / f[:, None]


def _wrap(fn, x):
    device, on_device = _compute_device(x)
    if isinstance(x, torch.Tensor):
        xd = x.to(device=device, dtype=torch.float64)
    else:
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device)
    out = fn(xd)
    return out if on_device else out.cpu().numpy()


def cltocorr(cl, closed: bool = False):
    """transformcl.cltocorr: angular power spectrum -> angular correlation function on
    ``theta(len(cl))``.  A trailing axis may hold several spectra (columns)."""
    if closed:
        raise NotImplementedError("closed=True (DCT-I nodes) is not used by GLASS and not built")
    return _wrap(cltocorr_dev, cl)


def corrtocl(corr, closed: bool = False):
    """transformcl.corrtocl: the inverse of :func:`cltocorr` on the same nodes."""
    if closed:
        raise NotImplementedError("closed=True (DCT-I nodes) is not used by GLASS and not built")
    return _wrap(corrtocl_dev, corr)


def cltovar(cl) -> float:
    """transformcl.cltovar: sum_l (2l+1)/(4 pi) C_l (host arithmetic on a tiny array, like
    glass_b200.fields.cltovar)."""
    cl = cl.detach().cpu().numpy() if isinstance(cl, torch.Tensor) else np.asarray(cl)
    ell = np.arange(cl.shape[0])
    return float(np.sum((2 * ell + 1) / (4 * np.pi) * cl))
