"""
glass_b200.pixwin -- the HEALPix pixel window functions w_l^T, w_l^P generated numerically
(``glass/healpix.py:313-356`` returns healpy's tabulated ones; used at ``glass/fields.py:289``
and ``glass/lensing.py:361, 421``).

healpy reads ``pixel_window_n????.fits`` from its data package, which is not available offline;
here the windows are computed from their definition (HEALPix primer, "pixel window functions")

    w_l^2 = 1/N_pix  sum_p  4 pi/(2l+1)  sum_m | 1/Omega_p  int_p  sY_lm dOmega |^2 ,   s = 0 (T), 2 (P)

without any transform.  By the addition theorem of the (spin-weighted) harmonics the sum over m is
a function of the two points only,

    sum_m sY*_lm(u) sY_lm(u') = (2l+1)/(4 pi)  d^l_ss(gamma) e^{-i s (chi - chi')} ,

(gamma the angle between u and u'; chi, chi' the angles between their great circle and the local
meridians), so that

    w_l^2 = <  d^l_ss(gamma)  cos(s (chi - chi'))  >   over pairs of points of the same pixel, over pixels,

and both d^l_00 = P_l(cos gamma) and d^l_22 = cos^4(gamma/2) P^(0,4)_(l-2)(cos gamma) are
polynomials in x = sin^2(gamma/2) (terminating 2F1 series) whose coefficients depend on l alone.
Hence ALL l come from one set of pair moments  M_k = < weight * x^k >:

    w_l^T^2 = sum_k c_k(l) M_k^T ,     c_k(l) = prod_{j<=k} [ -(l+j)(l-j+1) / j^2 ]
    w_l^P^2 = sum_k e_k(l) M_k^P ,     e_k(l) = prod_{j<=k} [ -(l-1-j)(l+2+j) / j^2 ] ,  weight (1-x)^2 cos 2(chi-chi')

The pixel integrals are Gauss-Legendre quadratures in the pixel's own (u, v) face coordinates, which
map to the sphere with constant Jacobian (HEALPix is equal-area); positions come from the library's
pixel -> angle kernel (``glb_ring2ang_uv``), the pair sums are torch reductions on the device.
Pixels of a ring are congruent within a quadrant of a polar-cap ring (i shapes in ring i) and all
around an equatorial ring, north and south are mirror images: 4 N(N+1)/2 + N + 1 shapes for nside N.

Like the tabulated windows, which are exact up to nside 128 and extrapolated above, windows for
nside > 128 are taken from the nside-128 moments at the scaled multipole
(l' + 1/2) = (l + 1/2) 128 / nside -- the series are polynomials in l, so they are evaluated AT the
real l', not interpolated; what is neglected is the O(nside^-2) curvature of a pixel, < 1e-5.
w^P_0 = w^P_1 = 0 (there is no spin-2 harmonic below l = 2).

Parity with healpy's data files: unpinned (files absent here); pinned mathematically against the
brute-force definition (sum over m of pixel-averaged scipy harmonics) in the CPU suite.
"""

from __future__ import annotations

import functools

import numpy as np
import torch

EXACT_NSIDE_MAX = 128  # above: scaled from this nside, as the tabulated windows are
_Q = None  # Gauss-Legendre nodes per pixel axis; None: by nside (see _nodes)
_K = 64  # terms of the series in sin^2(gamma/2): enough for l gamma <= 14, i.e. l <= 4 nside


def _compute_device():
    """The device the pair sums run on.  No CPU fallback: the generator is part of the CUDA path
    (the CPU suite patches this function to run the same code on CPU tensors)."""
    from . import healpix as hp

    return torch.device("cuda", hp._device_index())


def _positions(nside: int, ipix: torch.Tensor, u: torch.Tensor, v: torch.Tensor):
    """(theta, phi) of in-pixel offsets (u, v) of ring pixels ``ipix`` -- the library's kernel."""
    from . import healpix as hp

    return hp.ring2ang_uv(nside, ipix, u, v)


def _nodes(nside: int) -> int:
    """Gauss-Legendre order per axis of each of the two triangles of :func:`_pixel_rule`."""
    if _Q is not None:
        return _Q
    return 12 if nside <= 8 else 8  # 8: converged to 1e-11 (nside 1, 2, 16 against 16 nodes), 12: to 1e-15


def _pixel_rule(q: int):
    """Nodes (u, v) and weights (sum 1) of the quadrature over the unit pixel square.

    The (u, v) -> sphere map is analytic inside a pixel except (a) across the polar-cap / belt
    transition, which cuts the pixels of ring nside along their diagonal u + v = 1, and (b) at a
    pole, which is the corner (1, 1) of the four northern and (0, 0) of the four southern polar
    pixels.  A tensor Gauss rule on the square converges only algebraically for those pixels
    (3e-5 at nside 2 with 12 x 12 nodes).  The square is therefore cut along u + v = 1 into two
    triangles, each parametrised from the unit square by the Duffy map that collapses at (0, 0)
    resp. (1, 1): the integrand is smooth on either triangle and the Jacobian s removes the corner
    singularity, so the rule converges spectrally for every pixel."""
    g, w = np.polynomial.legendre.leggauss(q)
    g, w = 0.5 * (g + 1.0), 0.5 * w
    s, t = np.meshgrid(g, g, indexing="ij")
    wt = np.outer(w, w) * s  # Jacobian of (s, t) -> (s (1 - t), s t); each triangle has area 1/2
    u = np.concatenate([(s * (1.0 - t)).ravel(), (1.0 - s * (1.0 - t)).ravel()])
    v = np.concatenate([(s * t).ravel(), (1.0 - s * t).ravel()])
    return u, v, np.concatenate([wt.ravel(), wt.ravel()])


def _shapes(nside: int):
    """One pixel per distinct pixel shape of the northern hemisphere incl. the equator, with the
    number of pixels of the sphere that are congruent to it."""
    pix, mult = [], []
    for i in range(1, nside):  # polar-cap rings: i shapes (one quadrant), 4 quadrants, 2 hemispheres
        start = 2 * i * (i - 1)
        pix.extend(range(start, start + i))
        mult.extend([8.0] * i)
    for i in range(nside, 2 * nside + 1):  # equatorial rings: all 4 nside pixels congruent
        pix.append(2 * nside * (nside - 1) + (i - nside) * 4 * nside)
        mult.append(4.0 * nside * (1.0 if i == 2 * nside else 2.0))
    return np.asarray(pix, dtype=np.int64), np.asarray(mult, dtype=np.float64)


@functools.lru_cache(maxsize=8)
def _pair_moments(nside: int):
    """(x0, M^T[k], M^P[k]), k = 0.._K: moments of x / x0, x = sin^2(gamma/2), over pairs of points of
    one pixel, averaged over all pixels of the sphere; x0 = the largest x met (keeps x^k in range)."""
    dev = _compute_device()
    pix, mult = _shapes(nside)
    uu, vv, wq = _pixel_rule(_nodes(nside))
    q2 = wq.size
    wq = torch.as_tensor(wq, device=dev)
    ww = wq[:, None] * wq[None, :]
    chunk = max(1, (1 << 25) // (q2 * q2))
    mt = torch.zeros(_K + 1, dtype=torch.float64, device=dev)
    mp = torch.zeros(_K + 1, dtype=torch.float64, device=dev)
    u_all = torch.as_tensor(uu, device=dev)
    v_all = torch.as_tensor(vv, device=dev)
    # largest chord of any pixel: the coarsest (polar) pixels; an upper bound keeps every ratio <= 1
    x0 = float(min(1.0, (2.6 / nside) ** 2 / 4.0)) if nside > 1 else 1.0
    for a in range(0, pix.size, chunk):
        p = torch.as_tensor(pix[a : a + chunk], device=dev)
        m = torch.as_tensor(mult[a : a + chunk], device=dev)
        n = p.numel()
        th, ph = _positions(nside, p[:, None].expand(n, q2), u_all[None, :].expand(n, q2), v_all[None, :].expand(n, q2))
        th, ph = torch.as_tensor(th, device=dev), torch.as_tensor(ph, device=dev)
        st, ct, sp, cp = torch.sin(th), torch.cos(th), torch.sin(ph), torch.cos(ph)
        r = torch.stack([st * cp, st * sp, ct], dim=-1)  # [n, Q2, 3]
        e_th = torch.stack([ct * cp, ct * sp, -st], dim=-1)
        e_ph = torch.stack([-sp, cp, torch.zeros_like(sp)], dim=-1)
        d = r[:, None, :, :] - r[:, :, None, :]  # d[a, b] = r_b - r_a
        x = (d * d).sum(-1) * 0.25  # sin^2(gamma/2) = |r_b - r_a|^2 / 4, no cancellation
        # angle of the great circle a -> b against the meridian at a: the tangent at a is the
        # projection of (r_b - r_a); likewise at b with (r_a - r_b), i.e. pointing back, which adds
        # pi to chi' -- immaterial in cos(2 (chi - chi'))
        chi_a = torch.atan2((d * e_ph[:, :, None, :]).sum(-1), (d * e_th[:, :, None, :]).sum(-1))
        chi_b = torch.atan2((-d * e_ph[:, None, :, :]).sum(-1), (-d * e_th[:, None, :, :]).sum(-1))
        c2 = torch.cos(2.0 * (chi_a - chi_b))
        c2 = torch.where(x > 0, c2, torch.ones_like(c2))  # a == b: no direction, same basis
        wt = (m[:, None, None] * ww[None]).expand_as(x)
        wp = wt * (1.0 - x) ** 2 * c2
        xr = x / x0
        pw = torch.ones_like(xr)
        for k in range(_K + 1):
            mt[k] += (wt * pw).sum()
            mp[k] += (wp * pw).sum()
            pw = pw * xr
    # normalised by the total weight as summed (= 12 nside^2 up to rounding): w^T_0 = 1 exactly
    return x0, (mt / mt[0]).cpu().numpy(), (mp / mt[0]).cpu().numpy()


def _series(lreal: np.ndarray, x0: float, mt: np.ndarray, mp: np.ndarray):
    """w^T_l^2 and w^P_l^2 at (real) multipoles from the moments: the terminating 2F1 series of
    d^l_00 and d^l_22 / (1-x)^2 in x, coefficient recurrences in k."""
    l = np.asarray(lreal, dtype=np.float64)
    ct = np.ones_like(l)
    cp = np.ones_like(l)
    wt = np.zeros_like(l)
    wp = np.zeros_like(l)
    n = l - 2.0
    for k in range(mt.size):
        if k > 0:
            ct = ct * (-(l + k) * (l - k + 1.0) / (k * k)) * x0
            cp = cp * (-(n - k + 1.0) * (n + k + 4.0) / (k * k)) * x0
        wt += ct * mt[k]
        wp += cp * mp[k]
    return wt, wp


def pixwin(nside: int, *, lmax: int | None = None, pol: bool = False):
    """NumPy arrays w^T_l (and w^P_l with ``pol``), l = 0..lmax (default 3 nside - 1, as healpy)."""
    nside = int(nside)
    if nside < 1 or nside & (nside - 1):
        raise ValueError("nside must be a power of two")
    if lmax is None:
        lmax = 3 * nside - 1
    lmax = int(lmax)
    if lmax > 4 * nside:
        raise ValueError(f"pixel window function of nside {nside} is tabulated up to lmax = {4 * nside}")
    base = min(nside, EXACT_NSIDE_MAX)
    x0, mt, mp = _pair_moments(base)
    ell = np.arange(lmax + 1, dtype=np.float64)
    lreal = ell if base == nside else (ell + 0.5) * (base / nside) - 0.5
    wt2, wp2 = _series(lreal, x0, mt, mp)
    wt = np.sqrt(np.clip(wt2, 0.0, None))
    if not pol:
        return wt
    if base != nside:
        # the spin-2 series lives on l >= 2: above the exact range carry the RATIO w^P / w^T of the
        # base window (1 + O(pixel area)) to the scaled multipole
        lb = np.maximum(lreal, 2.0)
        t2, p2 = _series(lb, x0, mt, mp)
        wp = wt * np.sqrt(np.clip(p2 / t2, 0.0, None))
    else:
        wp = np.sqrt(np.clip(wp2, 0.0, None))
    wp[: min(2, wp.size)] = 0.0
    return wt, wp
