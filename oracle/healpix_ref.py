"""
oracle.healpix_ref -- numpy restatement of the third-party leaf maths behind
``glass/healpix.py``.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

What the reference delegates to (none of it is under /root/reference):

* ``healpy.alm2map``       (glass/healpix.py:71)   -> :func:`alm2map`
* ``healpy.alm2map_spin``  (glass/healpix.py:107)  -> :func:`alm2map_spin`
* ``healpy.map2alm``       (glass/healpix.py:270)  -> :func:`map2alm`
* ``healpy.almxfl``        (glass/healpix.py:136)  -> :func:`almxfl`
* ``healpix.randang``      (glass/healpix.py:426)  -> :func:`ring2ang_uv`, :func:`randang`
* ``healpix.ang2pix``      (glass/healpix.py:172)  -> :func:`ang2pix`
* ``healpix.npix2nside``   (glass/healpix.py:293)  -> :func:`npix2nside`

The restatement follows the published HEALPix (Gorski et al. 2005) ring
geometry and libsharp-style ring transforms, with the
*standard* three-term Legendre recurrence in l (NOT the x^2 recurrence the CUDA
kernels use, so the two are independent).  Ground truth that pins this module:
:func:`alm2map_direct` / :func:`alm2map_spin_direct`, brute-force sums of
``scipy.special.sph_harm_y`` and of the closed-form Goldberg spin-weighted
harmonics at pixel centres.  **Parity with the healpy binaries is unpinned.**
"""

from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------
# pixelisation bookkeeping
# --------------------------------------------------------------------------


def nside2npix(nside: int) -> int:
    return 12 * int(nside) * int(nside)


def npix2nside(npix: int) -> int:
    nside = math.isqrt(int(npix) // 12)
    if 12 * nside * nside != int(npix) or nside < 1:
        raise ValueError(f"invalid npix: {npix}")
    return nside


def alm_size(lmax: int, mmax: int | None = None) -> int:
    mmax = lmax if mmax is None else mmax
    return mmax * (2 * lmax + 1 - mmax) // 2 + lmax + 1


def alm_lmax(size: int) -> int:
    """lmax from the alm length assuming mmax == lmax (healpy.Alm.getlmax)."""
    lmax = (math.isqrt(8 * size + 1) - 3) // 2
    if (lmax + 1) * (lmax + 2) // 2 != size:
        raise ValueError(f"invalid alm size: {size}")
    return lmax


def alm_index(lmax: int, l, m):
    """m-major (HEALPix) index, glass/fields.py:959-962."""
    return m * (2 * lmax + 1 - m) // 2 + l


def ring_info(nside: int):
    """
    Ring table, rings 1..4*nside-1 (index 0 = northernmost).

    Returns dict of arrays: nphi, start (first pixel), z (cos theta),
    sth (sin theta), phi0 (azimuth of first pixel).
    SURVEY.md Appendix A.1.
    """
    n = int(nside)
    i = np.arange(1, 4 * n, dtype=np.int64)
    north = i < n
    south = i > 3 * n
    ip = np.where(south, 4 * n - i, i)  # mirrored ring number for caps
    cap = north | south
    nphi = np.where(cap, 4 * ip, 4 * n)
    fi = ip.astype(np.float64)
    zcap = 1.0 - fi * fi / (3.0 * n * n)
    # sin(theta) in the caps from the small quantity t = i^2/(3 n^2): 1-z^2 = t(2-t)
    tcap = fi * fi / (3.0 * n * n)
    zeq = (2.0 * n - i) * 2.0 / (3.0 * n)
    z = np.where(cap, np.where(south, -zcap, zcap), zeq)
    sth = np.where(cap, np.sqrt(np.maximum(tcap * (2.0 - tcap), 0.0)), np.sqrt(np.maximum((1.0 - zeq) * (1.0 + zeq), 0.0)))
    shifted = np.where(cap, True, ((i - n) % 2) == 0)
    phi0 = np.where(shifted, np.pi / nphi, 0.0)
    start_n = 2 * ip * (ip - 1)
    start_e = 2 * n * (n - 1) + (i - n) * 4 * n
    start_s = 12 * n * n - 2 * ip * (ip + 1)
    start = np.where(north, start_n, np.where(south, start_s, start_e))
    return {"nphi": nphi, "start": start, "z": z, "sth": sth, "phi0": phi0, "shifted": shifted}


def pix2ang_centers(nside: int):
    """theta, phi of all pixel centres in RING order."""
    ri = ring_info(nside)
    theta = np.empty(nside2npix(nside))
    phi = np.empty(nside2npix(nside))
    for r in range(4 * nside - 1):
        s, n = int(ri["start"][r]), int(ri["nphi"][r])
        theta[s : s + n] = math.atan2(ri["sth"][r], ri["z"][r])
        phi[s : s + n] = ri["phi0"][r] + 2.0 * np.pi * np.arange(n) / n
    return theta, phi


# --------------------------------------------------------------------------
# pixel <-> angle  (healpix_bare restatement)
# --------------------------------------------------------------------------

_JRLL = np.array([2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4], dtype=np.int64)
_JPLL = np.array([1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7], dtype=np.int64)


def _isqrt(a):
    a = np.asarray(a, dtype=np.int64)
    r = np.floor(np.sqrt(a.astype(np.float64))).astype(np.int64)
    r = np.where(r * r > a, r - 1, r)
    r = np.where((r + 1) * (r + 1) <= a, r + 1, r)
    return r


def ring2xyf(nside: int, pix):
    """ring pixel index -> (x, y, face), all int64 arrays."""
    n = int(nside)
    pix = np.asarray(pix, dtype=np.int64)
    ncap = 2 * n * (n - 1)
    npix = 12 * n * n
    x = np.empty_like(pix)
    y = np.empty_like(pix)
    f = np.empty_like(pix)

    iring = np.empty_like(pix)
    iphi = np.empty_like(pix)
    kshift = np.zeros_like(pix)
    nr = np.empty_like(pix)

    nc = pix < ncap
    sc = pix >= npix - ncap
    eq = ~(nc | sc)

    # north cap
    p = pix[nc]
    ir = (1 + _isqrt(1 + 2 * p)) >> 1
    ip = p + 1 - 2 * ir * (ir - 1)
    iring[nc], iphi[nc], nr[nc] = ir, ip, ir
    f[nc] = (ip - 1) // ir

    # equatorial belt
    p = pix[eq] - ncap
    tmp = p // (4 * n)
    ir = tmp + n
    ip = p - tmp * 4 * n + 1
    iring[eq], iphi[eq], nr[eq] = ir, ip, n
    kshift[eq] = (ir + n) & 1
    ire = tmp + 1
    irm = 2 * n + 1 - tmp
    ifm = (ip - ire // 2 + n - 1) // n
    ifp = (ip - irm // 2 + n - 1) // n
    f[eq] = np.where(ifp == ifm, ifp | 4, np.where(ifp < ifm, ifp, ifm + 8))

    # south cap
    p = npix - pix[sc]
    ir = (1 + _isqrt(2 * p - 1)) >> 1
    ip = 4 * ir + 1 - (p - 2 * ir * (ir - 1))
    iring[sc], iphi[sc], nr[sc] = 4 * n - ir, ip, ir
    f[sc] = 8 + (ip - 1) // ir

    irt = iring - _JRLL[f] * n + 1
    ipt = 2 * iphi - _JPLL[f] * nr - kshift - 1
    ipt = np.where(ipt >= 2 * n, ipt - 8 * n, ipt)
    x = (ipt - irt) >> 1
    y = (-ipt - irt) >> 1
    return x, y, f


def hpc2loc(nside: int, x, y, f, u, v):
    """continuous in-face coordinates -> (z, sth, phi).  Appendix A.4."""
    n = float(nside)
    X = (np.asarray(x, dtype=np.float64) + u) / n
    Y = (np.asarray(y, dtype=np.float64) + v) / n
    jr = _JRLL[f] - X - Y
    jp = _JPLL[f].astype(np.float64)
    north = jr < 1.0
    south = jr > 3.0
    nr = np.where(north, jr, np.where(south, 4.0 - jr, 1.0))
    tmp = nr * nr / 3.0
    zc = 1.0 - tmp
    sc = np.sqrt(tmp * (2.0 - tmp))
    ze = (2.0 - jr) * 2.0 / 3.0
    se = np.sqrt(np.maximum((1.0 - ze) * (1.0 + ze), 0.0))
    cap = north | south
    z = np.where(cap, np.where(south, -zc, zc), ze)
    sth = np.where(cap, sc, se)
    with np.errstate(divide="ignore", invalid="ignore"):
        tmpphi = np.where(cap, (X - Y) / nr, X - Y)
    tmpphi = jp + tmpphi
    tmpphi = np.where(tmpphi < 0, tmpphi + 8.0, tmpphi)
    tmpphi = np.where(tmpphi >= 8.0, tmpphi - 8.0, tmpphi)
    phi = (np.pi / 4.0) * tmpphi
    return z, sth, phi


def ring2ang_uv(nside: int, ipix, u, v, lonlat: bool = False):
    """
    Position inside ring-pixel ``ipix`` at in-pixel offsets (u, v) in [0,1)^2.
    u = v = 0.5 is the pixel centre.  Restates healpix ``_chp.ring2ang_uv``
    (reached from glass/healpix.py:426-431).
    Returns (theta, phi) in radians or (lon, lat) in degrees.
    """
    x, y, f = ring2xyf(nside, ipix)
    z, sth, phi = hpc2loc(nside, x, y, f, np.asarray(u, dtype=np.float64), np.asarray(v, dtype=np.float64))
    theta = np.arctan2(sth, z)
    if lonlat:
        return np.degrees(phi), 90.0 - np.degrees(theta)
    return theta, phi


def randang(nside: int, ipix, lonlat: bool = False, rng=None):
    """
    healpix.randang as called by glass/healpix.py:426-431 (fresh seed-42
    generator per call).  Draw order (all u, then all v) is UNVERIFIED against
    the binary; parity tests feed (u, v) explicitly instead.
    """
    rng = np.random.default_rng(42) if rng is None else rng
    ipix = np.asarray(ipix, dtype=np.int64)
    u = rng.random(ipix.shape)
    v = rng.random(ipix.shape)
    return ring2ang_uv(nside, ipix, u, v, lonlat=lonlat)


def zphi2pix_ring(nside: int, z, sth, phi):
    """(z, sin theta, phi) -> ring pixel.  Appendix A.5 (healpix_cxx loc2pix)."""
    n = int(nside)
    z = np.asarray(z, dtype=np.float64)
    sth = np.asarray(sth, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    za = np.abs(z)
    tt = np.mod(phi, 2.0 * np.pi) * (2.0 / np.pi)
    tt = np.where(tt >= 4.0, tt - 4.0, tt)
    ncap = 2 * n * (n - 1)
    npix = 12 * n * n

    # equatorial
    t1 = n * (0.5 + tt)
    t2 = n * z * 0.75
    jp = np.floor(t1 - t2).astype(np.int64)
    jm = np.floor(t1 + t2).astype(np.int64)
    ir = n + 1 + jp - jm
    kshift = 1 - (ir & 1)
    t = jp + jm - n + kshift + 1 + 8 * n
    ip = (t >> 1) % (4 * n)
    pix_eq = ncap + (ir - 1) * 4 * n + ip

    # caps
    tp = tt - np.floor(tt)
    use_sth = za > 0.99
    with np.errstate(invalid="ignore"):
        tmp = np.where(use_sth, n * sth / np.sqrt((1.0 + za) / 3.0), n * np.sqrt(3.0 * (1.0 - za)))
    jp = (tp * tmp).astype(np.int64)
    jm = ((1.0 - tp) * tmp).astype(np.int64)
    ir = jp + jm + 1
    ip = (tt * ir).astype(np.int64)
    ip = np.where(ip >= 4 * ir, ip - 4 * ir, ip)
    ip = np.where(ip < 0, ip + 4 * ir, ip)
    pix_n = 2 * ir * (ir - 1) + ip
    pix_s = npix - 2 * ir * (ir + 1) + ip
    return np.where(za <= 2.0 / 3.0, pix_eq, np.where(z > 0, pix_n, pix_s))


def ang2pix(nside: int, theta, phi, lonlat: bool = False):
    """healpix.ang2pix as called by glass/healpix.py:172 (RING scheme)."""
    theta = np.asarray(theta, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    if lonlat:
        lon, lat = theta, phi
        theta = np.radians(90.0 - lat)
        phi = np.radians(lon)
    return zphi2pix_ring(nside, np.cos(theta), np.sin(theta), phi)


# --------------------------------------------------------------------------
# Legendre / Wigner-d recurrences (standard, in l)
# --------------------------------------------------------------------------


def _log_lam_mm(m: int, sth):
    """log |lambda_mm| = log sqrt((2m+1)!!/(4 pi (2m)!!)) + m log sin(theta)."""
    k = np.arange(1, m + 1, dtype=np.float64)
    logc = 0.5 * (np.sum(np.log1p(0.5 / k)) - math.log(4.0 * math.pi))
    with np.errstate(divide="ignore"):
        return logc + m * np.log(sth)


def lam_lm(lmax: int, m: int, z, sth):
    """
    Normalised associated Legendre functions lambda_lm(z) = Y_lm(theta, 0),
    l = m..lmax, with Condon-Shortley phase.  Returns array [lmax-m+1, nz].
    Standard recurrence, SURVEY.md Appendix A.2.
    """
    z = np.asarray(z, dtype=np.float64)
    out = np.zeros((lmax - m + 1, z.size))
    lam0 = np.exp(_log_lam_mm(m, np.asarray(sth, dtype=np.float64))) * (-1.0) ** m
    out[0] = lam0
    if lmax > m:
        out[1] = z * math.sqrt(2.0 * m + 3.0) * lam0
    for l in range(m + 2, lmax + 1):
        a = math.sqrt((4.0 * l * l - 1.0) / (l * l - m * m))
        b = math.sqrt(((l - 1.0) ** 2 - m * m) / (4.0 * (l - 1.0) ** 2 - 1.0))
        out[l - m] = a * (z * out[l - m - 1] - b * out[l - m - 2])
    return out


def _log_fact(n):
    return math.lgamma(n + 1.0)


def wigner_d_l(lmax: int, m: int, mp: int, theta):
    """
    Wigner small-d d^l_{m,mp}(theta) for l = l0..lmax, l0 = max(|m|,|mp|), by the
    three-term recurrence in l.  Returns array [lmax-l0+1, ntheta] (empty if
    l0 > lmax).  Convention: d^j_{m'm} of Wikipedia / Varshalovich
    (d^1_{1,0} = -sin(theta)/sqrt(2)).
    """
    theta = np.asarray(theta, dtype=np.float64)
    l0 = max(abs(m), abs(mp))
    if l0 > lmax:
        return np.zeros((0, theta.size))
    c = np.cos(theta)
    ch = np.cos(0.5 * theta)
    sh = np.sin(0.5 * theta)
    out = np.zeros((lmax - l0 + 1, theta.size))
    # seed d^{l0}: use symmetries to reduce to m = l0 >= |mp| :
    # d^j_{j,mp} = sqrt((2j)!/((j+mp)!(j-mp)!)) cos^{j+mp}(t/2) (-sin(t/2))^{j-mp}  [d_{m',m} with m'=j]
    # symmetries: d_{m',m} = (-1)^{m-m'} d_{m,m'} = d_{-m,-m'}
    j = l0
    a, b, sign = m, mp, 1.0
    if abs(a) < abs(b):  # swap so that |a| = j
        a, b = b, a
        sign *= (-1.0) ** (a - b)
    if a < 0:  # flip both signs: d_{a,b} = d_{-b,-a} = (-1)^{a-b} d_{-a,-b}
        sign *= (-1.0) ** (a - b)
        a, b = -a, -b
    lognorm = 0.5 * (_log_fact(2 * j) - _log_fact(j + b) - _log_fact(j - b))
    with np.errstate(divide="ignore"):
        mag = np.exp(lognorm + (j + b) * np.log(ch) + (j - b) * np.log(sh))
    # d^j_{j,b} has sign (-1)^{j-b} in the Wikipedia d^j_{m'm} convention
    out[0] = sign * mag * (-1.0) ** (j - b)
    if lmax > l0:
        if l0 == 0:
            out[1] = c
        else:
            l = l0
            den = l * math.sqrt(((l + 1.0) ** 2 - m * m) * ((l + 1.0) ** 2 - mp * mp))
            out[1] = (2 * l + 1.0) * (l * (l + 1.0) * c - m * mp) * out[0] / den
    for l in range(l0 + 1, lmax):
        den = l * math.sqrt(((l + 1.0) ** 2 - m * m) * ((l + 1.0) ** 2 - mp * mp))
        t1 = (2 * l + 1.0) * (l * (l + 1.0) * c - m * mp)
        t2 = (l + 1.0) * math.sqrt((l * l - m * m) * (l * l - mp * mp))
        out[l + 1 - l0] = (t1 * out[l - l0] - t2 * out[l - 1 - l0]) / den
    return out


def slam_lm(lmax: int, m: int, s: int, theta):
    """
    Spin-weighted lambda: sY_lm(theta, 0) for l = 0..lmax (zeros below
    max(|m|,|s|)).  sY_lm = (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta).
    Returns array [lmax+1, ntheta].
    """
    theta = np.asarray(theta, dtype=np.float64)
    out = np.zeros((lmax + 1, theta.size))
    l0 = max(abs(m), abs(s))
    if l0 > lmax:
        return out
    d = wigner_d_l(lmax, m, -s, theta)
    l = np.arange(l0, lmax + 1, dtype=np.float64)
    out[l0:] = (-1.0) ** s * np.sqrt((2.0 * l + 1.0) / (4.0 * np.pi))[:, None] * d
    return out


def sYlm_goldberg(s: int, l: int, m: int, theta, phi):
    """Closed-form spin-weighted spherical harmonic (Goldberg et al. 1967)."""
    theta = np.asarray(theta, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    if l < max(abs(m), abs(s)):
        return np.zeros(theta.shape, dtype=np.complex128)
    f = math.factorial
    norm = math.sqrt(f(l + m) * f(l - m) * (2 * l + 1) / (4.0 * math.pi * f(l + s) * f(l - s)))
    sh = np.sin(0.5 * theta)
    ch = np.cos(0.5 * theta)
    acc = np.zeros(theta.shape)
    for r in range(0, l - s + 1):
        k2 = r + s - m
        if k2 < 0 or k2 > l + s:
            continue
        coef = math.comb(l - s, r) * math.comb(l + s, k2) * (-1.0) ** (l - r - s)
        # sin^{2l}(t/2) cot^{2r+s-m}(t/2) = cos^{2r+s-m} sin^{2l-2r-s+m}
        acc = acc + coef * ch ** (2 * r + s - m) * sh ** (2 * l - 2 * r - s + m)
    return (-1.0) ** m * norm * acc * np.exp(1j * m * phi)


# --------------------------------------------------------------------------
# ring <-> Fourier helpers
# --------------------------------------------------------------------------


def _phases_to_ring(F, nphi: int, phi0: float):
    """
    F[m], m=0..mmax (complex) on one ring -> nphi real samples
    f_j = Re F_0 + 2 Re sum_{m>0} F_m exp(i m (phi0 + 2 pi j / nphi)),
    via alias folding into half-complex bins and an inverse real FFT.
    SURVEY.md Appendix A.2.
    """
    mmax = F.shape[0] - 1
    m = np.arange(mmax + 1)
    t = F * np.exp(1j * m * phi0)
    G = np.zeros(nphi // 2 + 1, dtype=np.complex128)
    G[0] += t[0].real
    for mm in range(1, mmax + 1):
        k = mm % nphi
        if k == 0 or 2 * k == nphi:
            G[k] += 2.0 * t[mm].real
        elif k < nphi - k:
            G[k] += t[mm]
        else:
            G[nphi - k] += np.conj(t[mm])
    # f_j = G0 + 2 Re sum_{0<k<n/2} G_k w^{jk} + G_{n/2} (-1)^j  ==  n * irfft(G)
    return np.fft.irfft(G, nphi) * nphi


def _ring_to_phases(f, nphi: int, phi0: float, mmax: int):
    """adjoint-side helper: G_m = sum_j f_j exp(-i m phi_j), m = 0..mmax."""
    X = np.fft.rfft(f)  # X[k] = sum_j f_j e^{-2 pi i jk/n}, k=0..n/2
    m = np.arange(mmax + 1)
    k = m % nphi
    full = np.where(k <= nphi // 2, X[np.minimum(k, nphi // 2)], np.conj(X[np.minimum(nphi - k, nphi // 2)]))
    return full * np.exp(-1j * m * phi0)


# --------------------------------------------------------------------------
# transforms
# --------------------------------------------------------------------------


def alm2map_direct(alm, nside: int, lmax: int | None = None):
    """Ground truth: f(p) = sum_lm a_lm Y_lm(p) with scipy's Y_lm, real field."""
    from scipy.special import sph_harm_y

    alm = np.asarray(alm, dtype=np.complex128)
    lmax = alm_lmax(alm.size) if lmax is None else lmax
    theta, phi = pix2ang_centers(nside)
    out = np.zeros(theta.size)
    for m in range(lmax + 1):
        for l in range(m, lmax + 1):
            a = alm[alm_index(lmax, l, m)]
            y = sph_harm_y(l, m, theta, phi)
            out += (a * y).real if m == 0 else 2.0 * (a * y).real
    return out


def alm2map(alm, nside: int, lmax: int | None = None, mmax: int | None = None):
    """
    Scalar synthesis on HEALPix rings (healpy.alm2map with pol=False,
    pixwin=False; glass/healpix.py:71, called from glass/fields.py:429 and
    glass/lensing.py:326).  alm m-major complex128; returns RING map float64.
    """
    alm = np.asarray(alm, dtype=np.complex128)
    lmax = alm_lmax(alm.size) if lmax is None else lmax
    mmax = lmax if mmax is None else mmax
    return phase2map(alm2phase(alm, nside, lmax, mmax), nside)


def alm2phase(alm, nside: int, lmax: int, mmax: int | None = None):
    """Legendre stage: F[ring, m] = sum_l a_lm lambda_lm(cos theta_ring)."""
    alm = np.asarray(alm, dtype=np.complex128)
    mmax = lmax if mmax is None else mmax
    ri = ring_info(nside)
    nring = 4 * nside - 1
    F = np.zeros((nring, mmax + 1), dtype=np.complex128)
    for m in range(mmax + 1):
        lam = lam_lm(lmax, m, ri["z"], ri["sth"])  # [l-m, ring]
        a = alm[alm_index(lmax, m, m) : alm_index(lmax, lmax, m) + 1].copy()
        if m == 0:
            a = a.real + 0j
        F[:, m] = a @ lam
    return F


def phase2map(F, nside: int):
    """Fourier stage: F[ring, m] -> RING map."""
    ri = ring_info(nside)
    nring = 4 * nside - 1
    out = np.empty(nside2npix(nside))
    for r in range(nring):
        s, n = int(ri["start"][r]), int(ri["nphi"][r])
        out[s : s + n] = _phases_to_ring(F[r], n, float(ri["phi0"][r]))
    return out


def _spin_pm_coeffs(alm1, alm2, spin: int):
    """HEALPix alm2map_spin convention: +s and -s coefficients."""
    a_p = -(alm1 + 1j * alm2)
    a_m = -((-1.0) ** spin) * (alm1 - 1j * alm2)
    return a_p, a_m


def alm2map_spin_direct(alm1, alm2, nside: int, spin: int, lmax: int):
    """
    Ground truth for healpy.alm2map_spin:  map1 + i map2 = sum_{l,m} sa_lm sY_lm
    with sa_lm = -(alm1 + i alm2) for m >= 0 and the reality conditions of
    alm1/alm2 for m < 0 (alm_{l,-m} = (-1)^m conj(alm_{lm})).
    """
    alm1 = np.asarray(alm1, dtype=np.complex128)
    alm2 = np.asarray(alm2, dtype=np.complex128)
    theta, phi = pix2ang_centers(nside)
    out = np.zeros(theta.size, dtype=np.complex128)
    for l in range(abs(spin), lmax + 1):
        for m in range(-l, l + 1):
            i = alm_index(lmax, l, abs(m))
            if m >= 0:
                e, b = alm1[i], alm2[i]
            else:
                e, b = (-1.0) ** m * np.conj(alm1[i]), (-1.0) ** m * np.conj(alm2[i])
            out += -(e + 1j * b) * sYlm_goldberg(spin, l, m, theta, phi)
    return out.real, out.imag


def alm2map_spin(alm1, alm2, nside: int, spin: int, lmax: int, mmax: int | None = None):
    """
    Spin-s synthesis on HEALPix rings (healpy.alm2map_spin; glass/healpix.py:107,
    called from glass/lensing.py:343,366,428).  Returns (map1, map2).
    """
    if spin <= 0:
        raise ValueError("spin must be positive")
    alm1 = np.asarray(alm1, dtype=np.complex128)
    alm2 = np.asarray(alm2, dtype=np.complex128)
    mmax = lmax if mmax is None else mmax
    ri = ring_info(nside)
    theta = np.arctan2(ri["sth"], ri["z"])
    nring = 4 * nside - 1
    # complex field P = sum_{m=-mmax..mmax} P_m(theta) e^{i m phi}
    Fp = np.zeros((nring, mmax + 1), dtype=np.complex128)  # m >= 0
    Fn = np.zeros((nring, mmax + 1), dtype=np.complex128)  # -m  (index m)
    for m in range(mmax + 1):
        sl = slice(alm_index(lmax, m, m), alm_index(lmax, lmax, m) + 1)
        e = np.zeros(lmax + 1, dtype=np.complex128)
        b = np.zeros(lmax + 1, dtype=np.complex128)
        e[m:] = alm1[sl]
        b[m:] = alm2[sl]
        yp = slam_lm(lmax, m, spin, theta)  # sY_{l,m}(theta,0)
        Fp[:, m] = (-(e + 1j * b)) @ yp
        if m > 0:
            yn = slam_lm(lmax, -m, spin, theta)  # sY_{l,-m}
            en = (-1.0) ** m * np.conj(e)
            bn = (-1.0) ** m * np.conj(b)
            Fn[:, m] = (-(en + 1j * bn)) @ yn
    map1 = np.empty(nside2npix(nside))
    map2 = np.empty(nside2npix(nside))
    for r in range(nring):
        s, n = int(ri["start"][r]), int(ri["nphi"][r])
        # real part: Re P = sum_m [ (P_m + conj(P_{-m}))/2 ] e^{im phi} + c.c.
        A = 0.5 * (Fp[r] + np.conj(Fn[r]))
        B = -0.5j * (Fp[r] - np.conj(Fn[r]))
        A[0] = Fp[r, 0].real
        B[0] = Fp[r, 0].imag
        map1[s : s + n] = _phases_to_ring(A, n, float(ri["phi0"][r]))
        map2[s : s + n] = _phases_to_ring(B, n, float(ri["phi0"][r]))
    return map1, map2


def ring_weights_uniform(nside: int):
    return np.ones(4 * nside - 1)


def _analysis_pass(mp, nside: int, lmax: int, mmax: int, ring_w):
    ri = ring_info(nside)
    nring = 4 * nside - 1
    G = np.zeros((nring, mmax + 1), dtype=np.complex128)
    for r in range(nring):
        s, n = int(ri["start"][r]), int(ri["nphi"][r])
        G[r] = _ring_to_phases(mp[s : s + n], n, float(ri["phi0"][r]), mmax) * ring_w[r]
    G *= 4.0 * np.pi / nside2npix(nside)
    alm = np.zeros(alm_size(lmax, mmax), dtype=np.complex128)
    for m in range(mmax + 1):
        lam = lam_lm(lmax, m, ri["z"], ri["sth"])
        alm[alm_index(lmax, m, m) : alm_index(lmax, lmax, m) + 1] = lam @ G[:, m]
    return alm


def map2alm(mp, lmax: int | None = None, mmax: int | None = None, niter: int = 3, ring_w=None):
    """
    Scalar analysis (healpy.map2alm, pol=False; glass/healpix.py:270 called from
    glass/lensing.py:306,408):  a_lm = sum_p w_p f_p conj(Y_lm(p)) 4pi/npix, then
    ``niter`` Jacobi refinements alm += A(map - S(alm)) (healpy's default
    iter=3 is not overridden by GLASS).  ``ring_w`` stands in for healpy's
    pixel-weight data files, which cannot be obtained offline; default uniform.
    """
    mp = np.asarray(mp, dtype=np.float64)
    nside = npix2nside(mp.size)
    lmax = 3 * nside - 1 if lmax is None else lmax
    mmax = lmax if mmax is None else mmax
    ring_w = ring_weights_uniform(nside) if ring_w is None else np.asarray(ring_w)
    alm = _analysis_pass(mp, nside, lmax, mmax, ring_w)
    for _ in range(niter):
        res = mp - alm2map(alm, nside, lmax, mmax)
        alm = alm + _analysis_pass(res, nside, lmax, mmax, ring_w)
    return alm


def almxfl(alm, fl, mmax: int | None = None):
    """healpy.almxfl (glass/healpix.py:136): a_lm *= f_l, zero beyond len(fl)."""
    alm = np.array(alm, dtype=np.complex128, copy=True)
    lmax = alm_lmax(alm.size) if mmax is None else None
    if lmax is None:
        raise NotImplementedError("mmax != lmax")
    fl = np.asarray(fl)
    f = np.zeros(lmax + 1, dtype=fl.dtype if np.iscomplexobj(fl) else np.float64)
    n = min(lmax + 1, fl.shape[0])
    f[:n] = fl[:n]
    for m in range(lmax + 1):
        alm[alm_index(lmax, m, m) : alm_index(lmax, lmax, m) + 1] *= f[m:]
    return alm


# --------------------------------------------------------------------------
# visibility-mask helpers behind glass.vmap_galactic_ecliptic (glass/observations.py:96-100):
# healpy.query_strip (glass/healpix.py:389), healpy.Rotator(coord=).rotate_map_pixel
# (glass/healpix.py:471).  Third-party (healpy >= 1.15.0 / healpix_cxx, absent from the tree):
# restated from the published HEALPix C++ algorithms -- T_Healpix_Base::ring_above,
# query_strip_internal, get_ring_info2, get_interpol (bilinear weights of the four pixels on the two
# neighbouring rings) -- and healpy's rotator conventions (rotator.py: get_coordconv_matrix with the
# HEALPix ecliptic->galactic matrix and the J2000 obliquity, Rotator.I = inverse rotation,
# rotate_map_pixel = interpolate the input map at the back-rotated pixel centres).
# **Parity with the healpy binaries is unpinned.**
# --------------------------------------------------------------------------


def ring_above(nside: int, z):
    """Index of the ring just north of cos(theta) = z (0 = above the first ring)."""
    z = np.asarray(z, dtype=np.float64)
    az = np.abs(z)
    eq = (nside * (2.0 - 1.5 * z)).astype(np.int64)
    ir = (nside * np.sqrt(3.0 * (1.0 - az))).astype(np.int64)
    return np.where(az <= 2.0 / 3.0, eq, np.where(z > 0, ir, 4 * nside - ir - 1))


def _ring_small(nside: int, ring: int):
    """(first pixel, pixels) of ring 0..4 nside-1 (ring 0 is the empty 'ring' above the pole)."""
    n = nside
    if ring < n:
        return 2 * ring * (ring - 1), 4 * ring
    if ring < 3 * n:
        return 2 * n * (n - 1) + (ring - n) * 4 * n, 4 * n
    nr = 4 * n - ring
    return 12 * n * n - 2 * nr * (nr + 1), 4 * nr


def _strip_internal(nside: int, theta1: float, theta2: float, out):
    ring1 = max(1, 1 + int(ring_above(nside, math.cos(theta1))))
    ring2 = min(4 * nside - 1, int(ring_above(nside, math.cos(theta2))))
    sp1, _ = _ring_small(nside, ring1)
    sp2, rp2 = _ring_small(nside, ring2)
    if sp1 <= sp2 + rp2:
        out[sp1 : sp2 + rp2] = 1


def query_strip(nside: int, theta1: float, theta2: float):
    """0/1 mask of the pixels whose centres lie in the colatitude strip (healpy.query_strip,
    inclusive=False, RING): theta1 < theta2 the strip between them, otherwise its complement."""
    out = np.zeros(nside2npix(nside))
    if theta1 < theta2:
        _strip_internal(nside, theta1, theta2, out)
    else:
        _strip_internal(nside, 0.0, theta2, out)
        _strip_internal(nside, theta1, math.pi, out)
    return out


def _ring_info2(nside: int, ring):
    """(first pixel, pixels, theta, shifted) of rings 1..4 nside-1, vectorised."""
    n = nside
    ring = np.asarray(ring, dtype=np.int64)
    north = np.where(ring > 2 * n, 4 * n - ring, ring)
    cap = north < n
    tmp = north * north / (3.0 * n * n)
    th_cap = np.arctan2(np.sqrt(tmp * (2.0 - tmp)), 1.0 - tmp)
    th_eq = np.arccos(np.clip((2 * n - north) * (2.0 / (3.0 * n)), -1.0, 1.0))
    theta = np.where(cap, th_cap, th_eq)
    nr = np.where(cap, 4 * north, 4 * n)
    shifted = np.where(cap, True, ((north - n) & 1) == 0)
    sp = np.where(cap, 2 * north * (north - 1), 2 * n * (n - 1) + (north - n) * 4 * n)
    south = north != ring
    theta = np.where(south, np.pi - theta, theta)
    sp = np.where(south, 12 * n * n - sp - nr, sp)
    return sp, nr, theta, shifted


def get_interpol(nside: int, theta, phi):
    """Pixels (4, N) and weights (4, N) of HEALPix's bilinear interpolation (RING)."""
    n = nside
    npix = 12 * n * n
    theta = np.asarray(theta, dtype=np.float64)
    phi = np.mod(np.asarray(phi, dtype=np.float64), 2.0 * np.pi)
    phi = np.where(phi >= 2.0 * np.pi, 0.0, phi)  # fmodulo: a tiny negative angle must not round to 2 pi
    ir1 = ring_above(n, np.cos(theta))
    ir2 = ir1 + 1
    pix = np.zeros((4,) + theta.shape, dtype=np.int64)
    wgt = np.zeros((4,) + theta.shape)
    th = [None, None]
    for k, ir in enumerate((ir1, ir2)):
        valid = (ir > 0) & (ir < 4 * n)
        sp, nr, th[k], shift = _ring_info2(n, np.clip(ir, 1, 4 * n - 1))
        dphi = 2.0 * np.pi / nr
        tmp = phi / dphi - 0.5 * shift
        i1 = np.where(tmp < 0, tmp.astype(np.int64) - 1, tmp.astype(np.int64))
        w1 = (phi - (i1 + 0.5 * shift) * dphi) / dphi
        i2 = i1 + 1
        i1 = np.where(i1 < 0, i1 + nr, i1)
        i2 = np.where(i2 >= nr, i2 - nr, i2)
        pix[2 * k] = np.where(valid, sp + i1, 0)
        pix[2 * k + 1] = np.where(valid, sp + i2, 0)
        wgt[2 * k] = np.where(valid, 1.0 - w1, 0.0)
        wgt[2 * k + 1] = np.where(valid, w1, 0.0)
    theta1, theta2 = th
    top, bot = ir1 == 0, ir2 == 4 * n
    mid = ~(top | bot)
    # north of the first ring: the four pixels of ring 1, the opposite pair weighted by the rest
    wt = theta / theta2
    fac = (1.0 - wt) * 0.25
    w = wgt.copy()
    p = pix.copy()
    w[0] = np.where(top, fac, w[0])
    w[1] = np.where(top, fac, w[1])
    w[2] = np.where(top, wgt[2] * wt + fac, w[2])
    w[3] = np.where(top, wgt[3] * wt + fac, w[3])
    p[0] = np.where(top, (pix[2] + 2) & 3, p[0])
    p[1] = np.where(top, (pix[3] + 2) & 3, p[1])
    # south of the last ring
    wb = (theta - theta1) / (np.pi - theta1)
    facb = wb * 0.25
    w[0] = np.where(bot, wgt[0] * (1.0 - wb) + facb, w[0])
    w[1] = np.where(bot, wgt[1] * (1.0 - wb) + facb, w[1])
    w[2] = np.where(bot, facb, w[2])
    w[3] = np.where(bot, facb, w[3])
    p[2] = np.where(bot, ((pix[0] + 2) & 3) + npix - 4, p[2])
    p[3] = np.where(bot, ((pix[1] + 2) & 3) + npix - 4, p[3])
    # between two rings
    with np.errstate(invalid="ignore", divide="ignore"):
        wm = (theta - theta1) / (theta2 - theta1)
        w[0] = np.where(mid, wgt[0] * (1.0 - wm), w[0])
        w[1] = np.where(mid, wgt[1] * (1.0 - wm), w[1])
        w[2] = np.where(mid, wgt[2] * wm, w[2])
        w[3] = np.where(mid, wgt[3] * wm, w[3])
    return p, w


def get_interp_val(m, theta, phi):
    """healpy.get_interp_val (RING): sum over the four pixels of m[p] w, rows added in order."""
    m = np.asarray(m, dtype=np.float64)
    p, w = get_interpol(npix2nside(m.size), theta, phi)
    return np.sum(m[p] * w, 0)


def coordconv_matrix(coord: str):
    """healpy.rotator.get_coordconv_matrix for coord = two letters of G, E, C (C = Q, equatorial)."""
    a, b = (c.upper().replace("Q", "C") for c in coord)
    eps = (23.452294 - 0.0130125 - 1.63889e-6 + 5.02778e-7) * np.pi / 180.0
    e2g = np.array(
        [
            [-0.054882486, -0.993821033, -0.096476249],
            [0.494116468, -0.110993846, 0.862281440],
            [-0.867661702, -0.000346354, 0.497154957],
        ]
    )
    e2q = np.array([[1.0, 0.0, 0.0], [0.0, np.cos(eps), -np.sin(eps)], [0.0, np.sin(eps), np.cos(eps)]])
    g2e = np.linalg.inv(e2g)
    q2e = np.linalg.inv(e2q)
    table = {
        "EG": e2g,
        "GE": g2e,
        "EC": e2q,
        "CE": q2e,
        "GC": e2q @ g2e,
        "CG": e2g @ q2e,
    }
    return np.identity(3) if a == b else table[a + b]


def rotate_map_pixel(m, coord: str):
    """healpy.Rotator(coord=coord).rotate_map_pixel(m) for one scalar map."""
    m = np.asarray(m, dtype=np.float64)
    nside = npix2nside(m.size)
    theta, phi = pix2ang_centers(nside)
    mat = np.linalg.inv(coordconv_matrix(coord))  # Rotator.I: back to the original frame
    st = np.sin(theta)
    v = mat @ np.stack([st * np.cos(phi), st * np.sin(phi), np.cos(theta)])
    th = np.arctan2(np.hypot(v[0], v[1]), v[2])
    ph = np.arctan2(v[1], v[0])
    return get_interp_val(m, th, ph)


def vmap_galactic_ecliptic(nside: int, galactic=(30, 90), ecliptic=(20, 80)):
    """glass.vmap_galactic_ecliptic (glass/observations.py:51-101) on the functions above."""
    m = np.ones(nside2npix(nside))
    m *= 1 - query_strip(nside, *galactic)
    m = rotate_map_pixel(m, "GC")
    m *= 1 - query_strip(nside, *ecliptic)
    return rotate_map_pixel(m, "CE")
