"""
oracle.glass_ref -- NumPy restatement of the GLASS-side code on the hot path, with every
random draw replaced by *supplied* deviates so results are deterministic functions of
their inputs.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Each function cites the reference lines it follows.  Pinned by (i) the golden vectors in
``tests/golden/`` that ``tests/golden/make_golden.py`` produced by executing the
reference's own source, (ii) the reference's dependency-free known-answer tests, ported
in ``tests/test_oracle_glass.py``.
"""

from __future__ import annotations

import itertools
import math

import numpy as np

from . import healpix_ref as H

ARCMIN2_SPHERE = 60**6 // 100 / math.pi  # glass/points.py:72


# ----------------------------------------------------------------------------
# fields
# ----------------------------------------------------------------------------


def iternorm(cov_rows):
    """glass/fields.py:101-188: incremental banded Cholesky rows [a, s]."""
    out = []
    m = a = s = shape = None
    for idx, row in enumerate(cov_rows):
        row = np.asarray(row, dtype=np.float64)
        k = row.shape[-1] - 1
        if k < 0:
            raise ValueError("empty covariance matrix")
        if idx == 0:
            shape = row.shape[:-1]
            m = np.zeros(shape + (k, k))
            a = np.zeros(shape + (k,))
            s = np.ones(shape)
        else:
            if row.shape[:-1] != shape:
                raise ValueError("shape mismatch in covariance")
            atm = np.matmul(a[..., None, :], m)
            m = np.concatenate([m, np.zeros(shape + (m.shape[-2], 1))], axis=-1)
            u = (s > 0).astype(np.float64)
            newrow = np.concatenate([-atm, u[..., None, None]], axis=-1)
            s = np.where(s > 0, s, 1.0)
            newrow = newrow / s[..., None, None]
            m = np.concatenate([m, newrow], axis=-2)
        m = m[..., m.shape[-2] - k :, m.shape[-1] - k :]
        c = row[..., :0:-1]
        a = np.matmul(m, c[..., None])[..., 0]
        s = row[..., 0] - np.vecdot(a, a)
        if np.any(s < 0):
            raise ValueError("covariance matrix is not positive definite")
        s = np.sqrt(s)
        out.append(np.concatenate([a, s[..., None]], axis=-1))
    return out


def cls2cov_rows(cls, nl, nf, nc):
    """glass/fields.py:191-236 -- returns COPIES of each row block (list)."""
    rows = []
    cov = np.zeros((nl, nc + 1))
    end = 0
    for j in range(nf):
        begin, end = end, end + j + 1
        for i, cl in enumerate(cls[begin:end][: nc + 1]):
            cl = np.asarray(cl, dtype=np.float64)
            if i == 0 and np.any(cl < 0):
                raise ValueError("negative values in cl")
            n = cl.shape[0]
            cov[:n, i] = cl
            cov[n:, i] = 0.0
        cov /= 2
        rows.append(cov.copy())
    return rows


def multalm(alm, bl):
    """glass/harmonics.py:46-47."""
    bl = np.asarray(bl)
    return np.asarray(alm) * np.repeat(bl, np.arange(bl.size) + 1)


def glass_to_healpix_alm(alm):
    """glass/fields.py:943-962."""
    alm = np.asarray(alm)
    n = math.isqrt(2 * alm.size)
    if n * (n + 1) // 2 != alm.size:
        raise ValueError(f"not a triangle number: {alm.size}")
    ell = np.arange(n)
    parts = [alm[ell[m:] * (ell[m:] + 1) // 2 + m] for m in range(n)]
    return np.concatenate(parts) if parts else alm


def getcl(cls, i, j, lmax=None):
    """glass/fields.py:525-560."""
    if j > i:
        i, j = j, i
    cl = np.asarray(cls[i * (i + 1) // 2 + i - j])
    if lmax is not None:
        cl = cl[: lmax + 1] if cl.shape[0] > lmax + 1 else np.pad(cl, (0, lmax + 1 - cl.shape[0]))
    return cl


def cltovar(cl):
    """transformcl.cltovar (glass/fields.py:890): sum (2l+1)/(4pi) C_l."""
    cl = np.asarray(cl, dtype=np.float64)
    ell = np.arange(cl.shape[0])
    return float(np.sum((2 * ell + 1) / (4 * np.pi) * cl))


def generate_alms(gls, ncorr, zs):
    """
    The alm (m-major, m=0 made real) of every shell, glass/fields.py:377-425, given the
    complex standard normals ``zs[j]`` (GLASS order) that fields.py:407 would draw.
    """
    ngrf = math.isqrt(2 * len(gls))
    if ngrf * (ngrf + 1) // 2 != len(gls):
        raise ValueError(f"invalid number of spectra: {len(gls)}")
    if ncorr is None:
        ncorr = ngrf - 1
    n = max((np.asarray(g).shape[0] for g in gls), default=0)
    if n == 0:
        raise ValueError("all gls are empty")
    ws = iternorm(cls2cov_rows(gls, n, ngrf, ncorr))
    y = []
    alms = []
    for j, w in enumerate(ws):
        y.append(np.asarray(zs[j], dtype=np.complex128))
        while len(y) > w.shape[-1]:
            y.pop(0)
        mis = w.shape[-1] - len(y)
        alm = sum(multalm(z, w[..., i + mis]) for i, z in enumerate(y))
        alm = glass_to_healpix_alm(alm)
        alm[:n] = alm[:n].real + alm[:n].imag + 0j
        alms.append(alm)
    return alms


def lognormal(x, var, lamda=1.0):
    """glass/grf/_transformations.py:83-89."""
    x = np.expm1(x - var / 2)
    return lamda * x if lamda != 1.0 else x


def squared_normal(x, a, lamda=1.0):
    """glass/grf/_transformations.py:170-175."""
    x = (x - a) ** 2 - 1
    return lamda * x if lamda != 1.0 else x


def generate(transforms, gls, nside, ncorr, zs):
    """glass/fields.py:884-894 with fields.py:429 -> oracle alm2map.
    transforms: list of ('normal',) | ('lognormal', lamda) | ('squared', a, lamda)."""
    alms = generate_alms(gls, ncorr, zs)
    out = []
    for i, (t, alm) in enumerate(zip(transforms, alms)):
        x = H.alm2map(alm, nside)
        var = cltovar(getcl(gls, i, i))
        if t[0] == "lognormal":
            x = lognormal(x, var, t[1])
        elif t[0] == "squared":
            x = squared_normal(x, t[1], t[2])
        out.append(x)
    return out


# ----------------------------------------------------------------------------
# points
# ----------------------------------------------------------------------------


def linear_bias(delta, b):
    """glass/points.py:134."""
    return b * delta


def loglinear_bias(delta, b):
    """glass/points.py:157-160."""
    d = np.log1p(delta)
    d *= b
    return np.expm1(d)


def expected_count(delta, ngal, bias=None, vis=None, bias_model="linear", remove_monopole=False):
    """glass/points.py:243-249, 279-288, 314-316 for ONE population (1-D delta)."""
    delta = np.asarray(delta, dtype=np.float64)
    if bias is None:
        n = delta.copy()
    elif bias_model == "linear":
        n = linear_bias(delta, bias)
    elif bias_model == "loglinear":
        n = loglinear_bias(delta, bias)
    else:
        n = bias_model(delta, bias)
    if remove_monopole:
        n = n - np.mean(n, keepdims=True)
    n = n + 1
    n *= ARCMIN2_SPHERE / n.size * ngal
    if vis is not None:
        n *= vis
    return n


def batch_cuts(n, batch):
    """
    The (start, stop) pixel ranges of glass/points.py:409-437 for an int count map n:
    1000-pixel stepping, searchsorted(side='right'), 'first pixel alone' rule.
    """
    n = np.asarray(n)
    npix = n.shape[-1]
    count = int(np.sum(n))
    cuts = []
    step = 1000
    start = stop = size = 0
    while count:
        q = np.cumsum(n[stop : min(npix, stop + step)])
        if size + q[-1] < min(batch, count):
            stop += step
            size += q[-1]
        else:
            stop += int(np.searchsorted(q, batch - size, side="right"))
            if stop == start:
                stop += 1
            tot = int(np.sum(n[start:stop]))
            cuts.append((start, stop, tot))
            start, size = stop, 0
            count -= tot
    assert np.sum(n[stop:]) == 0
    return cuts


def positions_from_counts(n, nside, batch, uv):
    """
    glass/points.py:389-440 for one population: yields (lon, lat, count) per batch.
    ``uv(k)`` returns the (u, v) in-pixel offsets for a batch of k points (the reference
    draws them from a fresh default_rng(42) per batch, glass/healpix.py:430).
    """
    out = []
    for start, stop, tot in batch_cuts(n, batch):
        ipix = np.repeat(np.arange(start, stop), n[start:stop])
        u, v = uv(ipix.size)
        lon, lat = H.ring2ang_uv(nside, ipix, u, v, lonlat=True)
        out.append((lon, lat, ipix.size))
    return out


# ----------------------------------------------------------------------------
# lensing
# ----------------------------------------------------------------------------


class MultiPlaneConvergence:
    """glass/lensing.py:431-606 with an explicit cosmology object (duck-typed)."""

    def __init__(self, cosmo):
        self.cosmo = cosmo
        self.z2 = 0.0
        self.z3 = 0.0
        self.x3 = 0.0
        self.w3 = 0.0
        self.r23 = 1.0
        self.delta3 = np.asarray(0.0)
        self.kappa2 = None
        self.kappa3 = None

    def add_window(self, delta, za, wa, zeff):
        """glass/lensing.py:504-509."""
        lens_weight = float(np.trapezoid(wa, za) / np.interp(zeff, za, wa))
        self.add_plane(delta, zeff, lens_weight)

    def add_plane(self, delta, zsrc, wlens=1.0):
        """glass/lensing.py:535-586."""
        if zsrc <= self.z3:
            raise ValueError("source redshift must be increasing")
        delta2, self.delta3 = self.delta3, delta
        z1, self.z2, self.z3 = self.z2, self.z3, zsrc
        w2, self.w3 = self.w3, wlens
        x2, self.x3 = self.x3, self.cosmo.transverse_comoving_distance(self.z3) / self.cosmo.hubble_distance
        r12 = self.r23
        r13, self.r23 = self.cosmo.transverse_comoving_distance([z1, self.z2], self.z3) / (
            self.cosmo.hubble_distance * self.x3
        )
        t = r13 / r12
        f = 3 * self.cosmo.Omega_m0 / 2
        f *= x2 * self.r23
        f *= (1 + self.z2) / self.cosmo.H_over_H0(self.z2)
        f *= w2
        if self.kappa2 is None:
            self.kappa2 = np.zeros_like(delta)
            self.kappa3 = np.zeros_like(delta)
        self.kappa2, self.kappa3 = self.kappa3, self.kappa2
        self.kappa3 *= 1 - t
        self.kappa3 += t * self.kappa2
        self.kappa3 += f * delta2
        return t, f

    @property
    def kappa(self):
        return self.kappa3


def uniform_positions_from_uniforms(counts, u_lon, u_lat):
    """glass/points.py:586-604 with supplied deviates: for every population k (C order) of the
    count array, lon = uniform(-180, 180) = -180 + 360 u (NumPy's low + (high - low) * random()),
    lat = degrees(arcsin(uniform(-1, 1))); yields (lon, lat, count) like the reference."""
    counts = np.asarray(counts, dtype=np.int64)
    dims = counts.shape
    pos = 0
    for k in np.ndindex(dims):
        n = int(counts[k])
        lon = -180.0 + 360.0 * np.asarray(u_lon[pos : pos + n], dtype=np.float64)
        lat = np.degrees(np.arcsin(-1.0 + 2.0 * np.asarray(u_lat[pos : pos + n], dtype=np.float64)))
        pos += n
        if dims:
            count = np.zeros(dims, dtype=np.int64)
            count[k] = n
        else:
            count = n
        yield lon, lat, count


def kappa_to_shear_fl(lmax, discretized=False, pw0=None, pw2=None):
    """glass/lensing.py:413-421: combined kappa_lm -> gamma E-mode factor."""
    ell = np.arange(lmax + 1)
    fl = np.sqrt((ell + 2) * (ell + 1) * ell * (ell - 1))
    fl /= np.clip(ell * (ell + 1), 1, None)
    fl *= -1
    if discretized:
        fl *= pw2 / pw0
    return fl


def from_convergence_factors(lmax, discretized=False, pw0=None, pw2=None):
    """glass/lensing.py:316-363: the three successive almxfl factors."""
    ell = np.arange(lmax + 1, dtype=np.float64)
    with np.errstate(divide="ignore"):
        f_psi = np.where(ell > 0, -2.0 / np.where(ell > 0, ell * (ell + 1), 1.0), 0.0)
    f_alpha = np.sqrt(ell * (ell + 1))
    f_gamma = np.where(ell > 0, np.sqrt(np.maximum((ell - 1) * (ell + 2), 0.0)), 0.0) / 2
    if discretized:
        f_gamma = f_gamma * (pw2 / pw0)
    return f_psi, f_alpha, f_gamma


def from_convergence(kappa, lmax=None, potential=False, deflection=False, shear=False, niter=3, ring_w=None):
    """glass/lensing.py:296-371 with discretized=False (no pixel-window data offline)."""
    if not (potential or deflection or shear):
        return ()
    nside = H.npix2nside(np.asarray(kappa).shape[-1])
    if lmax is None:
        lmax = 3 * nside - 1
    alm = H.map2alm(kappa, lmax=lmax, niter=niter, ring_w=ring_w)
    f_psi, f_alpha, f_gamma = from_convergence_factors(lmax)
    res = ()
    alm = H.almxfl(alm, f_psi)
    if potential:
        res += (H.alm2map(alm, nside, lmax),)
    if not (deflection or shear):
        return res
    blm = np.zeros_like(alm)
    alm = H.almxfl(alm, f_alpha)
    if deflection:
        a1, a2 = H.alm2map_spin(alm, blm, nside, 1, lmax)
        res += (a1 + 1j * a2,)
    if not shear:
        return res
    alm = H.almxfl(alm, f_gamma)
    g1, g2 = H.alm2map_spin(alm, blm, nside, 2, lmax)
    res += (g1 + 1j * g2,)
    return res


def shear_from_convergence(kappa, lmax=None, niter=3, ring_w=None, pixwin=None):
    """glass/lensing.py:403-428; ``pixwin=(pw0, pw2)``: the discretized=True branch, lensing.py:420-422
    (``fl *= pw2 / pw0`` with the tables hp.pixwin would return), None: discretized=False."""
    nside = H.npix2nside(np.asarray(kappa).shape[-1])
    if lmax is None:
        lmax = 3 * nside - 1
    alm = H.map2alm(kappa, lmax=lmax, niter=niter, ring_w=ring_w)
    blm = np.zeros_like(alm)
    fl = kappa_to_shear_fl(lmax)
    if pixwin is not None:
        pw0, pw2 = (np.asarray(p, dtype=np.float64)[: lmax + 1] for p in pixwin)
        fl = fl * (pw2 / pw0)
    alm = H.almxfl(alm, fl)
    return list(H.alm2map_spin(alm, blm, nside, 2, lmax))


# ----------------------------------------------------------------------------
# galaxies / shapes
# ----------------------------------------------------------------------------


def galaxy_shear(lon, lat, eps, kappa, gamma1, gamma2, reduced_shear=True):
    """glass/galaxies.py:311-347 (the 10 000-galaxy chunking does not change results)."""
    nside = H.npix2nside(np.broadcast_arrays(kappa, gamma1, gamma2)[0].shape[-1])
    ipix = H.ang2pix(nside, np.asarray(lon), np.asarray(lat), lonlat=True)
    k = np.asarray(kappa)[ipix]
    g = np.asarray(gamma1)[ipix] + 1j * np.asarray(gamma2)[ipix]
    if reduced_shear:
        g = g / (1 - k)
        g = (eps + g) / (1 + np.conj(g) * eps)
    else:
        g = g + eps
    return g


def ellipticity_intnorm_from_normals(sigma, normals):
    """glass/shapes.py:323-362 for one population; ``normals`` = complex standard normal
    deviates (shapes.py:48-53 draws real and imaginary parts)."""
    if not (0 <= sigma < 0.5**0.5):
        raise ValueError("sigma must be between 0 and sqrt(0.5)")
    sigma_eta = sigma * ((8 + 5 * sigma**2) / (2 - 4 * sigma**2)) ** 0.5
    e = np.asarray(normals, dtype=np.complex128) * sigma_eta
    r = np.hypot(e.real, e.imag)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = e * np.where(r > 0, np.tanh(r / 2) / r, 1.0)
    return e


def redshifts_from_nz_uniform(z, nz, u):
    """glass/galaxies.py:77-89: inverse-CDF sampling given uniforms u."""
    z = np.asarray(z, dtype=np.float64)
    nz = np.asarray(nz, dtype=np.float64)
    cdf = np.concatenate([[0.0], np.cumsum((nz[1:] + nz[:-1]) * 0.5 * np.diff(z))])  # arraytools.py:197-226
    cdf /= cdf[-1]
    return np.interp(u, cdf, z)


def gaussian_phz_from_normals(z, sigma_0, normals, lower=None, upper=None):
    """glass/galaxies.py:411-455 for array z with supplied standard normals: ``normals`` is the
    list of full-size arrays the reference's successive ``xrng.normal(z, sigma)`` calls would be
    built from (normal(loc, scale) = loc + scale * standard_normal)."""
    z = np.asarray(z, dtype=np.float64)
    sigma = (1 + z) * np.asarray(sigma_0, dtype=np.float64)
    lo = np.asarray(0.0 if lower is None else lower, dtype=np.float64)
    hi = np.asarray(np.inf if upper is None else upper, dtype=np.float64)
    rounds = iter(normals)
    zphot = z + sigma * next(rounds)
    trunc = (zphot < lo) | (zphot > hi)
    while np.count_nonzero(trunc) > 0:
        zphot = np.where(trunc, z + sigma * next(rounds), zphot)
        trunc = (zphot < lo) | (zphot > hi)
    return zphot
