"""GPU checks at BASELINE.json's FULL sizes (nside 4096, lmax 8191), where the oracle cannot run:
size-independent properties and closed forms.

* synthesis of single harmonics against lambda_lm evaluated ring by ring in 80-bit arithmetic
  on exact ring geometry with an exponent-tracked standard recurrence (independent of the
  kernel's recurrence) and the exact pixel azimuths -- covers l up to lmax, zonal, sectoral and mixed modes, the range
  machinery (lambda_mm ~ 1e-30000 near the poles) and the Bluestein / power-of-two ring FFTs;
* linearity of the batched synthesis; band-limited round trip S(A(S a)) = S a;
* galaxy counts: offsets are the exclusive scan of the counts, positions come out sorted by
  pixel and fall into their own pixel, totals follow the Poisson expectation;
* the multi-plane update against separately rounded torch arithmetic (bit-exact);
* two correlated lognormal shells through ``generate``: moments and cross-correlation.
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import healpix_ref as H

pytestmark = pytest.mark.gpu

NSIDE, LMAX = 4096, 8191


LD = np.longdouble  # 80-bit on x86-64


def exact_ring_geometry(nside):
    """cos(theta), sin(theta) of all rings in extended precision straight from the integers of
    the HEALPix definition -- no rounded double z in between."""
    n = int(nside)
    i = np.arange(1, 4 * n, dtype=np.int64)
    ip = np.where(i > 3 * n, 4 * n - i, i).astype(LD)
    cap = (i < n) | (i > 3 * n)
    t = ip * ip / (LD(3) * n * n)
    zeq = (LD(2) * n - i.astype(LD)) * 2 / (LD(3) * n)
    z = np.where(cap, np.where(i > 3 * n, t - 1, 1 - t), zeq)
    sth = np.where(cap, np.sqrt(np.abs(t * (2 - t))), np.sqrt(np.abs((1 - zeq) * (1 + zeq))))
    return z, sth


def lam_single(l, m, z, sth):
    """lambda_lm(theta) on all rings: standard three-term recurrence in l (SURVEY.md appendix
    A.2) on (mantissa, binary exponent) pairs so that sin^m(theta) may underflow; computed in
    the precision of ``z`` (the full-size test passes 80-bit geometry: in double the rounding of
    z alone moves lambda_l0 near the poles by l^2/2 * 1e-16 = 2e-9 at l = 8191)."""
    T = z.dtype.type
    k = np.arange(1, m + 1, dtype=z.dtype)
    pi = T("3.14159265358979323846264338327950288") if T is LD else T(math.pi)
    logc = T(0.5) * (np.sum(np.log1p(T(0.5) / k)) - np.log(T(4) * pi))
    ln2 = np.log(T(2))
    logseed = logc + m * np.log(sth)  # natural log of |lambda_mm|
    e = np.floor(logseed / ln2)
    p = np.exp(logseed - e * ln2) * T((-1.0) ** m)
    e = e.astype(np.int64)
    pp = np.zeros_like(p)
    for ll in range(m + 1, l + 1):
        a = np.sqrt((T(4) * ll * ll - 1) / (T(ll) * ll - T(m) * m))
        b = np.sqrt(((T(ll) - 1) ** 2 - T(m) * m) / (4 * (T(ll) - 1) ** 2 - 1)) if ll > m + 1 else T(0)
        pp, p = p, a * (z * p - b * pp)
        if (ll - m) % 16 == 0:
            sh = np.frexp(np.maximum(np.abs(p), np.abs(pp)))[1]
            p, pp = np.ldexp(p, -sh), np.ldexp(pp, -sh)
            e += sh
    return np.ldexp(p, np.clip(e, -2000, 2000).astype(np.int32)).astype(np.float64)


def _single_harmonics_error(dev, nside, lmax, modes, nmaps=1):
    """Synthesis of a few single harmonics at (nside, lmax) against lambda_lm evaluated ring by ring in
    80-bit arithmetic on exact ring geometry; returns the largest error relative to the map maximum.
    nmaps > 1: the modes are dealt over that many maps of ONE batched call (mode i in map i % nmaps)."""
    from glass_b200 import _lib
    from glass_b200.healpix import alm2map_batch, get_plan

    ri = H.ring_info(nside)
    zx, sx = exact_ring_geometry(nside)
    alm = np.zeros((nmaps, H.alm_size(lmax)), dtype=np.complex128)
    for i, (l, m, a) in enumerate(modes):
        alm[i % nmaps, H.alm_index(lmax, l, m)] = a
    got = alm2map_batch(torch.as_tensor(alm).to(dev), nside, lmax)
    # like libsharp2 (sharp_get_mlim) the transform skips m > mlim(ring) = lmax sin(theta) + max(100,
    # lmax / 100): apply the same rule to the reference and bound what it removes
    pl = get_plan(nside, lmax, 1, dev)
    mlim = (C.c_int * (2 * nside))()
    _lib.check(pl.lib.glb_debug_mlim(pl.handle, mlim), "mlim")
    mlim = np.array(mlim[:])
    nring = 4 * nside - 1
    mlim_ring = np.array([mlim[r if r < 2 * nside else nring - 1 - r] for r in range(nring)])
    # expected map, assembled on the device from per-ring factors
    nphi = torch.as_tensor(ri["nphi"], device=dev)
    ring = torch.repeat_interleave(torch.arange(nphi.numel(), device=dev), nphi)
    j = torch.arange(12 * nside * nside, device=dev) - torch.as_tensor(ri["start"], device=dev)[ring]
    nphi_p = nphi[ring]
    shifted = torch.as_tensor(ri["shifted"].astype(np.int64), device=dev)[ring]
    worst = 0.0
    for b in range(nmaps):
        want = torch.zeros(12 * nside * nside, dtype=torch.float64, device=dev)
        for l, m, a in modes[b::nmaps]:
            lam_np = lam_single(l, m, zx, sx)
            cut = mlim_ring < m
            if cut.any():
                assert np.abs(lam_np[cut]).max() < 1e-8 * np.abs(lam_np).max()  # size of the truncation
                lam_np = np.where(cut, 0.0, lam_np)
            lam = torch.as_tensor(lam_np, device=dev)[ring]
            if m == 0:
                want += a.real * lam
                continue
            # m phi_j = pi (2 m j + m shifted) / nphi, reduced exactly in integers modulo 2 nphi
            num = (2 * m * j + m * shifted) % (2 * nphi_p)
            ang = math.pi * num.to(torch.float64) / nphi_p.to(torch.float64)
            want += 2.0 * lam * (a.real * torch.cos(ang) - a.imag * torch.sin(ang))
        if len(modes[b::nmaps]):
            worst = max(worst, (got[b] - want).abs().max().item() / want.abs().max().item())
    return worst


def test_single_harmonics_fullsize(cuda_device):
    """Measured on B200: every mode within 9e-12 of the 80-bit exact-geometry value (2e-9 for the
    zonal modes before the recurrence variable was chosen per warp and the coefficient tables
    were computed in double-double, see csrc/sht_tables.cuh)."""
    modes = [(0, 0, 1.0 + 0j), (1, 0, -0.7 + 0j), (8000, 0, 0.9 + 0j), (8191, 0, 0.4 + 0j), (8191, 8191, 1.1 + 0.5j),
             (5000, 3000, -0.6 + 0.8j), (8191, 4000, 0.5 + 0.1j), (7000, 6999, 0.2 - 0.9j), (6001, 17, 0.3 + 0.3j)]
    err = _single_harmonics_error(cuda_device, NSIDE, LMAX, modes)
    assert err < 1e-10, err  # BASELINE.json north_star: maps from identical alm within 1e-10 relative


def exact_half_angles(nside):
    """cos(theta/2), sin(theta/2) of all rings in extended precision (1 - z is exact in the caps)."""
    n = int(nside)
    i = np.arange(1, 4 * n, dtype=np.int64)
    ip = np.where(i > 3 * n, 4 * n - i, i).astype(LD)
    cap = (i < n) | (i > 3 * n)
    t = ip * ip / (LD(3) * n * n)
    zeq = (LD(2) * n - i.astype(LD)) * 2 / (LD(3) * n)
    omz = np.where(cap, np.where(i > 3 * n, 2 - t, t), 1 - zeq)  # 1 - z
    opz = np.where(cap, np.where(i > 3 * n, t, 2 - t), 1 + zeq)  # 1 + z
    return np.sqrt(opz / 2), np.sqrt(omz / 2)


def wigner_d_single(l, m, mp, z, ch, sh):
    """d^l_{m,mp}(theta) on all rings in the precision of ``z``: oracle.healpix_ref.wigner_d_l
    (three-term recurrence in l) keeping two rows.  80-bit range covers sin^(m+s)(theta/2) down
    to 1e-4900, enough for the small m used here."""
    T = z.dtype.type
    j = max(abs(m), abs(mp))
    a, b, sign = m, mp, 1.0
    if abs(a) < abs(b):
        a, b = b, a
        sign *= (-1.0) ** (a - b)
    if a < 0:
        sign *= (-1.0) ** (a - b)
        a, b = -a, -b
    lognorm = T(0.5) * (T(math.lgamma(2 * j + 1.0)) - T(math.lgamma(j + b + 1.0)) - T(math.lgamma(j - b + 1.0)))
    # exact norm for small j: sqrt((2j)! / ((j+b)! (j-b)!))
    if j <= 20:
        lognorm = T(0.5) * np.log(T(math.factorial(2 * j)) / (T(math.factorial(j + b)) * T(math.factorial(j - b))))
    with np.errstate(divide="ignore"):
        mag = np.exp(lognorm + (j + b) * np.log(ch) + (j - b) * np.log(sh))
    cur = T(sign * (-1.0) ** (j - b)) * mag
    if l == j:
        return cur
    prev = np.zeros_like(cur)
    for ll in range(j, l):
        den = T(ll) * np.sqrt(((T(ll) + 1) ** 2 - T(m) * m) * ((T(ll) + 1) ** 2 - T(mp) * mp)) if ll > 0 else T(1)
        if ll == 0:
            nxt = z * cur
        else:
            t1 = (2 * T(ll) + 1) * (T(ll) * (T(ll) + 1) * z - T(m) * mp)
            t2 = (T(ll) + 1) * np.sqrt((T(ll) * ll - T(m) * m) * (T(ll) * ll - T(mp) * mp))
            nxt = (t1 * cur - t2 * prev) / den
        prev, cur = cur, nxt
    return cur


def slam_single(l, m, s, z, ch, sh):
    """sY_lm(theta, 0) = (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta) (oracle.healpix_ref.slam_lm)."""
    T = z.dtype.type
    pi = T("3.14159265358979323846264338327950288") if T is LD else T(math.pi)
    return T((-1.0) ** s) * np.sqrt((2 * T(l) + 1) / (4 * pi)) * wigner_d_single(l, m, -s, z, ch, sh)


def spin_mode_maps(modes, spin, nside, ring, j, nphi_p, shifted, z, ch, sh, dev, mlim_ring=None):
    """Expected (map1, map2) of E-only single modes [(l, m, e)]: P = map1 + i map2 =
    sum_m P_m e^{i m phi} with P_m = -e sY_{l,m}, P_{-m} = -(-1)^m conj(e) sY_{l,-m}
    (oracle.healpix_ref.alm2map_spin)."""
    npix = 12 * nside * nside
    w1 = torch.zeros(npix, dtype=torch.float64, device=dev)
    w2 = torch.zeros(npix, dtype=torch.float64, device=dev)
    for l, m, e in modes:
        yp = slam_single(l, m, spin, z, ch, sh).astype(np.float64)
        if mlim_ring is not None:
            yp = np.where(mlim_ring < m, 0.0, yp)
        yp_t = torch.as_tensor(yp, device=dev)[ring]
        if m == 0:
            w1 += -e.real * yp_t
            continue
        yn = slam_single(l, -m, spin, z, ch, sh).astype(np.float64)
        if mlim_ring is not None:
            yn = np.where(mlim_ring < m, 0.0, yn)
        yn_t = torch.as_tensor(yn, device=dev)[ring]
        num = (2 * m * j + m * shifted) % (2 * nphi_p)
        ang = math.pi * num.to(torch.float64) / nphi_p.to(torch.float64)
        c, sn = torch.cos(ang), torch.sin(ang)
        sg = (-1.0) ** m
        # P_m e^{i a} + P_{-m} e^{-i a},  P_m = -(er + i ei) yp,  P_{-m} = -sg (er - i ei) yn
        w1 += -(e.real * c - e.imag * sn) * yp_t - sg * (e.real * c - e.imag * sn) * yn_t
        w2 += -(e.real * sn + e.imag * c) * yp_t + sg * (e.real * sn + e.imag * c) * yn_t
    return w1, w2


def test_spin2_single_harmonics_fullsize(cuda_device):
    """Spin-2 synthesis (kappa -> shear) at nside 4096 against 80-bit spin-weighted harmonics on
    exact ring geometry: low m, where the polar rings amplify rounding of cos(theta) by l^2/2."""
    from glass_b200 import healpix as hp

    dev = cuda_device
    ri = H.ring_info(NSIDE)
    zx, _ = exact_ring_geometry(NSIDE)
    chx, shx = exact_half_angles(NSIDE)
    modes = [(2, 0, 1.0 + 0j), (8191, 0, 0.8 + 0j), (8000, 1, 0.5 - 0.4j), (8191, 2, -0.7 + 0.2j), (6000, 3, 0.3 + 0.9j), (4097, 40, 0.6 - 0.1j)]
    alm = np.zeros(H.alm_size(LMAX), dtype=np.complex128)
    for l, m, e in modes:
        alm[H.alm_index(LMAX, l, m)] = e
    g1, g2 = hp.alm2map_spin([torch.as_tensor(alm).to(dev), None], NSIDE, 2, LMAX)
    nphi = torch.as_tensor(ri["nphi"], device=dev)
    ring = torch.repeat_interleave(torch.arange(nphi.numel(), device=dev), nphi)
    j = torch.arange(12 * NSIDE * NSIDE, device=dev) - torch.as_tensor(ri["start"], device=dev)[ring]
    nphi_p = nphi[ring]
    shifted = torch.as_tensor(ri["shifted"].astype(np.int64), device=dev)[ring]
    w1, w2 = spin_mode_maps(modes, 2, NSIDE, ring, j, nphi_p, shifted, zx, chx, shx, dev)
    scale = max(w1.abs().max().item(), w2.abs().max().item())
    err = max((g1 - w1).abs().max().item(), (g2 - w2).abs().max().item()) / scale
    assert err < 1e-10, err


def test_linearity_and_batching_fullsize(cuda_device):
    """S(a) + S(b) == S(a + b) to rounding, and a map does not depend on its batch slot."""
    from glass_b200.healpix import alm2map_batch

    g = torch.Generator(device=cuda_device)
    g.manual_seed(5)
    n = H.alm_size(LMAX)
    ab = torch.view_as_complex(torch.randn((2, n, 2), dtype=torch.float64, device=cuda_device, generator=g))
    ab[:, : LMAX + 1] = ab[:, : LMAX + 1].real.to(torch.complex128)  # the m = 0 block comes first
    ell = np.concatenate([np.arange(m, LMAX + 1) for m in range(LMAX + 1)])
    ab = ab * torch.as_tensor((ell + 1.0) ** -1.25, device=cuda_device)
    stack = torch.stack([ab[0], ab[1], ab[0] + ab[1], ab[1]])
    maps = alm2map_batch(stack, NSIDE, LMAX)
    ref = maps[2].abs().max().item()
    assert (maps[0] + maps[1] - maps[2]).abs().max().item() < 1e-11 * ref
    assert torch.equal(maps[1], maps[3])
    single = alm2map_batch(ab[1:2], NSIDE, LMAX)[0]
    assert (single - maps[1]).abs().max().item() < 1e-11 * ref


def test_roundtrip_band_limited_fullsize(cuda_device):
    """S(A(S(alm))) == S(alm) at nside 4096 for input band-limited to lmax = nside."""
    from glass_b200.healpix import alm2map_batch, clear_plans, map2alm

    lmax = NSIDE
    g = torch.Generator(device=cuda_device)
    g.manual_seed(6)
    n = H.alm_size(lmax)
    alm = torch.view_as_complex(torch.randn((1, n, 2), dtype=torch.float64, device=cuda_device, generator=g))
    alm[:, : lmax + 1] = alm[:, : lmax + 1].real.to(torch.complex128)
    m1 = alm2map_batch(alm, NSIDE, lmax)[0]
    a2 = map2alm(m1, lmax=lmax, pol=False, niter=3)
    assert (a2 - alm[0]).abs().max().item() < 1e-6 * alm.abs().max().item()
    m2 = alm2map_batch(a2[None], NSIDE, lmax)[0]
    assert (m2 - m1).abs().max().item() < 1e-6 * m1.abs().max().item()
    clear_plans()


def test_points_fullsize_properties(cuda_device):
    """nside 4096, 0.083 galaxies per pixel (1e9 galaxies over 60 shells): offsets == exclusive
    scan of counts, positions sorted by pixel and inside their pixel, Poisson total."""
    from glass_b200 import _lib
    from glass_b200 import healpix as hp

    dev = cuda_device
    lib = _lib.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    npix = 12 * NSIDE * NSIDE
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    delta = torch.expm1(0.5 * torch.randn(npix, dtype=torch.float64, device=dev, generator=g) - 0.125)
    counts = torch.empty(npix, dtype=torch.int64, device=dev)
    off = torch.empty(npix + 1, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.glb_points_workspace_bytes(npix)), dtype=torch.uint8, device=dev)
    bias, scale = 1.2, 0.083
    tot_d = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.check(lib.glb_points_counts(npix, delta.data_ptr(), None, 1, bias, scale, 0, None, C.c_uint64(42), C.c_uint32(0), None,
                                     counts.data_ptr(), off.data_ptr(), None, 0, tot_d.data_ptr(), ws.data_ptr(), st))
    torch.cuda.synchronize()
    assert int(tot_d) == int(off[-1])
    assert int(counts.min()) >= 0
    assert int(off[0]) == 0
    assert torch.equal(off[1:], torch.cumsum(counts, 0))
    tot = int(off[-1])
    lam = torch.clamp((bias * delta + 1.0) * scale, min=0.0)
    mean, sd = float(lam.sum()), math.sqrt(float(lam.sum()))
    assert abs(tot - mean) < 6 * sd, (tot, mean, sd)
    # zero-inflation matches Poisson: P(0) = mean of exp(-lambda)
    p0 = float(torch.exp(-lam).mean())
    f0 = float((counts == 0).double().mean())
    assert abs(f0 - p0) < 6 * math.sqrt(p0 * (1 - p0) / npix)
    lon = torch.empty(tot, dtype=torch.float64, device=dev)
    lat = torch.empty(tot, dtype=torch.float64, device=dev)
    _lib.check(lib.glb_points_fill(NSIDE, counts.data_ptr(), off.data_ptr(), 0, npix, None, None, C.c_uint64(42), C.c_uint32(0),
                                   lon.data_ptr(), lat.data_ptr(), None, st))
    torch.cuda.synchronize()
    ipix = hp.ang2pix(NSIDE, lon, lat, lonlat=True)
    want = torch.repeat_interleave(torch.arange(npix, device=dev), counts)
    assert torch.equal(ipix, want)  # sorted by ring pixel like the reference, every galaxy in its own pixel
    assert float(lon.min()) >= 0.0 and float(lon.max()) < 360.0 and float(lat.abs().max()) <= 90.0
    # LIST mode (what positions_from_delta uses for sparse maps): the galaxy -> pixel list alone, first
    # with a capacity that is too small (total still exact, prefix written), then one that fits; the
    # list is np.repeat(arange, counts) and the positions are bit-identical to the scan path's
    small = tot // 2
    gpix = torch.full((tot + 8,), -1, dtype=torch.int64, device=dev)
    _lib.check(lib.glb_points_counts(npix, delta.data_ptr(), None, 1, bias, scale, 0, None, C.c_uint64(42), C.c_uint32(0), None,
                                     None, None, gpix.data_ptr(), small, tot_d.data_ptr(), ws.data_ptr(), st))
    torch.cuda.synchronize()
    assert int(tot_d) == tot and torch.equal(gpix[:small], want[:small]) and bool((gpix[small:] == -1).all())
    _lib.check(lib.glb_points_counts(npix, delta.data_ptr(), None, 1, bias, scale, 0, None, C.c_uint64(42), C.c_uint32(0), None,
                                     None, None, gpix.data_ptr(), tot + 8, tot_d.data_ptr(), ws.data_ptr(), st))
    torch.cuda.synchronize()
    assert int(tot_d) == tot and torch.equal(gpix[:tot], want) and bool((gpix[tot:] == -1).all())
    lon2, lat2 = torch.empty_like(lon), torch.empty_like(lat)
    _lib.check(lib.glb_points_fill_list(NSIDE, gpix.data_ptr(), 0, tot, None, None, C.c_uint64(42), C.c_uint32(0), lon2.data_ptr(),
                                        lat2.data_ptr(), st))
    g0 = tot // 3  # a sub-range: indices are global, outputs relative to g0
    lon3, lat3 = torch.empty(1000, dtype=torch.float64, device=dev), torch.empty(1000, dtype=torch.float64, device=dev)
    _lib.check(lib.glb_points_fill_list(NSIDE, gpix.data_ptr(), g0, g0 + 1000, None, None, C.c_uint64(42), C.c_uint32(0),
                                        lon3.data_ptr(), lat3.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(lon2, lon) and torch.equal(lat2, lat)
    assert torch.equal(lon3, lon[g0 : g0 + 1000]) and torch.equal(lat3, lat[g0 : g0 + 1000])


def test_multiplane_update_fullsize_bit_exact(cuda_device):
    from glass_b200 import _lib

    dev = cuda_device
    lib = _lib.load()
    st = torch.cuda.current_stream(dev).cuda_stream
    npix = 12 * NSIDE * NSIDE
    g = torch.Generator(device=dev)
    g.manual_seed(8)
    k3 = torch.randn(npix, dtype=torch.float64, device=dev, generator=g)
    k2 = torch.randn(npix, dtype=torch.float64, device=dev, generator=g)
    d2 = torch.randn(npix, dtype=torch.float64, device=dev, generator=g)
    t, f = 1.37, 0.0123
    want = k3 * (1 - t)  # glass/lensing.py:584-586: three separately rounded NumPy passes
    want += t * k2
    want += f * d2
    _lib.check(lib.glb_multiplane_update(k3.data_ptr(), k2.data_ptr(), d2.data_ptr(), 0.0, npix, t, f, st))
    torch.cuda.synchronize()
    assert torch.equal(k3, want)


def test_generate_two_lognormal_shells_fullsize(cuda_device):
    """Two correlated lognormal shells at nside 4096 through the public API: zero mean, variance
    e^{var} - 1 and the cross-correlation the Gaussian cross-spectrum implies (statistical)."""
    import glass_b200
    from helpers import synthetic_gls

    gls = [torch.as_tensor(x).to(cuda_device) for x in synthetic_gls(2, LMAX, 1)]
    fields = [glass_b200.grf.Lognormal(), glass_b200.grf.Lognormal()]
    maps = [m.clone() for m in glass_b200.generate(fields, gls, NSIDE, ncorr=1, rng=123)]
    assert len(maps) == 2 and all(m.shape == (12 * NSIDE * NSIDE,) for m in maps)
    l = np.arange(LMAX + 1)
    g = 1e-2 * (l + 1.0) ** -1.5
    g[0] = 0.0
    var = float(((2 * l + 1) * g).sum() / (4 * math.pi))
    for m in maps:
        assert float(m.min()) > -1.0
        assert abs(float(m.mean())) < 5e-3
        assert abs(float(m.var()) / math.expm1(var) - 1.0) < 0.05
    cov = float((maps[0] * maps[1]).mean() - maps[0].mean() * maps[1].mean())
    assert abs(cov / math.expm1(0.5 * var) - 1.0) < 0.08


def test_long_rings_nside_8192(cuda_device):
    """nside 8192: rings of up to 32768 pixels, whose FFT (16384 complex points) does not fit one SM's
    shared memory and runs on global work buffers (sht_ringfft_*_long_kernel).  With a small lmax the
    Legendre stage is cheap and the C oracle (independent FFT: radix-2 / plain Bluestein per ring) can
    check EVERY ring: synthesis with alias folding on the polar rings, the fused lognormal store, and
    the analysis direction with ring weights."""
    from glass_b200 import _lib
    from glass_b200 import healpix as hp
    from oracle import sht_c

    nside, lmax = 8192, 150
    rng = np.random.default_rng(8192)
    n = (lmax + 1) * (lmax + 2) // 2
    alm = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    alm[: lmax + 1] = alm[: lmax + 1].real
    ref = sht_c.alm2map(alm, nside, lmax)
    d_alm = torch.as_tensor(alm[None]).to(cuda_device)
    got = hp.alm2map_batch(d_alm, nside, lmax)[0]
    err = float((got - torch.as_tensor(ref, device=cuda_device)).abs().max()) / np.abs(ref).max()
    assert err < 1e-11, err
    logn = hp.alm2map_batch(0.01 * d_alm, nside, lmax, transforms=[(_lib.T_LOGNORMAL, 0.05, 0.7)])[0]
    want = 0.7 * np.expm1(0.01 * ref - 0.05)
    assert float((logn - torch.as_tensor(want, device=cuda_device)).abs().max()) < 1e-12 * np.abs(want).max() + 1e-14
    del logn
    # analysis of the synthesised map plus pixel noise, with ring weights
    mp = ref / np.abs(ref).max() + 1e-3 * rng.standard_normal(ref.size)
    w = 1.0 + 0.05 * rng.random(4 * nside - 1)
    ga = hp.map2alm(torch.as_tensor(mp, device=cuda_device), lmax=lmax, pol=False, niter=1, ring_weights=w).cpu().numpy()
    ra = sht_c.map2alm(mp, lmax, niter=1, ring_w=w)
    assert np.abs(ga - ra).max() < 1e-10 * np.abs(ra).max(), np.abs(ga - ra).max() / np.abs(ra).max()
    hp.clear_plans()
    torch.cuda.empty_cache()


def test_single_harmonics_nside_8192(cuda_device):
    """The whole transform at nside 8192, lmax 16383 -- the size the m-split axis exists for (range
    scaling down to lambda_mm ~ 1e-60000, pole skipping, long-ring FFTs): single harmonics against the
    80-bit ring-by-ring evaluation, as test_single_harmonics_fullsize does at nside 4096."""
    from glass_b200.healpix import clear_plans

    modes = [(0, 0, 1.0 + 0j), (16383, 0, 0.4 + 0j), (16383, 16383, 1.1 + 0.5j), (12001, 7000, -0.6 + 0.8j), (9000, 1, 0.3 + 0.3j),
             (16383, 8000, 0.5 + 0.1j)]
    err = _single_harmonics_error(cuda_device, 8192, 16383, modes)
    assert err < 1e-10, err
    clear_plans()
    torch.cuda.empty_cache()
