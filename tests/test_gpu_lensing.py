"""GPU parity: multi-plane convergence, galaxy_shear, ellipticities, redshifts."""
import numpy as np
import pytest
import torch

from helpers import MockCosmology, triangular_shells
from oracle import glass_ref as G
from oracle import healpix_ref as H

pytestmark = pytest.mark.gpu


def test_multi_plane_bit_exact(cuda_device):
    import glass_b200

    cosmo = MockCosmology()
    shells = triangular_shells(5)
    rng = np.random.default_rng(42)
    deltas = rng.random((5, 12 * 4**2))
    a = glass_b200.MultiPlaneConvergence(cosmo)
    b = G.MultiPlaneConvergence(cosmo)
    for i, w in enumerate(shells):
        a.add_window(deltas[i], w)
        b.add_window(deltas[i].copy(), w.za, w.wa, w.zeff)
        assert np.array_equal(a.kappa, b.kappa), i  # same roundings as NumPy's three passes
        assert a.zsrc == w.zeff
    with pytest.raises(ValueError, match="source redshift must be increasing"):
        a.add_plane(deltas[0], 1.0)
    # device tensors stay on the device
    c = glass_b200.MultiPlaneConvergence(cosmo)
    for i, w in enumerate(shells):
        c.add_window(torch.as_tensor(deltas[i]).to(cuda_device), w)
    assert c.kappa.is_cuda and np.array_equal(c.kappa.cpu().numpy(), b.kappa)


def test_multi_plane_matrix_and_weights(cuda_device):
    """Port of tests/core/test_lensing.py:64-117."""
    import glass_b200

    cosmo = MockCosmology()
    shells = triangular_shells(5)
    mat = glass_b200.multi_plane_matrix(shells, cosmo)
    assert np.array_equal(mat, np.tril(mat)) and np.all(np.triu(mat, 1) == 0)
    conv = glass_b200.MultiPlaneConvergence(cosmo)
    rng = np.random.default_rng(42)
    deltas = rng.random((5, 10))
    kappas = []
    for i in range(5):
        conv.add_window(deltas[i], shells[i])
        kappas.append(np.array(conv.kappa, copy=True))
    np.testing.assert_allclose(mat @ deltas, np.stack(kappas), rtol=1e-12, atol=1e-14)
    w_out = glass_b200.multi_plane_weights(np.eye(5), shells, cosmo)
    assert np.array_equal(w_out, np.triu(w_out, 1))
    weights = rng.random((5, 3))
    conv = glass_b200.MultiPlaneConvergence(cosmo)
    kappa = 0
    for i in range(5):
        conv.add_window(deltas[i], shells[i])
        kappa = kappa + weights[i][..., None] * conv.kappa
    kappa /= weights.sum(axis=0)[..., None]
    wmat = glass_b200.multi_plane_weights(weights, shells, cosmo)
    np.testing.assert_allclose(np.einsum("ij,ik", wmat, deltas), kappa, rtol=1e-12)
    with pytest.raises(ValueError, match="shape mismatch between weights and shells"):
        glass_b200.multi_plane_weights(np.ones((4, 2)), shells, cosmo)


@pytest.mark.parametrize("reduced", [True, False])
def test_galaxy_shear_vs_oracle(cuda_device, reduced):
    import glass_b200

    nside = 16
    npix = 12 * nside**2
    rng = np.random.default_rng(1)
    kappa, g1, g2 = 0.1 * rng.standard_normal((3, npix))
    n = 5000
    ipix = rng.integers(0, npix, n)
    lon, lat = H.ring2ang_uv(nside, ipix, rng.random(n), rng.random(n), lonlat=True)
    eps = 0.3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    got = glass_b200.galaxy_shear(lon, lat, eps, kappa, g1, g2, reduced_shear=reduced)
    ref = G.galaxy_shear(lon, lat, eps, kappa, g1, g2, reduced_shear=reduced)
    assert got.dtype == np.complex128 and got.shape == (n,)
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-15)
    got2 = glass_b200.galaxy_shear(lon, lat, eps, kappa, g1, g2, reduced_shear=reduced, ipix=ipix)
    assert np.array_equal(got, got2)
    # shapes-only cases of tests/core/test_galaxies.py:175-232
    k12 = np.zeros(12)
    assert glass_b200.galaxy_shear(np.zeros(0), np.zeros(0), np.zeros(0, dtype=complex), k12, k12, k12).shape == (0,)
    assert glass_b200.galaxy_shear(np.zeros(5), np.zeros(5), np.zeros(5, dtype=complex), k12, k12, k12).shape == (5,)


def test_ellipticities(cuda_device):
    """Supplied normals -> oracle formula; Philox draws -> the reference's statistical
    checks (tests/core/test_shapes.py:130-252)."""
    import glass_b200
    from glass_b200.rng import Deviates

    rng = np.random.default_rng(2)
    nrm = rng.standard_normal(1000) + 1j * rng.standard_normal(1000)
    got = glass_b200.ellipticity_intnorm(1000, 0.256, rng=Deviates(normal=nrm))
    np.testing.assert_allclose(got, G.ellipticity_intnorm_from_normals(0.256, nrm), rtol=1e-14, atol=1e-16)
    n = 1_000_000
    for fn in (glass_b200.ellipticity_intnorm, glass_b200.ellipticity_gaussian):
        eps = fn(n, 0.256, rng=3)
        assert eps.shape == (n,) and np.all(np.abs(eps) < 1)
        assert abs(np.std(eps.real) - 0.256) < 1e-3 and abs(np.std(eps.imag) - 0.256) < 1e-3
    eps = glass_b200.ellipticity_intnorm([n, n], [0.128, 0.256], rng=4)
    assert eps.shape == (2 * n,)
    assert abs(np.std(eps.real[:n]) - 0.128) < 1e-3 and abs(np.std(eps.real[n:]) - 0.256) < 1e-3
    with pytest.raises(ValueError, match="sigma must be between 0 and sqrt\\(0.5\\)"):
        glass_b200.ellipticity_intnorm(1, 0.71)


def test_redshifts(cuda_device):
    import glass_b200
    from glass_b200.rng import Deviates
    from scipy import stats

    z = np.linspace(0.0, 2.0, 201)
    nz = z**2 * np.exp(-((z / 0.5) ** 1.5))
    u = np.random.default_rng(0).random(10_000)
    w = glass_b200.RadialWindow(z, nz, 1.0)
    got = glass_b200.redshifts(u.size, w, rng=Deviates(uniform=u))
    np.testing.assert_allclose(got, G.redshifts_from_nz_uniform(z, nz, u), rtol=1e-13, atol=1e-15)
    zs = glass_b200.redshifts(200_000, w, rng=5)
    cdf = G.redshifts_from_nz_uniform  # noqa: F841
    cd = np.concatenate([[0], np.cumsum((nz[1:] + nz[:-1]) * 0.5 * np.diff(z))])
    cd /= cd[-1]
    assert stats.kstest(zs, lambda x: np.interp(x, z, cd)).pvalue > 1e-4
    with pytest.warns(UserWarning):
        glass_b200.redshifts_from_nz(10, z, nz, rng=1)
    out = glass_b200.redshifts_from_nz(np.array([3, 4]), z, np.stack([nz, nz[::-1]]), rng=1, warn=False)
    assert out.shape == (7,)


@pytest.mark.parametrize("nside,lmax", [(8, 16), (16, None)])
def test_from_convergence_vs_oracle(cuda_device, nside, lmax):
    import glass_b200

    rng = np.random.default_rng(nside)
    kappa = 0.05 * rng.standard_normal(12 * nside**2)
    got = glass_b200.from_convergence(kappa, lmax, potential=True, deflection=True, shear=True, discretized=False)
    ref = G.from_convergence(kappa, lmax, potential=True, deflection=True, shear=True)
    assert len(got) == 3 and got[1].dtype == np.complex128 and got[2].dtype == np.complex128
    for g, r in zip(got, ref):
        assert np.abs(g - r).max() < 1e-10 * np.abs(r).max()
    # tuple lengths for the 8 flag combinations (tests/core/test_lensing.py:21-56)
    for p in (False, True):
        for d in (False, True):
            for s in (False, True):
                res = glass_b200.from_convergence(kappa, lmax, potential=p, deflection=d, shear=s, discretized=False, niter=0)
                assert len(res) == p + d + s
    g1, g2 = glass_b200.shear_from_convergence(kappa, lmax, discretized=False)
    r1, r2 = G.shear_from_convergence(kappa, lmax)
    assert np.abs(g1 - r1).max() < 1e-10 * np.abs(r1).max() and np.abs(g2 - r2).max() < 1e-10 * np.abs(r1).max()
    np.testing.assert_allclose(g1 + 1j * g2, got[2], rtol=0, atol=1e-12 * np.abs(got[2]).max())
    # the reference's canonical call: discretized=True by default, pixel windows from hp.pixwin
    # (glass/lensing.py:414-422) -- generated here (glass_b200.pixwin) instead of read from healpy's files
    lm = 3 * nside - 1 if lmax is None else lmax
    pw0, pw2 = glass_b200.healpix.pixwin(nside, lmax=lm, pol=True)
    gd = glass_b200.shear_from_convergence(kappa, lmax)
    rd = G.shear_from_convergence(kappa, lmax, pixwin=(pw0, pw2))
    assert np.abs(gd[0] - rd[0]).max() < 1e-10 * np.abs(rd[0]).max() and np.abs(gd[1] - rd[1]).max() < 1e-10 * np.abs(rd[0]).max()
    assert np.abs(gd[0] - g1).max() > 1e-6 * np.abs(g1).max()  # the window ratio does something
    fc = glass_b200.from_convergence(kappa, lmax, shear=True)  # discretized=True default (glass/lensing.py:353-363)
    np.testing.assert_allclose(fc[0], gd[0] + 1j * gd[1], rtol=0, atol=1e-12 * np.abs(rd[0]).max())
    # caller-supplied tables (e.g. healpy's own) still take precedence
    pw = (np.ones(3 * nside), np.linspace(1.0, 0.9, 3 * nside))
    gu = glass_b200.shear_from_convergence(kappa, lmax, discretized=True, pixwin=pw)
    ru = G.shear_from_convergence(kappa, lmax, pixwin=pw)
    assert np.abs(gu[0] - ru[0]).max() < 1e-10 * np.abs(ru[0]).max()


def test_pixwin_on_device(cuda_device):
    """hp.pixwin (glass/healpix.py:313-356) generated on the device: the brute-force definition at
    nside 2 (sum over m of pixel-averaged harmonics, oracle geometry), the values of the CPU suite,
    shapes / defaults / errors, the scaled range above nside 128, alm2map(pixwin=True)."""
    import glass_b200
    from glass_b200 import healpix as hp

    wt, wp = hp.pixwin(2, lmax=8, pol=True)
    want_t = [1.0, 0.977303, 0.93310702, 0.86971852, 0.79038278, 0.69905215, 0.60011811, 0.49813949, 0.39760902]
    assert np.abs(wt - want_t).max() < 1e-8 and wp[0] == wp[1] == 0.0 and np.abs(wp[2:] / wt[2:] - 1).max() < 0.09
    w = hp.pixwin(64)
    assert w.shape == (3 * 64,) and w[0] == 1.0 and np.all(np.diff(w) < 0)
    wt, wp = hp.pixwin(256, lmax=1024, pol=True)  # scaled from the nside-128 moments
    w128 = hp.pixwin(128, lmax=512)
    # (l + 1/2) / nside is the scaling variable: l = 2j + 1 at nside 256 sits at l' = j + 1/4 of nside 128
    assert np.abs(wt[1::2][:256] - (0.75 * w128[:256] + 0.25 * w128[1:257])).max() < 1e-4
    assert abs(wt[1024] - (0.25 * w128[511] + 0.75 * w128[512])) < 1e-4
    assert wp[0] == wp[1] == 0.0 and np.abs(wp[2:] / wt[2:] - 1).max() < 1e-3
    t = hp.pixwin(8, lmax=5, xp=torch)
    assert t.is_cuda and t.shape == (6,)
    with pytest.raises(ValueError, match="tabulated up to"):
        hp.pixwin(8, lmax=33)
    rng = np.random.default_rng(1)
    alm = rng.standard_normal(45) + 1j * rng.standard_normal(45)
    alm[:9] = alm[:9].real
    a = hp.alm2map(alm, 4, pixwin=True, pol=False)
    b = hp.alm2map(hp.almxfl(alm, hp.pixwin(4, lmax=8)), 4, pol=False)
    assert np.array_equal(a, b)
    # discretized_cls picks the generated window up (glass/fields.py:288-299)
    cl = np.ones(10)
    got = glass_b200.discretized_cls([cl], nside=4, lmax=8)[0]
    assert np.allclose(got, hp.pixwin(4, lmax=8) ** 2, rtol=1e-15)


def test_gaussian_phz(cuda_device):
    """glass/galaxies.py:350-455: bit-exact against the reference's golden vector with supplied
    normals (several rejection rounds), bounds / shapes / errors, Philox draws statistically."""
    import os

    from scipy import stats

    import glass_b200
    from glass_b200.rng import Deviates

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    got = glass_b200.gaussian_phz(gold["phz_z"], 0.2, lower=0.1, upper=1.2, rng=Deviates(normal=list(gold["phz_normals"])))
    assert isinstance(got, np.ndarray) and np.array_equal(got, gold["phz_out"])
    # Philox path: unbounded -> (zphot - z) / ((1 + z) sigma_0) is standard normal; default lower bound 0
    z = torch.linspace(0.2, 2.0, 200_000, dtype=torch.float64, device=cuda_device)
    zp = glass_b200.gaussian_phz(z, 0.05, rng=3)
    assert zp.is_cuda and zp.shape == z.shape and float(zp.min()) >= 0.0
    r = ((zp - z) / ((1 + z) * 0.05)).cpu().numpy()
    assert stats.kstest(r, "norm").pvalue > 1e-4
    # truncation: everything inside the bounds, per-galaxy sigma and bounds arrays
    lo, hi = torch.full_like(z, 0.5), torch.full_like(z, 1.5)
    zp = glass_b200.gaussian_phz(z, torch.full_like(z, 0.3), lower=lo, upper=hi, rng=4)
    assert float(zp.min()) >= 0.5 and float(zp.max()) <= 1.5
    # scalar in, 0-d out
    assert glass_b200.gaussian_phz(1.0, 0.0).shape == () and float(glass_b200.gaussian_phz(1.0, 0.0)) == 1.0
    with pytest.raises(ValueError, match="requires lower < upper"):
        glass_b200.gaussian_phz(z, 0.1, lower=1.0, upper=0.5)
    with pytest.raises(ValueError, match="lower and upper must best scalars"):
        glass_b200.gaussian_phz(z, 0.1, lower=torch.zeros(3, device=cuda_device), upper=torch.ones(3, device=cuda_device))


def test_batched_map2alm_and_shear_match_single(cuda_device):
    """healpy.map2alm accepts a sequence of maps; here up to four maps of a group share the
    recurrence of every refinement synthesis.  Results against the one-map-at-a-time path
    (1e-10: the batched synthesis kernel accumulates in the same order, its ring FFT too)."""
    import glass_b200
    from glass_b200 import healpix as hp

    nside, lmax = 32, 64
    rng = np.random.default_rng(12)
    maps = 0.01 * rng.standard_normal((7, 12 * nside**2))
    singles = [hp.map2alm(m, lmax=lmax, pol=False, niter=2) for m in maps]
    batched = hp.map2alm(list(maps), lmax=lmax, pol=False, niter=2)  # groups of 4 + 2 + 1
    assert len(batched) == 7
    for a, b in zip(batched, singles):
        assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max()
    ref = H.map2alm(maps[5], lmax=lmax, niter=2)
    assert np.abs(batched[5] - ref).max() <= 1e-10 * np.abs(ref).max()
    g1, g2 = glass_b200.shear_from_convergence(maps[:3], lmax, discretized=False, niter=1)
    assert g1.shape == g2.shape == (3, 12 * nside**2)
    for b in range(3):
        s1, s2 = glass_b200.shear_from_convergence(maps[b], lmax, discretized=False, niter=1)
        assert np.abs(g1[b] - s1).max() <= 1e-10 * np.abs(s1).max()
        assert np.abs(g2[b] - s2).max() <= 1e-10 * np.abs(s2).max()
    kd = torch.as_tensor(maps[:2]).to(cuda_device)
    d1, d2 = glass_b200.shear_from_convergence(kd, lmax, discretized=False, niter=1)
    assert d1.is_cuda and np.abs(d1.cpu().numpy() - g1[:2]).max() <= 1e-12 * np.abs(g1).max()


@pytest.mark.parametrize("nside,lmax,spin", [(16, 40, 2), (64, 128, 2), (32, 64, 1), (256, 511, 2)])
def test_batched_spin_synthesis(cuda_device, nside, lmax, spin, monkeypatch):
    """glb_alm2map_spin_batch: the E modes of 1..7 map pairs on shared Wigner-d recurrences (launch
    groups of 4, 2, 1 and both ring-pairs-per-thread variants of the four-map kernel) against the
    one-pair-at-a-time transform and, at small sizes, the oracle (glass/healpix.py:81-108)."""
    from glass_b200 import healpix as hp

    rng = np.random.default_rng(nside + spin)
    nalm = (lmax + 1) * (lmax + 2) // 2
    alms = rng.standard_normal((7, nalm)) + 1j * rng.standard_normal((7, nalm))
    alms[:, : lmax + 1] = alms[:, : lmax + 1].real
    ad = torch.as_tensor(alms).to(cuda_device)
    singles = [hp.alm2map_spin([ad[b], None], nside, spin, lmax) for b in range(7)]
    scale = max(float(s[0].abs().max()) for s in singles)
    for nb in (1, 2, 3, 4, 7):
        m1, m2 = hp.alm2map_spin_batch(ad[:nb], nside, spin, lmax)
        assert m1.shape == m2.shape == (nb, 12 * nside**2)
        for b in range(nb):
            assert float((m1[b] - singles[b][0]).abs().max()) <= 1e-12 * scale, (nb, b)
            assert float((m2[b] - singles[b][1]).abs().max()) <= 1e-12 * scale, (nb, b)
    if nside <= 32:
        r1, r2 = H.alm2map_spin(alms[3], np.zeros(nalm, dtype=complex), nside, spin, lmax)
        assert np.abs(m1[3].cpu().numpy() - r1).max() <= 1e-10 * np.abs(r1).max() and np.abs(m2[3].cpu().numpy() - r2).max() <= 1e-10 * np.abs(r1).max()
