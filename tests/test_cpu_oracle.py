"""CPU tests (-m "not gpu"): the oracle against the golden vectors produced by executing the
reference's own source (tests/golden/make_golden.py), the reference's dependency-free
known-answer tests, and the mathematical ground truth for the SHT restatements."""
import os

import numpy as np
import pytest

from helpers import MockCosmology, synthetic_gls
from oracle import glass_ref as G
from oracle import healpix_ref as H

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))


# ------------------------------------------------------------------ golden vectors
def test_golden_iternorm_cls2cov():
    cov = np.array([[1.0, 0.2, 0.1], [0.2, 0.5, 0.2], [0.1, 0.2, 0.3]])
    for k in (0, 1, 2):
        rows = [np.pad(cov[i, i::-1][: min(i, k) + 1], (0, k + 1 - min(i + 1, k + 1))) for i in range(3)]
        assert np.array_equal(np.stack(G.iternorm(rows)), GOLD[f"iternorm_k{k}"])
    for name, (nshell, lmax, ncorr, ragged) in {"a": (4, 12, 2, False), "b": (5, 9, None, False), "c": (4, 10, 1, True)}.items():
        nc = nshell - 1 if ncorr is None else ncorr
        gls = synthetic_gls(nshell, lmax, nc, ragged)
        rows = G.cls2cov_rows(gls, lmax + 1, nshell, nc)
        assert np.array_equal(np.stack(rows), GOLD[f"cls2cov_{name}"])
        assert np.array_equal(np.stack(G.iternorm(rows)), GOLD[f"iternorm_{name}"])


def test_golden_generate_alm():
    """alm of _generate_grf with the seed-42 NumPy normals the reference draws."""
    for name, (nshell, lmax, ncorr, ragged) in {"a": (4, 12, 2, False), "b": (5, 9, None, False), "c": (4, 10, 1, True)}.items():
        nc = nshell - 1 if ncorr is None else ncorr
        gls = synthetic_gls(nshell, lmax, nc, ragged)
        rng = np.random.default_rng(42)
        n = (lmax + 1) * (lmax + 2) // 2
        zs = [rng.standard_normal((n, 2)) @ np.array([1, 1j]) for _ in range(nshell)]
        alms = G.generate_alms(gls, ncorr, zs)
        assert np.array_equal(np.stack(alms), GOLD[f"grf_alm_{name}"])


def test_golden_transformations_and_reorder():
    x = GOLD["lognormal_x"]
    assert np.array_equal(G.lognormal(x, 0.35, 0.7), GOLD["lognormal_y"])
    assert np.array_equal(G.lognormal(x, 0.35, 1.0), GOLD["lognormal_y1"])
    assert np.array_equal(G.squared_normal(x, 0.3, 1.5), GOLD["sqnormal_y"])
    assert np.array_equal(G.glass_to_healpix_alm(GOLD["g2h_in"]), GOLD["g2h_out"])
    assert np.array_equal(G.multalm(GOLD["g2h_in"][:6], np.array([2.0, 0.5, 1.0])), GOLD["multalm_out"])


def test_golden_points():
    delta, vis = GOLD["pt_delta"], GOLD["pt_vis"]
    assert float(GOLD["ARCMIN2_SPHERE"]) == G.ARCMIN2_SPHERE
    for tag, (bias, v, model, rm) in {
        "none": (None, None, "linear", False),
        "lin_vis": (0.8, vis, "linear", False),
        "loglin": (1.3, None, "loglinear", False),
        "lin_vis_rm": (0.8, vis, "linear", True),
    }.items():
        assert np.array_equal(G.expected_count(delta, 1e-3, bias, v, model, rm), GOLD[f"pt_nbar_{tag}"]), tag
    counts = GOLD["pt_counts"]
    for batch in (1_000_000, 500, 37, 1):
        cuts = G.batch_cuts(counts, batch)
        assert [c[2] for c in cuts] == list(GOLD[f"pt_batches_{batch}"])
        ipix = np.concatenate([np.repeat(np.arange(a, b), counts[a:b]) for a, b, _ in cuts])
        assert np.array_equal(ipix, GOLD[f"pt_ipix_{batch}"])


def test_golden_lensing_galaxies_shapes():
    cosmo = MockCosmology()
    deltas = GOLD["mpc_deltas"]
    mpc = G.MultiPlaneConvergence(cosmo)
    for i in range(5):
        mpc.add_window(deltas[i].copy(), np.array([i, i + 1.0, i + 2.0]), np.array([0.0, 1.0, 0.0]), i + 1.0)
        assert np.array_equal(mpc.kappa, GOLD["mpc_kappas"][i])
    for red, key in ((True, "gs_reduced"), (False, "gs_plain")):
        got = G.galaxy_shear(GOLD["gs_lon"], GOLD["gs_lat"], GOLD["gs_eps"], GOLD["gs_kappa"], GOLD["gs_g1"], GOLD["gs_g2"], red)
        np.testing.assert_allclose(got, GOLD[key], rtol=1e-15, atol=0)
    np.testing.assert_allclose(G.ellipticity_intnorm_from_normals(0.256, GOLD["eps_normals"]), GOLD["eps_intnorm"], rtol=1e-15)
    np.testing.assert_allclose(G.redshifts_from_nz_uniform(GOLD["z_grid"], GOLD["z_nz"], GOLD["z_uniform"]), GOLD["z_samples"], rtol=1e-15)


# ------------------------------------------------------------------ reference known-answer tests, ported
@pytest.mark.parametrize("k", [0, 1, 2])
@pytest.mark.parametrize("nd", [False, True])
def test_iternorm_against_explicit(k, nd):
    """tests/core/test_fields.py:31-104."""
    cov = np.array([[1.0, 0.2, 0.1], [0.2, 0.5, 0.2], [0.1, 0.2, 0.3]])
    if nd:
        cov = np.stack([cov, np.array([[1.4, 0.4, 0.5], [0.4, 1.5, 0.8], [0.5, 0.8, 1.3]])])
    rows = [cov[..., i, i::-1][..., : min(i, k) + 1] for i in range(3)]
    rows = [np.concatenate([r, np.zeros(r.shape[:-1] + (k + 1 - r.shape[-1],))], axis=-1) for r in rows]
    for n, w in enumerate(G.iternorm(rows)):
        Sn = cov[..., :n, :n].copy()
        cn = cov[..., :n, n].copy()
        vn = cov[..., n, n]
        if n > k:
            Sn *= np.abs(np.arange(n)[:, None] - np.arange(n)) <= k
            cn *= np.arange(n) >= n - k
        Sninv = np.linalg.pinv(Sn)
        An = np.swapaxes(np.linalg.cholesky(Sninv + 1e-100 * np.eye(n)), -1, -2) if n else np.zeros(Sn.shape)
        an = (An @ cn[..., None])[..., 0]
        sn = np.sqrt(vn - np.vecdot(an, an))
        a, s = w[..., :-1], w[..., -1]
        np.testing.assert_allclose(np.vecdot(a, a), np.vecdot(an, an), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(s, sn, rtol=1e-12)


def test_iternorm_errors():
    """tests/core/test_fields.py:107-119."""
    with pytest.raises(ValueError, match="empty covariance"):
        G.iternorm([np.ones(0)])
    with pytest.raises(ValueError, match="shape mismatch"):
        G.iternorm([np.ones(1), np.ones((5, 2))])
    with pytest.raises(ValueError, match="not positive definite"):
        G.iternorm([np.array([1.0]), np.array([0.1, 1.0])])


def test_multi_plane_matrix_relation():
    """tests/core/test_lensing.py:64-86 on the oracle: kappa_i = sum_j M_ij delta_j."""
    mat = GOLD["mpc_matrix"]
    assert np.array_equal(mat, np.tril(mat))
    np.testing.assert_allclose(mat @ GOLD["mpc_deltas"], GOLD["mpc_kappas"], rtol=1e-12)


# ------------------------------------------------------------------ SHT restatement vs ground truth
def _alm(lmax, seed):
    rng = np.random.default_rng(seed)
    n = H.alm_size(lmax)
    a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    a[: lmax + 1] = a[: lmax + 1].real
    return a


@pytest.mark.parametrize("nside,lmax", [(1, 3), (2, 6), (4, 11), (3, 8)])
def test_alm2map_vs_direct_sum(nside, lmax):
    alm = _alm(lmax, nside)
    d = H.alm2map_direct(alm, nside, lmax)
    assert np.abs(H.alm2map(alm, nside, lmax) - d).max() < 1e-12 * np.abs(d).max()


def test_analytic_maps():
    nside, lmax = 8, 4
    th, _ = H.pix2ang_centers(nside)
    alm = np.zeros(H.alm_size(lmax), dtype=complex)
    alm[H.alm_index(lmax, 0, 0)] = 2.0
    np.testing.assert_allclose(H.alm2map(alm, nside, lmax), 2.0 / np.sqrt(4 * np.pi), rtol=1e-14)
    alm[:] = 0
    alm[H.alm_index(lmax, 1, 0)] = 1.5
    np.testing.assert_allclose(H.alm2map(alm, nside, lmax), np.sqrt(3 / (4 * np.pi)) * 1.5 * np.cos(th), atol=1e-14)


@pytest.mark.parametrize("spin", [1, 2])
def test_alm2map_spin_vs_direct_sum(spin):
    nside, lmax = 4, 9
    e, b = _alm(lmax, 20 + spin), _alm(lmax, 30 + spin)
    for l in range(spin):
        for m in range(l + 1):
            e[H.alm_index(lmax, l, m)] = b[H.alm_index(lmax, l, m)] = 0
    d1, d2 = H.alm2map_spin_direct(e, b, nside, spin, lmax)
    r1, r2 = H.alm2map_spin(e, b, nside, spin, lmax)
    assert max(np.abs(r1 - d1).max(), np.abs(r2 - d2).max()) < 1e-12 * np.abs(d1).max()


def test_spin0_goldberg_equals_scipy():
    from scipy.special import sph_harm_y

    th, ph = np.linspace(0.1, 3.0, 9), np.linspace(0.0, 6.0, 9)
    for l, m in [(3, 2), (5, -3), (4, 0), (7, 7)]:
        assert np.abs(H.sYlm_goldberg(0, l, m, th, ph) - sph_harm_y(l, m, th, ph)).max() < 1e-14


def test_map2alm_roundtrip():
    nside, lmax = 8, 10
    alm = _alm(lmax, 5)
    mp = H.alm2map(alm, nside, lmax)
    e0 = np.abs(H.map2alm(mp, lmax, niter=0) - alm).max()
    e3 = np.abs(H.map2alm(mp, lmax, niter=3) - alm).max()
    assert e3 < 1e-3 * e0 and e3 < 1e-4


def test_pixel_functions():
    rng = np.random.default_rng(0)
    for nside in (1, 2, 4, 16, 3, 48):
        npix = 12 * nside**2
        p = np.arange(npix)
        th, ph = H.pix2ang_centers(nside)
        t2, p2 = H.ring2ang_uv(nside, p, 0.5, 0.5)
        assert np.allclose(th, t2, atol=1e-14) and np.allclose(np.mod(ph, 2 * np.pi), p2, atol=1e-13)
        u, v = rng.random(npix), rng.random(npix)
        assert np.array_equal(H.ang2pix(nside, *H.ring2ang_uv(nside, p, u, v)), p)
        lon, lat = H.ring2ang_uv(nside, p, u, v, lonlat=True)
        assert lon.min() >= 0 and lon.max() < 360 and lat.min() >= -90 and lat.max() <= 90
    with pytest.raises(ValueError):
        H.npix2nside(13)


def test_c_oracle_matches_numpy_oracle():
    from oracle import sht_c

    for nside, lmax in [(4, 11), (3, 7), (16, 47), (48, 100)]:
        alm = _alm(lmax, nside)
        ref = H.alm2map(alm, nside, lmax)
        for ld in (False, True):
            got = sht_c.alm2map(alm, nside, lmax, long_double=ld)
            assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max()
    # the mlim cut-off (shared with the CUDA kernels) does not change the result
    alm = _alm(300, 1)
    a = sht_c.alm2map(alm, 128, 300)
    b = sht_c.alm2map(alm, 128, 300, use_mlim=True)
    assert np.abs(a - b).max() < 1e-13 * np.abs(a).max()


def test_timed_cpu_arm_matches_checker():
    """oracle/sht_fast.cpp (the SIMD/OpenMP synthesis that bench.py times as the CPU arm) against
    the scalar C oracle and the 80-bit one, incl. high-m rings that start scaled down, a band
    limit above 2 nside and the smallest supported nside."""
    from oracle import sht_c

    for nside, lmax in [(2, 5), (4, 11), (16, 47), (64, 191), (128, 255), (256, 700)]:
        alm = _alm(lmax, nside)
        ref = sht_c.alm2map(alm, nside, lmax, use_mlim=True)
        tm = {}
        got = sht_c.alm2map_fast(alm, nside, lmax, nthreads=2, timings=tm)
        assert np.abs(got - ref).max() < 2e-12 * np.abs(ref).max(), (nside, lmax)
        assert tm["legendre_s"] >= 0 and tm["fft_s"] >= 0
    truth = sht_c.alm2map(alm, 256, 700, use_mlim=True, long_double=True)
    assert np.abs(got - truth).max() < 1e-11 * np.abs(truth).max()
    with pytest.raises(ValueError):
        sht_c.alm2map_fast(_alm(5, 0), 3, 5)


def test_golden_uniform_positions():
    """uniform_positions (glass/points.py:543-607) executed from the reference source; the
    oracle replays it bit-exactly from the same Poisson totals and uniform deviates."""
    got = list(G.uniform_positions_from_uniforms(GOLD["up_totals"], GOLD["up_u_lon"], GOLD["up_u_lat"]))
    assert np.array_equal(np.concatenate([g[0] for g in got]), GOLD["up_lon"])
    assert np.array_equal(np.concatenate([g[1] for g in got]), GOLD["up_lat"])
    assert np.array_equal(np.stack([g[2] for g in got]), GOLD["up_count"])


def test_golden_gaussian_phz():
    """gaussian_phz (glass/galaxies.py:350-455) executed from the reference source with bounds
    that need several rejection rounds; the oracle replays it from the same normal deviates."""
    got = G.gaussian_phz_from_normals(GOLD["phz_z"], 0.2, list(GOLD["phz_normals"]), 0.1, 1.2)
    assert np.array_equal(got, GOLD["phz_out"])
    assert got.min() >= 0.1 and got.max() <= 1.2


@pytest.mark.parametrize("nside,lmax,spin", [(8, 20, 2), (16, 40, 1), (8, 16, 3)])
def test_c_oracle_spin_and_analysis_match_numpy_oracle(nside, lmax, spin):
    """oracle/sht_ref.c's spin-weighted synthesis and analysis pass (used by the GPU parity tests at
    nside 512 / 1024, where NumPy is too slow) restate oracle/healpix_ref.py, which is validated
    against Goldberg's closed form and direct sums: the two agree to rounding, double and 80-bit."""
    from oracle import healpix_ref as H
    from oracle import sht_c

    rng = np.random.default_rng(nside * spin)
    nalm = (lmax + 1) * (lmax + 2) // 2
    e = rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm)
    b = rng.standard_normal(nalm) + 1j * rng.standard_normal(nalm)
    e[: lmax + 1], b[: lmax + 1] = e[: lmax + 1].real, b[: lmax + 1].real
    for blm in (b, None):
        r1, r2 = H.alm2map_spin(e, np.zeros_like(e) if blm is None else blm, nside, spin, lmax)
        for ld in (False, True):
            m1, m2 = sht_c.alm2map_spin(e, blm, nside, spin, lmax, long_double=ld)
            assert np.abs(m1 - r1).max() <= 1e-12 * np.abs(r1).max() and np.abs(m2 - r2).max() <= 1e-12 * np.abs(r1).max()
    mp = rng.standard_normal(12 * nside * nside)
    w = 1 + 0.01 * rng.standard_normal(4 * nside - 1)
    for niter, rw in ((0, w), (2, None)):
        a = sht_c.map2alm(mp, lmax, niter=niter, ring_w=rw)
        r = H.map2alm(mp, lmax, niter=niter, ring_w=rw)
        assert np.abs(a - r).max() <= 1e-12 * np.abs(r).max()


def test_visibility_mask_oracle_properties():
    """oracle.healpix_ref.query_strip / get_interpol / rotate_map_pixel (restated from the published
    HEALPix C++ algorithms; healpy itself is absent): the reference's own known answers
    (tests/core/test_observations.py:20-52) and the defining properties of the interpolation."""
    from oracle import healpix_ref as H

    for nside in (1, 2, 4, 16):
        v = H.vmap_galactic_ecliptic(nside)
        assert v.shape == (12 * nside**2,) and v.min() >= 0 and v.max() <= 1 + 1e-15
        z = H.vmap_galactic_ecliptic(nside, galactic=(0, 0), ecliptic=(0, 0))
        assert np.array_equal(z, np.zeros_like(z))  # "no rotation" case of the reference's test
    nside = 8
    theta, phi = H.pix2ang_centers(nside)
    # strip = pixels whose centre colatitude lies between the bounds; reversed bounds = complement
    s = H.query_strip(nside, 0.7, 2.1)
    assert np.array_equal(s == 1, (theta > 0.7) & (theta < 2.1))
    c = H.query_strip(nside, 2.1, 0.7)
    assert np.array_equal(c == 1, (theta < 0.7) | (theta > 2.1))
    assert np.all(H.query_strip(nside, 0.0, np.pi) == 1)
    assert np.all(H.query_strip(nside, 0.0, 0.0) == 1)  # equal bounds: the complement branch, everything
    # interpolation: weights are a partition of unity, pixels valid, exact at pixel centres
    rng = np.random.default_rng(5)
    th, ph = rng.uniform(0, np.pi, 5000), rng.uniform(-7, 7, 5000)
    p, w = H.get_interpol(nside, th, ph)
    assert np.abs(w.sum(0) - 1).max() < 1e-15 and w.min() > -1e-15 and p.min() >= 0 and p.max() < 12 * nside**2
    m = rng.random(12 * nside**2)
    np.testing.assert_allclose(H.get_interp_val(m, theta, phi), m, rtol=0, atol=1e-13)
    np.testing.assert_allclose(H.rotate_map_pixel(m, "CC"), m, rtol=0, atol=1e-13)
    # second-order accurate for a smooth function away from the poles
    f = lambda t, q: np.cos(t) + 0.3 * np.sin(t) * np.cos(q)  # noqa: E731
    errs = []
    away = (th > 0.3) & (th < np.pi - 0.3)
    for n in (16, 32, 64):
        t, q = H.pix2ang_centers(n)
        errs.append(np.abs(H.get_interp_val(f(t, q), th, ph) - f(th, ph))[away].max())
    assert errs[1] < errs[0] / 3 and errs[2] < errs[1] / 3
    # coordinate systems: rotations, the poles where the almanac puts them
    for cpair in ("GC", "CE", "EG"):
        M = H.coordconv_matrix(cpair)
        assert np.abs(M @ M.T - np.eye(3)).max() < 2e-9
    x, y, z = H.coordconv_matrix("GC") @ np.array([0.0, 0.0, 1.0])
    assert abs(np.degrees(np.arctan2(y, x)) % 360 - 192.86) < 0.01 and abs(np.degrees(np.arcsin(z)) - 27.13) < 0.01
    x, y, z = H.coordconv_matrix("EC") @ np.array([0.0, 0.0, 1.0])
    assert abs(np.degrees(np.arctan2(y, x)) % 360 - 270) < 1e-9 and abs(np.degrees(np.arcsin(z)) - 66.56) < 0.01


def test_int8_digit_scheme_emulation():
    """The arithmetic of the INT8 tensor-core Legendre kernel (csrc/sht_ozaki.cu), emulated step by step with exact
    integers (tests/studies/ozaki_device_scheme.py): magic-constant fixed point, balanced base-256 digits as bytes ^ 0x80,
    the 21 digit products with i + j >= 5 grouped by significance, the pair-wise int32 combination -- reproduces the FP64
    contraction F = sum_k lambda_k A_k to 1e-12 of its maximum, with every integer inside its range."""
    import importlib.util
    import os

    from oracle import healpix_ref as H

    path = os.path.join(os.path.dirname(__file__), "studies", "ozaki_device_scheme.py")
    spec = importlib.util.spec_from_file_location("ozaki_device_scheme", path)
    oz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(oz)
    # the bytes of V = rint(x s) + BIAS, sign-flipped, ARE the balanced digits
    x = np.array([0.0, 1.0, -1.0, 0.123456789, -0.987654321, 1.999, -1.999])
    s_, inv = oz.scales(np.full(x.shape, 1.999))
    d = oz.digits(x, s_)
    assert d.min() >= -128 and d.max() <= 127
    rec = sum(d[j].astype(np.float64) * 256.0**j for j in range(oz.ND)) * inv
    assert np.abs(rec - x).max() <= 2.0**-46
    assert np.all(oz.digits(np.zeros(3), s_[:3]) == 0)
    nside, lmax = 64, 127
    ri = H.ring_info(nside)
    z, sth = ri["z"][: 2 * nside], ri["sth"][: 2 * nside]
    rng = np.random.default_rng(2)
    for m in (0, 3, 40, 101):
        lam = H.lam_lm(lmax, m, z, sth).T[:, ::2]
        l = np.arange(m, lmax + 1, 2)
        a = rng.standard_normal((l.size, 16)) * np.sqrt(1e-2 * (l + 1.0) ** -1.5)[:, None]
        ref = lam @ a
        F, worst = oz.contraction(lam, a, 64)
        assert np.abs(F - ref).max() <= 1e-12 * np.abs(ref).max(), m
        assert worst < 6 * 64 * 128 * 128
