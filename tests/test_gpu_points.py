"""GPU parity: galaxy counts / positions / pixel functions against the oracle."""
import numpy as np
import pytest
import torch

from oracle import glass_ref as G
from oracle import healpix_ref as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nside", [1, 2, 4, 16, 3, 48, 128])
def test_ring2ang_uv_and_ang2pix(cuda_device, nside):
    from glass_b200 import healpix as hp

    rng = np.random.default_rng(nside)
    npix = 12 * nside * nside
    ipix = np.arange(npix) if npix <= 4096 else rng.integers(0, npix, 4096)
    u, v = rng.random(ipix.size), rng.random(ipix.size)
    for lonlat in (False, True):
        a, b = hp.ring2ang_uv(nside, ipix, u, v, lonlat=lonlat)
        ra, rb = H.ring2ang_uv(nside, ipix, u, v, lonlat=lonlat)
        np.testing.assert_allclose(a, ra, rtol=0, atol=1e-12 * (57.3 if lonlat else 1.0))
        np.testing.assert_allclose(b, rb, rtol=0, atol=1e-12 * (57.3 if lonlat else 1.0))
        # pixel indexing bit-exact, and identical to the oracle's on the same angles
        p = hp.ang2pix(nside, ra, rb, lonlat=lonlat)
        assert p.dtype == np.int64
        assert np.array_equal(p, H.ang2pix(nside, ra, rb, lonlat=lonlat))
    # pixel centres map back to their pixel
    th, ph = hp.ring2ang_uv(nside, ipix, 0.5, 0.5)
    assert np.array_equal(hp.ang2pix(nside, th, ph), ipix)
    lon, lat = hp.randang(nside, ipix, lonlat=True)
    assert lon.min() >= 0 and lon.max() < 360 and lat.min() >= -90 and lat.max() <= 90
    assert np.array_equal(hp.ang2pix(nside, lon, lat, lonlat=True), ipix)
    lon2, _ = hp.randang(nside, ipix, lonlat=True)  # same seed-42 stream on every call
    assert np.array_equal(lon, lon2)


def _nbar(delta, ngal, bias, vis, model, remove_monopole, cuda_device):
    from glass_b200.points import _Population, linear_bias, loglinear_bias

    bm = {"linear": linear_bias, "loglinear": loglinear_bias}[model]
    pop = _Population(delta, vis, ngal, bias, bm, remove_monopole, 42, 0, np.zeros(delta.size, dtype=np.int64), cuda_device, want_nbar=True)
    return pop.nbar.cpu().numpy()


@pytest.mark.parametrize("bias,vis,model,rm", [(None, False, "linear", False), (0.8, True, "linear", False), (1.3, False, "loglinear", False), (0.8, True, "linear", True)])
def test_expected_count_vs_oracle(cuda_device, bias, vis, model, rm):
    nside = 16
    rng = np.random.default_rng(3)
    delta = np.expm1(0.5 * rng.standard_normal(12 * nside**2) - 0.125)
    v = rng.random(delta.size) if vis else None
    got = _nbar(delta, 1e-3, bias, v, model, rm, cuda_device)
    ref = G.expected_count(delta, 1e-3, bias, v, model, rm)
    if model == "linear" and not rm:
        assert np.array_equal(got, ref)  # same roundings as NumPy: bit-identical
    else:
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())


@pytest.mark.parametrize("batch", [1_000_000, 500, 37, 1])
def test_positions_supplied_deviates(cuda_device, batch):
    """Counts, pixel indices and batch cuts bit-exact given supplied Poisson/uniform deviates."""
    import glass_b200
    from glass_b200.rng import Deviates

    nside = 8
    npix = 12 * nside**2
    rng = np.random.default_rng(5)
    delta = rng.random(npix) - 0.5
    counts = rng.poisson(3.0 * (1 + delta))
    counts[10:40] = 0
    counts[100] = 60  # a pixel larger than the small batches
    uv = lambda n: (np.random.default_rng(42).random(n), np.random.default_rng(43).random(n))  # noqa: E731
    ref = G.positions_from_counts(counts, nside, batch, uv)
    got = list(glass_b200.positions_from_delta(1e-3, delta, rng=Deviates(poisson=[counts], uv=uv), batch=batch))
    assert len(got) == len(ref)
    for (lon, lat, n), (rlon, rlat, rn) in zip(got, ref):
        assert n == rn and lon.shape == (rn,)
        np.testing.assert_allclose(lon, rlon, rtol=0, atol=1e-10)
        np.testing.assert_allclose(lat, rlat, rtol=0, atol=1e-10)
        assert np.array_equal(H.ang2pix(nside, lon, lat, lonlat=True), H.ang2pix(nside, rlon, rlat, lonlat=True))
    assert sum(g[2] for g in got) == counts.sum()


def test_positions_from_delta_structure(cuda_device):
    """Port of the reference's structural tests (tests/core/test_points.py:235-384)."""
    import glass_b200

    nside = 32
    npix = 12 * nside**2
    # zero delta, no bias: expected total = ngal * ARCMIN2_SPHERE
    ngal = 1e-3
    tot = 0
    for lon, lat, cnt in glass_b200.positions_from_delta(ngal, np.zeros(npix), rng=7):
        assert isinstance(cnt, int) and lon.shape == lat.shape == (cnt,)
        assert lon.min() >= 0 and lon.max() < 360 and lat.min() >= -90 and lat.max() <= 90
        tot += cnt
    mean = ngal * glass_b200.points.ARCMIN2_SPHERE
    assert abs(tot - mean) < 6 * np.sqrt(mean)
    # populations: ngal (2,), delta (3,1,npix) -> dims (3,2); count is one-hot * n
    delta = np.zeros((3, 1, npix))
    ngal2 = np.array([1e-3, 2e-3])
    vis = np.ones(npix)
    vis[: npix // 2] = 0.0
    seen = np.zeros((3, 2), dtype=np.int64)
    for lon, lat, cnt in glass_b200.positions_from_delta(ngal2, delta, 0.8, vis, rng=8):
        assert cnt.shape == (3, 2) and np.count_nonzero(cnt) == 1
        assert lat.max() < 2.0  # southern half only (vis zero in the north; boundary pixels straddle the equator)
        seen += cnt
    assert np.all(seen > 0) and abs(seen[:, 1].sum() / seen[:, 0].sum() - 2) < 0.1
    # all-zero visibility: nothing is yielded
    assert list(glass_b200.positions_from_delta(ngal, np.zeros(npix), None, np.zeros(npix))) == []
    with pytest.raises(TypeError, match="bias_model must be callable"):
        next(glass_b200.positions_from_delta(ngal, np.zeros(npix), bias_model=0))
    # device in -> device out, custom callable bias model
    d = torch.zeros(npix, dtype=torch.float64, device=cuda_device)
    out = list(glass_b200.positions_from_delta(ngal, d, 1.0, bias_model=lambda dd, b: b * dd * 0.5, rng=9))
    assert out and out[0][0].is_cuda


def test_poisson_and_position_statistics(cuda_device):
    """Random draws validated statistically (north_star): KS/moment tests on counts and
    uniformity of in-pixel positions."""
    from scipy import stats

    from glass_b200.points import _Population, linear_bias

    nside = 64
    npix = 12 * nside**2
    for lam in (0.08, 3.0, 9.5, 25.0, 400.0):
        ngal = lam / (G.ARCMIN2_SPHERE / npix)
        pop = _Population(np.zeros(npix), None, ngal, None, linear_bias, False, 123, 0, None, cuda_device, mode="scan")
        c = pop.counts.cpu().numpy()
        if lam < 10:  # the galaxy list of the same draw (what "auto" picks for sparse maps): np.repeat(arange, counts)
            lst = _Population(np.zeros(npix), None, ngal, None, linear_bias, False, 123, 0, None, cuda_device, mode="list")
            assert lst.total == c.sum() and lst.counts is None and lst.off is None
            assert np.array_equal(lst.gpix[: lst.total].cpu().numpy(), np.repeat(np.arange(npix), c))
            for batch in (1000, 37):
                assert list(lst.cuts(batch)) == list(pop.cuts(batch))
            a, b, n = list(pop.cuts(1000))[3]
            for x, y in zip(lst.fill(a, b, n, want_ipix=True), pop.fill(a, b, n, want_ipix=True)):
                assert torch.equal(x, y)
        assert abs(c.mean() - lam) < 6 * np.sqrt(lam / npix)
        assert abs(c.var() - lam) < 6 * lam * np.sqrt(2.0 / npix) + 6 * np.sqrt(lam / npix)
        # chi-square of the histogram against the Poisson pmf
        kmax = int(lam + 6 * np.sqrt(lam) + 6)
        obs = np.bincount(np.minimum(c, kmax), minlength=kmax + 1).astype(float)
        pmf = stats.poisson.pmf(np.arange(kmax + 1), lam)
        pmf[-1] += stats.poisson.sf(kmax, lam)
        keep = pmf * npix > 5
        chi2 = ((obs[keep] - pmf[keep] * npix) ** 2 / (pmf[keep] * npix)).sum()
        assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-5, (lam, chi2)
        assert pop.off[-1].item() == c.sum() and np.array_equal(pop.off.cpu().numpy()[:-1], np.cumsum(c) - c)
    # in-pixel (u, v) uniformity: invert positions of a single big pixel population
    pop = _Population(np.zeros(npix), None, 200.0 / (G.ARCMIN2_SPHERE / npix), None, linear_bias, False, 5, 0, None, cuda_device)
    assert pop.mode == "scan"
    lon, lat, ipix = pop.fill(0, npix, pop.total, None, want_ipix=True)
    ip = ipix.cpu().numpy()
    assert np.array_equal(H.ang2pix(nside, lon.cpu().numpy(), lat.cpu().numpy(), lonlat=True), ip)
    assert np.all(np.diff(ip) >= 0)  # sorted by ring pixel index (points.py:426)
    # z = sin(lat) is uniform over the sphere for uniform-in-pixel sampling on an equal-area grid
    z = np.sin(np.radians(lat.cpu().numpy()))
    assert stats.kstest(z, "uniform", args=(-1, 2)).pvalue > 1e-4
    assert stats.kstest(lon.cpu().numpy(), "uniform", args=(0, 360)).pvalue > 1e-4


def test_batch_cut_corner_cases(cuda_device):
    """Exact fit ending at a 1000-pixel group boundary, empty batches before an oversize
    pixel, trailing zeros: the cut sequence equals the reference loop's (points.py:409-437)."""
    import glass_b200
    from glass_b200.rng import Deviates

    nside = 16
    npix = 12 * nside**2
    uv = lambda n: (np.full(n, 0.5), np.full(n, 0.5))  # noqa: E731
    cases = []
    c = np.zeros(npix, dtype=np.int64)
    c[0], c[1500], c[1501], c[2999] = 5, 9, 1, 2
    cases.append((c, 5))
    c = np.zeros(npix, dtype=np.int64)
    c[999], c[1000], c[2500] = 3, 2, 4
    cases.append((c, 3))
    c = np.random.default_rng(0).poisson(0.01, npix)
    cases.append((c, 2))
    c = np.random.default_rng(1).poisson(2.0, npix)
    cases.append((c, 1000))
    for counts, batch in cases:
        ref = G.batch_cuts(counts, batch)
        got = list(glass_b200.positions_from_delta(1e-3, np.zeros(npix), rng=Deviates(poisson=[counts], uv=uv), batch=batch))
        assert [g[2] for g in got] == [r[2] for r in ref], batch
        assert sum(g[2] for g in got) == counts.sum()


def test_uniform_positions(cuda_device):
    """glass/points.py:543-607: supplied deviates against the golden vectors of the reference
    (lon bit-exact, lat to 1 ulp of asin), count semantics, and the Philox draws statistically."""
    import os

    from scipy import stats

    import glass_b200
    from glass_b200.rng import Deviates

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    pos = {"i": 0}

    def uniforms(n):
        a = pos["i"]
        pos["i"] += n
        return gold["up_u_lon"][a : a + n], gold["up_u_lat"][a : a + n]

    got = list(glass_b200.uniform_positions(gold["up_ngal"], rng=Deviates(poisson=[gold["up_totals"]], uniform=uniforms)))
    assert len(got) == 2 and all(isinstance(g[0], np.ndarray) for g in got)
    assert np.array_equal(np.concatenate([g[0] for g in got]), gold["up_lon"])
    np.testing.assert_allclose(np.concatenate([g[1] for g in got]), gold["up_lat"], rtol=0, atol=2e-14)
    assert np.array_equal(np.stack([g[2] for g in got]), gold["up_count"])
    # scalar density: count is a Python int, arrays follow the input's device
    (lon, lat, cnt), = list(glass_b200.uniform_positions(torch.tensor(3e-3, device=cuda_device), rng=5))
    assert isinstance(cnt, int) and lon.is_cuda and lon.numel() == cnt
    lam = G.ARCMIN2_SPHERE * 3e-3
    assert abs(cnt - lam) < 6 * np.sqrt(lam)
    lo, la = lon.cpu().numpy(), lat.cpu().numpy()
    assert lo.min() >= -180 and lo.max() < 180 and la.min() >= -90 and la.max() <= 90
    assert stats.kstest(lo, "uniform", args=(-180, 360)).pvalue > 1e-4
    assert stats.kstest(np.sin(np.radians(la)), "uniform", args=(-1, 2)).pvalue > 1e-4
    assert abs(np.corrcoef(lo, la)[0, 1]) < 6 / np.sqrt(cnt)
    # zero density: an empty batch is still yielded, like the reference
    (lon, lat, cnt), = list(glass_b200.uniform_positions(0.0, rng=1))
    assert cnt == 0 and lon.size == 0


def test_poisson_with_visibility_and_bias_variants(cuda_device):
    """The sampling instantiations of the fused count/offset kernel (visibility, log-linear bias,
    monopole removal): zero where the visibility is zero, Poisson moments elsewhere, and offsets
    that are the exclusive scan of the counts."""
    from glass_b200.points import _Population, linear_bias, loglinear_bias

    nside = 64
    npix = 12 * nside**2
    rng = np.random.default_rng(2)
    delta = np.expm1(0.4 * rng.standard_normal(npix) - 0.08)
    vis = (np.arange(npix) % 3 != 0).astype(float) * 0.5
    for model, bias, rm in ((linear_bias, 0.9, False), (loglinear_bias, 1.4, False), (linear_bias, 0.9, True)):
        ngal = 6.0 / (G.ARCMIN2_SPHERE / npix)
        pop = _Population(delta, vis, ngal, bias, model, rm, 77, 3, None, cuda_device, want_nbar=True, mode="scan")
        c = pop.counts.cpu().numpy()
        lst = _Population(delta, vis, ngal, bias, model, rm, 77, 3, None, cuda_device, want_nbar=True, mode="list")
        assert torch.equal(lst.nbar, pop.nbar) and np.array_equal(lst.gpix[: lst.total].cpu().numpy(), np.repeat(np.arange(npix), c))
        nbar = np.clip(pop.nbar.cpu().numpy(), 0, None)
        name = "linear" if model is linear_bias else "loglinear"
        ref = G.expected_count(delta, ngal, bias, vis, name, rm)
        np.testing.assert_allclose(pop.nbar.cpu().numpy(), ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
        assert np.all(c[vis == 0] == 0)
        assert abs(c.sum() - nbar.sum()) < 6 * np.sqrt(nbar.sum())
        # per-pixel standardised residuals have unit variance for Poisson draws
        m = nbar > 0.5
        zres = (c[m] - nbar[m]) / np.sqrt(nbar[m])
        assert abs(zres.var() - 1.0) < 0.05 and abs(zres.mean()) < 0.02
        off = pop.off.cpu().numpy()
        assert off[-1] == c.sum() and np.array_equal(off[:-1], np.cumsum(c) - c)


def _unit(lon, lat):
    lo, la = np.radians(lon), np.radians(lat)
    return np.stack([np.cos(la) * np.cos(lo), np.cos(la) * np.sin(lo), np.sin(la)])


def test_displace_deflect_displacement_golden(cuda_device):
    """glass.displace / glass.deflect / glass.displacement against vectors produced by executing
    the reference's own source (tests/golden/make_golden.py --displace).  Transcendental
    functions differ from NumPy's libm in the last bits: 1e-11 degrees, and the displaced POINT
    (unit vector) to 1e-13 where the longitude itself is ill-conditioned (poles)."""
    import os

    import glass_b200

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "glass_reference_displace.npz"))
    lon, lat, alpha = g["lon"], g["lat"], g["alpha"]
    ok = np.abs(lat) < 89.9
    for fn, key in ((glass_b200.displace, "displace"), (glass_b200.deflect, "deflect")):
        a, b = fn(lon, lat, alpha)
        assert isinstance(a, np.ndarray) and a.shape == lon.shape
        ra, rb = g[key + "_lon"], g[key + "_lat"]
        assert np.abs(_unit(a, b) - _unit(ra, rb)).max() < 1e-13
        okk = ok & (np.abs(rb) < 89.9)
        assert np.abs(a - ra)[okk].max() < 1e-11 and np.abs(b - rb)[okk].max() < 1e-11
        a2, b2 = fn(lon, lat, np.stack([alpha.real, alpha.imag]))  # leading axis of size 2
        assert np.array_equal(a, a2) and np.array_equal(b, b2)
    d = glass_b200.displacement(lon, lat, g["to_lon"], g["to_lat"])
    assert d.dtype == np.complex128
    assert np.abs(d - g["displacement"]).max() < 1e-12
    # CUDA tensors in -> CUDA tensors out
    tl, tb = glass_b200.displace(torch.as_tensor(lon).to(cuda_device), torch.as_tensor(lat).to(cuda_device), torch.as_tensor(alpha).to(cuda_device))
    assert tl.is_cuda and np.array_equal(tl.cpu().numpy(), glass_b200.displace(lon, lat, alpha)[0])


def test_displace_roundtrip_large(cuda_device):
    """Size-independent property on 2e7 points: displacing by the displacement between two
    positions lands on the second one."""
    import glass_b200

    n = 20_000_000
    g = torch.Generator(device=cuda_device)
    g.manual_seed(3)
    lon = 360.0 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=g)
    lat = torch.rad2deg(torch.asin(2.0 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=g) - 1.0))
    lon2 = 360.0 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=g)
    lat2 = torch.rad2deg(torch.asin(2.0 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=g) - 1.0))
    a = glass_b200.displacement(lon, lat, lon2, lat2)
    lo, la = glass_b200.displace(lon, lat, a)

    def unit(lo, la):
        lo, la = torch.deg2rad(lo), torch.deg2rad(la)
        return torch.stack([torch.cos(la) * torch.cos(lo), torch.cos(la) * torch.sin(lo), torch.sin(la)])

    assert (unit(lo, la) - unit(lon2, lat2)).abs().max().item() < 1e-12


def test_write_catalog_from_cuda_columns(cuda_device, tmp_path):
    """Catalogue sink fed with CUDA tensors (double-buffered pinned staging on a side stream):
    rows arrive in write order and bit-identical, including a mixed host/CUDA call."""
    import glass_b200
    from glass_b200.user import read_catalog

    g = torch.Generator(device=cuda_device)
    g.manual_seed(11)
    path = tmp_path / "cat.fits"
    parts = []
    with glass_b200.write_catalog(path, ext="GALAXIES") as out:
        for i, n in enumerate((1000, 250_000, 0, 3, 90_000, 1_000_000)):
            lon = 360.0 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=g)
            lat = 180.0 * torch.rand(n, dtype=torch.float64, device=cuda_device, generator=g) - 90.0
            she = torch.view_as_complex(torch.randn((n, 2), dtype=torch.float64, device=cuda_device, generator=g))
            z = np.full(n, 0.1 * i) if i == 3 else torch.full((n,), 0.1 * i, dtype=torch.float64, device=cuda_device)
            out.write(RA=lon, DEC=lat, Z_TRUE=z, G=she)
            parts.append((lon.cpu().numpy(), lat.cpu().numpy(), np.full(n, 0.1 * i), she.cpu().numpy()))
            del lon, lat, she  # the writer must hold on to what it still copies
    cat = read_catalog(path)
    for k, name in enumerate(("RA", "DEC", "Z_TRUE", "G")):
        assert np.array_equal(cat[name], np.concatenate([p[k] for p in parts]))
    assert cat["__extname__"] == "GALAXIES" and path.stat().st_size % 2880 == 0
