"""GPU parity: visibility-mask construction (glass/observations.py:51-101 on hp.query_strip and
hp.Rotator.rotate_map_pixel) and the device forms of discretized_cls / effective_cls against the
oracle and the golden vectors made by executing the reference's source."""
import os

import numpy as np
import pytest
import torch

from oracle import healpix_ref as H

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz")


@pytest.mark.parametrize("nside", [1, 2, 4, 16, 3, 64])
def test_query_strip_vs_oracle(cuda_device, nside):
    from glass_b200 import healpix as hp

    rng = np.random.default_rng(nside)
    cases = [(30, 90), (20, 80), (0, 0), (0.0, np.pi), (np.pi, 0.0), (1.0, 1.0), (2.5, 0.3)]
    cases += [tuple(rng.uniform(0, np.pi, 2)) for _ in range(6)]
    for t in cases:
        got = hp.query_strip(nside, t)
        assert got.dtype == np.int64 and got.shape == (12 * nside**2,)
        assert np.array_equal(got, H.query_strip(nside, *t).astype(np.int64)), t
    d = hp.query_strip(nside, (0.4, 1.9), dtype=np.float64)
    assert d.dtype == np.float64
    t = hp.query_strip(nside, (0.4, 1.9), dtype=torch.float64, xp=torch)
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), d)
    # ring semantics: a pixel is in iff its centre colatitude is (checked away from ring boundaries)
    theta, _ = H.pix2ang_centers(nside)
    inside = (theta > 0.4 + 1e-9) & (theta < 1.9 - 1e-9)
    assert np.all(d[inside] == 1)
    assert np.all(d[(theta < 0.4 - 1e-9) | (theta > 1.9 + 1e-9)] == 0)


@pytest.mark.parametrize("nside", [1, 2, 8, 32, 5])
@pytest.mark.parametrize("coord", ["GC", "CE", "EG", "CG", "CC"])
def test_rotate_map_pixel_vs_oracle(cuda_device, nside, coord):
    from glass_b200 import healpix as hp

    m = np.random.default_rng(nside).random(12 * nside**2)
    got = hp.Rotator(coord=coord).rotate_map_pixel(m)
    ref = H.rotate_map_pixel(m, coord)
    # interpolation weights agree to rounding; a back-rotated centre that falls within rounding of a
    # ring or pixel boundary may pick the neighbouring cell, where bilinear interpolation is continuous
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-10)
    dev = hp.Rotator(coord=coord).rotate_map_pixel(torch.as_tensor(m, device=cuda_device))
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), got)
    if coord == "CC":
        np.testing.assert_allclose(got, m, rtol=0, atol=1e-13)
    # constants are reproduced exactly up to the weights' rounding
    one = hp.Rotator(coord=coord).rotate_map_pixel(np.ones_like(m))
    np.testing.assert_allclose(one, 1.0, rtol=0, atol=4e-16)


def test_rotation_matrices_are_rotations():
    from glass_b200.healpix import _coordconv_matrix

    for c in ("GC", "CG", "CE", "EC", "EG", "GE"):
        M = _coordconv_matrix(c)
        assert np.abs(M @ M.T - np.eye(3)).max() < 2e-9 and abs(np.linalg.det(M) - 1) < 2e-9
        assert np.allclose(_coordconv_matrix(c[::-1]), np.linalg.inv(M), atol=1e-15)
        assert np.array_equal(M, H.coordconv_matrix(c))
    # the galactic pole in equatorial coordinates (J2000: RA 192.86 deg, Dec 27.13 deg)
    x, y, z = _coordconv_matrix("GC") @ np.array([0.0, 0.0, 1.0])
    assert abs(np.degrees(np.arctan2(y, x)) % 360 - 192.86) < 0.01 and abs(np.degrees(np.arcsin(z)) - 27.13) < 0.01
    with pytest.raises(TypeError):
        _coordconv_matrix("XY")


def test_rotate_smooth_function(cuda_device):
    """Rotating the map of a smooth function f(direction) gives f(R^-1 direction) up to the O(h^2)
    interpolation error away from the poles of the INPUT grid."""
    from glass_b200 import healpix as hp

    a = np.array([0.3, -0.5, 0.8])
    errs = []
    for nside in (32, 64, 128):
        th, ph = H.pix2ang_centers(nside)
        v = np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)])
        f = a @ v
        got = hp.Rotator(coord="GC").rotate_map_pixel(f)
        back = np.linalg.inv(H.coordconv_matrix("GC")) @ v
        away = np.abs(back[2]) < 0.9
        errs.append(np.abs(got - a @ back)[away].max())
    assert errs[0] < 2e-3 and errs[1] < errs[0] / 3 and errs[2] < errs[1] / 3, errs


@pytest.mark.parametrize("nside", [4, 16, 64])
def test_vmap_galactic_ecliptic_vs_oracle(cuda_device, nside):
    """The reference's own test (tests/core/test_observations.py:20-52) plus parity with the oracle."""
    import glass_b200 as glass

    vmap = glass.vmap_galactic_ecliptic(nside)
    assert isinstance(vmap, np.ndarray) and vmap.shape[0] == 12 * nside**2
    np.testing.assert_allclose(vmap, H.vmap_galactic_ecliptic(nside), rtol=0, atol=1e-10)
    assert vmap.min() >= 0 and vmap.max() <= 1 + 1e-15 and 0.2 < vmap.mean() < 0.8
    t = glass.vmap_galactic_ecliptic(nside, xp=torch)
    assert t.is_cuda and np.array_equal(t.cpu().numpy(), vmap)
    # no rotation
    z = glass.vmap_galactic_ecliptic(nside, galactic=(0, 0), ecliptic=(0, 0))
    assert np.array_equal(z, np.zeros_like(z))
    other = glass.vmap_galactic_ecliptic(nside, galactic=(0.5, 1.1), ecliptic=(2.0, 1.0))
    np.testing.assert_allclose(other, H.vmap_galactic_ecliptic(nside, (0.5, 1.1), (2.0, 1.0)), rtol=0, atol=1e-10)
    with pytest.raises(TypeError, match="galactic stripe must be a pair of numbers"):
        glass.vmap_galactic_ecliptic(nside, galactic=(1,))
    with pytest.raises(TypeError, match="ecliptic stripe must be a pair of numbers"):
        glass.vmap_galactic_ecliptic(nside, ecliptic=(1,))
    with pytest.raises(TypeError, match="galactic stripe must be a pair of numbers"):
        glass.vmap_galactic_ecliptic(nside, galactic=(1, 2, 3))
    with pytest.raises(TypeError, match="ecliptic stripe must be a pair of numbers"):
        glass.vmap_galactic_ecliptic(nside, ecliptic=(1, 2, 3))


def test_vmap_fullsize_properties(cuda_device):
    """nside 4096 (the size of BASELINE configs 4/5): values in [0, 1], the visible fraction equal to
    the nside-64 map's within the strips' edge width, and the mask usable by positions_from_delta."""
    import glass_b200 as glass

    v = glass.vmap_galactic_ecliptic(4096, xp=torch)
    assert v.shape == (12 * 4096**2,) and float(v.min()) >= 0 and float(v.max()) <= 1 + 1e-15
    small = glass.vmap_galactic_ecliptic(64)
    assert abs(float(v.mean()) - small.mean()) < 0.02
    assert float(((v > 1e-9) & (v < 1 - 1e-9)).double().mean()) < 0.01  # partial values only along the strip edges


def test_discretized_and_effective_cls_on_device(cuda_device):
    """glass/fields.py:239-300, 607-694 with CUDA spectra: the kernels' output is bit-identical to the
    vectors made by executing the reference's source."""
    import glass_b200
    from helpers import synthetic_gls

    gold = np.load(GOLD)
    gls_h = synthetic_gls(4, 12, 3)
    gls = [torch.as_tensor(g, device=cuda_device) for g in gls_h]
    pw = torch.as_tensor(gold["dcl_pw"], device=cuda_device)
    before = glass_b200._lib.load().glb_kernel_launch_count()
    res = glass_b200.discretized_cls(gls, lmax=9, ncorr=2, nside=4, pixwin=pw)
    assert glass_b200._lib.load().glb_kernel_launch_count() == before + 1  # all spectra in one launch
    assert all(r.is_cuda for r in res)
    assert np.array_equal(np.array([r.shape[0] for r in res]), gold["dcl_all_len"])
    assert np.array_equal(torch.cat(res).cpu().numpy(), gold["dcl_all"])
    res = glass_b200.discretized_cls(gls, lmax=9, ncorr=2, nside=4, pixwin=gold["dcl_pw"])  # host table
    assert np.array_equal(torch.cat(res).cpu().numpy(), gold["dcl_all"])
    for tag, kw in {"lmax": {"lmax": 8}, "ncorr": {"ncorr": 1}}.items():
        res = glass_b200.discretized_cls(gls, **kw)
        assert np.array_equal(torch.cat(res).cpu().numpy(), gold[f"dcl_{tag}"])
    # ragged spectra, a window shorter than the spectra
    rag = [g[: 5 + i] for i, g in enumerate(gls)]
    res = glass_b200.discretized_cls(rag, nside=4, pixwin=pw[:9])
    for r, g in zip(res, rag):
        n = min(g.shape[0], 9)
        assert np.array_equal(r.cpu().numpy(), (g[:n].cpu().numpy() * gold["dcl_pw"][:n] ** 2))

    w1 = torch.as_tensor(gold["ecl_w1"], device=cuda_device)
    w2 = torch.as_tensor(gold["ecl_w2"], device=cuda_device)
    auto = glass_b200.effective_cls(gls, w1)
    assert auto.is_cuda and np.array_equal(auto.cpu().numpy(), gold["ecl_auto"])
    cross = glass_b200.effective_cls(gls, w1, w2, lmax=7)
    assert np.array_equal(cross.cpu().numpy(), gold["ecl_cross"])
    # device weights with host spectra, and padding beyond the longest spectrum
    assert np.array_equal(glass_b200.effective_cls(gls_h, w1).cpu().numpy(), gold["ecl_auto"])
    long = glass_b200.effective_cls(gls, w1, w2, lmax=20).cpu().numpy()
    assert np.array_equal(long[..., :8], gold["ecl_cross"]) and np.all(long[..., 13:] == 0)
    with pytest.raises(ValueError, match="shape mismatch between fields and weights1"):
        glass_b200.effective_cls(gls, torch.ones((3, 2), device=cuda_device))
    with pytest.raises(ValueError, match="shape mismatch between fields and weights2"):
        glass_b200.effective_cls(gls, w1, torch.ones((3, 2), device=cuda_device))
