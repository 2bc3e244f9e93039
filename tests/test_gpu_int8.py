"""GPU parity of the INT8 tensor-core Legendre path (csrc/sht_ozaki.cu: the contraction over l as exact
integer digit products on tcgen05.mma kind::i8) against the oracle and against the FP64 kernel.
Tolerance: maps from identical alm within 1e-10 relative (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import healpix_ref as H

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def random_alm(lmax, seed, nmaps=1, power_law=True):
    rng = np.random.default_rng(seed)
    n = H.alm_size(lmax)
    a = rng.standard_normal((nmaps, n)) + 1j * rng.standard_normal((nmaps, n))
    a[:, : lmax + 1] = a[:, : lmax + 1].real
    if power_law:  # the bench's spectrum: C_l ~ (l + 1)^-1.5, different amplitudes per map
        l = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])
        a *= np.sqrt(1e-2 * (l + 1.0) ** -1.5) * (1.0 + np.arange(nmaps))[:, None]
    return a


def relerr(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture
def int8_plans():
    """Plans created inside the test use the INT8 path for groups of four and eight maps at any nside."""
    import os

    from glass_b200.healpix import clear_plans

    clear_plans()
    old = os.environ.get("GLB_LEGENDRE")
    os.environ["GLB_LEGENDRE"] = "int8"
    yield
    if old is None:
        del os.environ["GLB_LEGENDRE"]
    else:
        os.environ["GLB_LEGENDRE"] = old
    clear_plans()


@pytest.mark.parametrize("nside,lmax", [(1, 2), (2, 5), (3, 7), (8, 23), (16, 47), (32, 64), (64, 191), (48, 100), (256, 300)])
@pytest.mark.parametrize("nmaps", [4, 8])
def test_int8_phases_match_fp64_kernel(cuda_device, nside, lmax, nmaps):
    """Stage tap: F_m(ring) of the two Legendre kernels on the same alm (every m <= mlim(ring))."""
    from glass_b200 import _lib
    from glass_b200.healpix import get_plan

    alm = torch.as_tensor(random_alm(lmax, 7 * nside + nmaps, nmaps)).to(cuda_device)
    pl = get_plan(nside, lmax, 4, cuda_device)
    nring = 4 * nside - 1
    ref = torch.zeros((nmaps, nring, lmax + 1), dtype=torch.complex128, device=cuda_device)
    got = torch.full_like(ref, float("nan"))
    for b0 in range(0, nmaps, 4):
        _lib.check(pl.lib.glb_debug_alm2phase(pl.handle, alm[b0:].data_ptr(), 4, ref[b0:].data_ptr(), pl.stream_ptr()), "fp64")
    _lib.check(pl.lib.glb_debug_alm2phase_int8(pl.handle, alm.data_ptr(), nmaps, got.data_ptr(), pl.stream_ptr()), "int8")
    torch.cuda.synchronize()
    assert torch.isfinite(torch.view_as_real(got)).all()
    for b in range(nmaps):
        err = float((got[b] - ref[b]).abs().max() / ref[b].abs().max())
        assert err < 2e-11, (b, err)


@pytest.mark.parametrize("nside,lmax,nmaps", [(4, 11, 4), (16, 47, 8), (32, 64, 9), (64, 191, 13), (128, 383, 8), (20, 50, 12)])
def test_int8_alm2map_vs_oracle(cuda_device, int8_plans, nside, lmax, nmaps):
    """glb_alm2map with the INT8 path forced: groups of eight and four (the remainder runs on the FP64 pipe)."""
    from glass_b200.healpix import alm2map_batch, get_plan

    alm = random_alm(lmax, 100 + nside, nmaps)
    assert get_plan(nside, lmax, 4, cuda_device).legendre_mode == "int8"
    maps = alm2map_batch(torch.as_tensor(alm).to(cuda_device), nside, lmax).cpu().numpy()
    for b in range(nmaps):
        ref = H.alm2map(alm[b], nside, lmax)
        assert relerr(maps[b], ref) < RTOL, (b, relerr(maps[b], ref))


def test_int8_fused_transforms_and_generate(cuda_device, int8_plans):
    """Eight lognormal shells through generate() (one INT8 group with the fused lognormal store) against the oracle
    on supplied deviates."""
    import glass_b200
    from glass_b200.rng import Deviates
    from oracle import glass_ref as G

    nside, lmax, nshell, ncorr = 32, 64, 8, 2
    l = np.arange(lmax + 1)
    g = 1e-2 * (l + 1.0) ** -1.5
    g[0] = 0.0
    gls = [0.5 ** (i - j) * g if i - j <= ncorr else g[:0] for i in range(nshell) for j in range(i, -1, -1)]
    rng = np.random.default_rng(3)
    n = (lmax + 1) * (lmax + 2) // 2
    zs = [rng.standard_normal((n, 2)) @ np.array([1, 1j]) for _ in range(nshell)]
    before = glass_b200._lib.load().glb_kernel_launch_count()
    got = list(glass_b200.generate([glass_b200.grf.Lognormal(1.0)] * nshell, gls, nside, ncorr=ncorr, rng=Deviates(normal_alm=zs)))
    assert glass_b200._lib.load().glb_kernel_launch_count() > before
    ref = G.generate([("lognormal", 1.0)] * nshell, gls, nside, ncorr, zs)
    for a, b in zip(got, ref):
        assert relerr(a, b) < RTOL, relerr(a, b)


def test_int8_alm2map_vs_long_double_oracle(cuda_device, int8_plans):
    """nside 512, lmax 1023 against the 80-bit C oracle (range scaling, silent warps, runs across tiles all active)."""
    from glass_b200.healpix import alm2map_batch
    from oracle import sht_c

    nside, lmax = 512, 1023
    alm = random_alm(lmax, nside, 8)
    got = alm2map_batch(torch.as_tensor(alm).to(cuda_device), nside, lmax).cpu().numpy()
    for b in (0, 5):
        ref = sht_c.alm2map(alm[b], nside, lmax, long_double=True)
        assert relerr(got[b], ref) < 5e-11, (b, relerr(got[b], ref))


@pytest.mark.parametrize("nside,lmax", [(1024, 2047), (2048, 4095)])
def test_int8_alm2map_at_configured_sizes(cuda_device, nside, lmax):
    """BASELINE.json configs[1] and [2] in the DEFAULT mode (groups of eight maps take the INT8 path at nside >= 1024):
    random alm against the CPU arm's transform (oracle/sht_fast.cpp) and against the FP64 kernel."""
    from glass_b200.healpix import alm2map_batch, clear_plans, get_plan
    from oracle import sht_c

    clear_plans()
    alm = random_alm(lmax, 17 * nside, 8, power_law=(nside == 1024))
    d_alm = torch.as_tensor(alm).to(cuda_device)
    assert get_plan(nside, lmax, 4, cuda_device).legendre_mode == "auto"
    got = alm2map_batch(d_alm, nside, lmax)
    for b in (1, 6):
        ref = sht_c.alm2map_fast(alm[b], nside, lmax)
        assert relerr(got[b].cpu().numpy(), ref) < RTOL, (b, relerr(got[b].cpu().numpy(), ref))
    fp64 = alm2map_batch(d_alm[:4], nside, lmax)  # four maps: the FP64 kernel
    for b in range(4):
        err = float((got[b] - fp64[b]).abs().max() / fp64[b].abs().max())
        assert err < 5e-11, (b, err)
    clear_plans()


def test_int8_single_harmonics_fullsize(cuda_device):
    """nside 4096, lmax 8191, eight maps of single harmonics (zonal, sectoral, mixed, l up to lmax) through the
    default path (INT8 for a group of eight) against lambda_lm in 80-bit arithmetic on exact ring geometry."""
    from test_gpu_fullsize import LMAX, NSIDE, _single_harmonics_error

    modes = [(0, 0, 1.0 + 0j), (1, 0, -0.7 + 0j), (8000, 0, 0.9 + 0j), (8191, 0, 0.4 + 0j), (8191, 8191, 1.1 + 0.5j),
             (5000, 3000, -0.6 + 0.8j), (8191, 4000, 0.5 + 0.1j), (7000, 6999, 0.2 - 0.9j), (6001, 17, 0.3 + 0.3j)]
    err = _single_harmonics_error(cuda_device, NSIDE, LMAX, modes, nmaps=8)
    assert err < RTOL, err
