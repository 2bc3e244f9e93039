"""GPU parity of the WHOLE user loop on BASELINE.json configs[0] (10 lognormal matter shells,
nside 128, lmax 383; examples/2-advanced/stage_4_galaxies.ipynb cell 13) against the
oracle with identical supplied deviates, the binding stub of INTEGRATION.md executed verbatim, and a
guarded comparison with a real healpy when one is importable."""
import os
import re

import numpy as np
import pytest
import torch

from helpers import MockCosmology, synthetic_gls, triangular_shells
from oracle import glass_ref as G
from oracle import healpix_ref as H

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config1_chain_against_oracle(cuda_device):
    """generate (correlated lognormal shells) -> MultiPlaneConvergence.add_window -> shear_from_convergence
    -> positions_from_delta -> redshifts -> ellipticity_intnorm -> galaxy_shear, every stage fed by the
    previous one's OUTPUT on both sides, with supplied normal / Poisson / uniform deviates:
    maps within 1e-10 relative, pixel indices, counts and batch sizes bit-exact, per-galaxy values
    within 1e-9 of their scale (they gather the maps compared before)."""
    import glass_b200 as glass
    from glass_b200.rng import Deviates

    nshell, nside, lmax, ncorr = 10, 128, 383, 3
    npix = 12 * nside * nside
    rng = np.random.default_rng(2024)
    gls = synthetic_gls(nshell, lmax, ncorr)
    nalm = (lmax + 1) * (lmax + 2) // 2
    zs = [rng.standard_normal((nalm, 2)) @ np.array([1, 1j]) for _ in range(nshell)]
    shells = triangular_shells(nshell, dz=0.1)
    cosmo = MockCosmology()
    fields = glass.lognormal_fields(shells)
    ref_maps = G.generate([("lognormal", 1.0)] * nshell, gls, nside, ncorr, zs)
    conv, ref_conv = glass.MultiPlaneConvergence(cosmo), G.MultiPlaneConvergence(cosmo)
    ngal = 0.02  # per arcmin^2: ~0.25 galaxies per pixel
    ntot = 0
    for i, delta in enumerate(glass.generate(fields, gls, nside, ncorr=ncorr, rng=Deviates(normal_alm=zs))):
        assert np.abs(delta - ref_maps[i]).max() <= 1e-10 * np.abs(ref_maps[i]).max(), i
        # lensing: the recurrence is fed with each side's own matter plane
        conv.add_window(delta, shells[i])
        ref_conv.add_window(ref_maps[i], shells[i].za, shells[i].wa, shells[i].zeff)
        kappa, ref_kappa = conv.kappa, ref_conv.kappa
        assert np.abs(kappa - ref_kappa).max() <= 1e-10 * max(np.abs(ref_kappa).max(), 1e-300), i
        if i in (1, 5, 9):  # three planes through the transforms (the oracle's pure-NumPy spin transform is slow)
            g1, g2 = glass.shear_from_convergence(kappa, lmax, discretized=False, niter=1)
            r1, r2 = G.shear_from_convergence(ref_kappa, lmax, niter=1)
            scale = np.abs(r1).max()
            assert np.abs(g1 - r1).max() <= 1e-10 * scale and np.abs(g2 - r2).max() <= 1e-10 * scale, i
        else:
            continue
        # galaxies: supplied Poisson counts (drawn from the ORACLE's expected counts) and in-pixel offsets
        lam = np.clip(G.expected_count(ref_maps[i], ngal, 1.2), 0, None)
        counts = rng.poisson(lam)
        uvs = []

        def uv(n, uvs=uvs):
            u, v = rng.random(n), rng.random(n)
            uvs.append((u, v))
            return u, v

        got = list(glass.positions_from_delta(ngal, delta, 1.2, batch=20_000, rng=Deviates(poisson=[counts], uv=uv)))
        it = iter(uvs)
        ref = G.positions_from_counts(counts, nside, 20_000, lambda n: next(it))
        assert [c for _lo, _la, c in got] == [c for _lo, _la, c in ref] and sum(c for _lo, _la, c in got) == counts.sum()
        for (lon, lat, cnt), (rlon, rlat, _c) in zip(got, ref):
            assert np.abs(lon - rlon).max() <= 1e-10 and np.abs(lat - rlat).max() <= 1e-10
            assert np.array_equal(H.ang2pix(nside, lon, lat, lonlat=True), H.ang2pix(nside, rlon, rlat, lonlat=True))
            u = rng.random(cnt)
            z = glass.redshifts(cnt, shells[i], rng=Deviates(uniform=u))
            assert np.abs(z - G.redshifts_from_nz_uniform(shells[i].za, shells[i].wa, u)).max() <= 1e-13
            nrm = rng.standard_normal((cnt, 2)) @ np.array([1, 1j])
            eps = glass.ellipticity_intnorm(cnt, 0.27, rng=Deviates(normal=nrm))
            assert np.abs(eps - G.ellipticity_intnorm_from_normals(0.27, nrm)).max() <= 1e-14
            she = glass.galaxy_shear(lon, lat, eps, kappa, g1, g2)
            ref_she = G.galaxy_shear(rlon, rlat, eps, ref_kappa, r1, r2)
            assert np.abs(she - ref_she).max() <= 1e-9 * max(np.abs(ref_she).max(), 1.0)
            ntot += cnt
    assert ntot > 100_000


def test_integration_stub_verbatim(cuda_device, monkeypatch):
    """The ctypes stub of INTEGRATION.md section 2 -- the code a GLASS maintainer would drop into
    glass/_b200.py -- executed as written (only the library's search path is resolved), binding
    glb_plan_create / glb_alm2map_host on HOST buffers, against the oracle and the device-buffer API;
    a map larger than the entry's page-locked staging chunks exercises the pipelined copies."""
    import ctypes as C

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "glb_alm2map_host" in b)
    real_cdll = C.CDLL
    monkeypatch.setattr(C, "CDLL", lambda name, *a, **k: real_cdll(os.path.join(ROOT, "glass_b200", name) if name == "libglassb200.so" else name, *a, **k))
    ns: dict = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)
    rng = np.random.default_rng(5)
    for nside, lmax in ((8, 23), (64, 150)):
        n = (lmax + 1) * (lmax + 2) // 2
        alm = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        alm[: lmax + 1] = alm[: lmax + 1].real
        got = ns["alm2map"](alm, nside)
        ref = H.alm2map(alm, nside, lmax)
        assert isinstance(got, np.ndarray) and np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()
    # 100 MB of map, 34 MB of alm: several staging chunks in each direction
    from glass_b200.healpix import alm2map_batch

    nside, lmax = 1024, 2047
    n = (lmax + 1) * (lmax + 2) // 2
    alm = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    got = ns["alm2map"](alm, nside)
    dev = alm2map_batch(torch.as_tensor(alm[None]).to(cuda_device), nside, lmax)[0].cpu().numpy()
    assert np.array_equal(got, dev)


def test_against_real_healpy_when_available(cuda_device):
    """healpy / healpix are absent from the offline image, so parity with their own arithmetic is
    unpinned.  Wherever they ARE importable this test pins it: alm2map, the sign and
    order conventions of alm2map_spin (spin 1 and 2), map2alm with healpy's iterations (uniform
    weights), pixwin (T and P) and ang2pix / the in-pixel offsets of healpix.randang's kernel."""
    healpy = pytest.importorskip("healpy")
    from glass_b200 import healpix as hp

    nside, lmax = 64, 150
    rng = np.random.default_rng(9)
    n = (lmax + 1) * (lmax + 2) // 2
    alm = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    alm[: lmax + 1] = alm[: lmax + 1].real
    ref = healpy.alm2map(alm, nside, lmax=lmax, pol=False)
    got = hp.alm2map(alm, nside, pol=False)
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()
    for spin in (1, 2):
        e = alm.copy()
        for l in range(spin):
            for m in range(l + 1):
                e[H.alm_index(lmax, l, m)] = 0
        r1, r2 = healpy.alm2map_spin([e, np.zeros_like(e)], nside, spin, lmax)
        g1, g2 = hp.alm2map_spin([e, None], nside, spin, lmax)
        scale = np.abs(r1).max()
        assert np.abs(g1 - r1).max() <= 1e-10 * scale and np.abs(g2 - r2).max() <= 1e-10 * scale, spin
    ra = healpy.map2alm(ref, lmax=lmax, pol=False, iter=3, use_weights=False, use_pixel_weights=False)
    ga = hp.map2alm(ref, lmax=lmax, pol=False, niter=3)
    assert np.abs(ga - ra).max() <= 1e-9 * np.abs(ra).max()
    try:
        wt, wp = healpy.pixwin(nside, pol=True, lmax=lmax)
    except Exception:  # data files missing
        wt = None
    if wt is not None:
        gt, gp = hp.pixwin(nside, lmax=lmax, pol=True)
        assert np.abs(gt - wt).max() <= 1e-5 and np.abs(gp[2:] - wp[2:]).max() <= 1e-4
    lon, lat = rng.uniform(0, 360, 1000), np.degrees(np.arcsin(rng.uniform(-1, 1, 1000)))
    assert np.array_equal(hp.ang2pix(nside, lon, lat, lonlat=True), healpy.ang2pix(nside, lon, lat, lonlat=True))
