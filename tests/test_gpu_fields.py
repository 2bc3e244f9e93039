"""GPU parity: correlated alm sampling + generate() against the oracle restatement."""
import numpy as np
import pytest
import torch

from oracle import glass_ref as G
from oracle import healpix_ref as H

pytestmark = pytest.mark.gpu


def synthetic_gls(nshell, lmax, ncorr, ragged=False):
    """g_l = 1e-2 (l+1)^-1.5 (l>=1), cross 0.5^|i-j| within ncorr."""
    l = np.arange(lmax + 1)
    g = 1e-2 * (l + 1.0) ** -1.5
    g[0] = 0.0
    gls = []
    for i in range(nshell):
        for j in range(i, -1, -1):
            d = i - j
            if d <= ncorr:
                gl = 0.5**d * g
                if ragged and d == 1:
                    gl = gl[: lmax // 2]
                gls.append(gl)
            else:
                gls.append(np.zeros(0))
    return gls


def supplied_z(nshell, lmax, seed=42):
    rng = np.random.default_rng(seed)
    n = (lmax + 1) * (lmax + 2) // 2
    return [rng.standard_normal((n, 2)) @ np.array([1, 1j]) for _ in range(nshell)]


@pytest.mark.parametrize("nshell,lmax,ncorr,ragged", [(4, 16, 2, False), (5, 33, None, False), (6, 20, 1, True), (3, 8, 0, False), (70, 6, None, False)])
def test_alm_with_supplied_z(cuda_device, nshell, lmax, ncorr, ragged):
    """alm given identical z.  K2 (draw order, combine, re-ordering, m = 0 fix) is BIT-EXACT when it
    is handed the oracle's weights through the C ABI (products and sums rounded as NumPy does,
    any number of terms -- the 70-shell case crosses the 64-pointer launch chunk); through
    ``generate``'s sampler the weights come from K1, whose summation order differs from NumPy's
    matmul: 1e-13 of the largest alm."""
    import ctypes as C

    from glass_b200 import _lib
    from glass_b200.fields import _ShellSampler
    from glass_b200.rng import Deviates

    nc = nshell - 1 if ncorr is None else ncorr
    gls = synthetic_gls(nshell, lmax, nc, ragged)
    if nshell > 64:  # 70 fully correlated shells, weakly enough to stay positive definite
        gls = [gls[0] * (1.0 if i == j else 0.02) for i in range(nshell) for j in range(i, -1, -1)]
    zs = supplied_z(nshell, lmax)
    ref = G.generate_alms(gls, ncorr, zs)
    s = _ShellSampler(gls, 8, ncorr, Deviates(normal_alm=zs), cuda_device)
    for j in range(nshell):
        out = torch.empty(s.nalm, dtype=torch.complex128, device=cuda_device)
        assert s.next_alm(out) == j
        got = out.cpu().numpy()
        assert np.abs(got - ref[j]).max() <= 1e-13 * np.abs(ref[j]).max(), (j, np.abs(got - ref[j]).max())
    assert s.next_alm(out) is None
    # K2 alone, on the oracle's weights: bit for bit
    lib = _lib.load()
    st = torch.cuda.current_stream(cuda_device).cuda_stream
    n = lmax + 1
    ws = G.iternorm(G.cls2cov_rows(gls, n, nshell, nc))
    zd = []
    for z in zs:
        zg = torch.as_tensor(z).to(cuda_device)
        zm = torch.empty_like(zg)
        _lib.check(lib.glb_alm_glass_to_healpix(lmax, zg.data_ptr(), zm.data_ptr(), st), "glb_alm_glass_to_healpix")
        zd.append(zm)
    for j in range(nshell):
        nterms = min(j + 1, nc + 1)
        w = torch.as_tensor(np.ascontiguousarray(ws[j][:, nc + 1 - nterms :])).to(cuda_device)
        ptrs = (C.c_void_p * nterms)(*[zd[t].data_ptr() for t in range(j - nterms + 1, j + 1)])
        out = torch.empty(s.nalm, dtype=torch.complex128, device=cuda_device)
        _lib.check(lib.glb_alm_combine(lmax, nterms, ptrs, w.data_ptr(), nterms, out.data_ptr(), st), "glb_alm_combine")
        assert np.array_equal(out.cpu().numpy(), ref[j]), j


@pytest.mark.parametrize("nside,lmax,nshell,ncorr", [(16, 32, 5, 2), (32, 64, 6, 3), (8, 23, 3, None)])
def test_generate_lognormal_vs_oracle(cuda_device, nside, lmax, nshell, ncorr):
    import glass_b200
    from glass_b200.rng import Deviates

    nc = nshell - 1 if ncorr is None else ncorr
    gls = synthetic_gls(nshell, lmax, nc)
    zs = supplied_z(nshell, lmax, seed=7)
    fields = [glass_b200.grf.Lognormal(1.0 if i % 2 == 0 else 0.7) for i in range(nshell)]
    ref = G.generate([("lognormal", f.lamda) for f in fields], gls, nside, ncorr, zs)
    got = list(glass_b200.generate(fields, gls, nside, ncorr=ncorr, rng=Deviates(normal_alm=zs)))
    assert len(got) == nshell
    for j in range(nshell):
        assert isinstance(got[j], np.ndarray)
        err = np.abs(got[j] - ref[j]).max() / np.abs(ref[j]).max()
        assert err < 1e-10, (j, err)
    # device-resident variant
    gls_d = [torch.as_tensor(g).to(cuda_device) for g in gls]
    got_d = list(glass_b200.generate(fields, gls_d, nside, ncorr=ncorr, rng=Deviates(normal_alm=zs)))
    for j in range(nshell):
        assert got_d[j].is_cuda
        assert np.array_equal(got_d[j].cpu().numpy(), got[j])


def test_generate_mixed_transforms_and_custom(cuda_device):
    import glass_b200
    from glass_b200.rng import Deviates

    nside, lmax, nshell = 8, 16, 3
    gls = synthetic_gls(nshell, lmax, 2)
    zs = supplied_z(nshell, lmax, seed=3)

    class Custom:
        def __call__(self, x, var, /):
            return 2.0 * x + var

    fields = [glass_b200.grf.Normal(), glass_b200.grf.SquaredNormal(0.3, 1.5), Custom()]
    got = list(glass_b200.generate(fields, gls, nside, ncorr=2, rng=Deviates(normal_alm=zs)))
    alms = G.generate_alms(gls, 2, zs)
    x = [H.alm2map(a, nside) for a in alms]
    ref = [x[0], G.squared_normal(x[1], 0.3, 1.5), 2.0 * x[2] + G.cltovar(G.getcl(gls, 2, 2))]
    for j in range(3):
        assert np.abs(got[j] - ref[j]).max() / np.abs(ref[j]).max() < 1e-10


def test_generate_errors(cuda_device):
    import glass_b200

    with pytest.raises(ValueError, match="mismatch between number of fields and gls"):
        next(glass_b200.generate([glass_b200.grf.Normal()], [np.ones(3), np.ones(3)], 4))
    with pytest.raises(ValueError, match="all gls are empty"):
        next(glass_b200.generate([glass_b200.grf.Normal()], [np.zeros(0)], 4))
    with pytest.raises(ValueError, match="negative values in cl"):
        next(glass_b200.generate([glass_b200.grf.Normal()], [-np.ones(3)], 4))
    # not positive definite at the second shell: first shell is still yielded
    gls = [np.ones(4), np.ones(4), 2 * np.ones(4)]
    g = glass_b200.generate([glass_b200.grf.Normal()] * 2, gls, 4)
    next(g)
    with pytest.raises(ValueError, match="covariance matrix is not positive definite"):
        next(g)


@pytest.mark.parametrize("nshell,lmax,ncorr", [(14, 40, 9), (12, 25, 11), (5, 12, 0), (9, 300, 8), (6, 64, 3)])
def test_device_iternorm_vs_oracle(cuda_device, nshell, lmax, ncorr):
    """K1 (glb_iternorm_step) against the oracle's recursion (glass/fields.py:101-188) on the rows
    of cls2cov, shell by shell.  Tolerance 1e-11 relative: the summation order of NumPy's
    batched matmul is not specified, so this path is not bit-exact by construction."""
    from glass_b200.fields import _DeviceIterNorm, cls2cov

    gls = synthetic_gls(nshell, lmax, ncorr, ragged=True)
    want = G.iternorm(G.cls2cov_rows(gls, lmax + 1, nshell, ncorr))
    dn = _DeviceIterNorm(lmax + 1, ncorr, nshell, cuda_device)
    for j, row in enumerate(cls2cov(gls, lmax + 1, nshell, ncorr)):
        w = dn.step(row).cpu().numpy()
        assert w.shape == want[j].shape
        assert np.abs(w - want[j]).max() <= 1e-11 * np.abs(want[j]).max(), j
    assert dn.first_failure() is None


def test_iternorm_public_function(cuda_device):
    """glass.iternorm (glass/fields.py:101-188) through the public generator: the reference's
    known answers (golden vectors from executing its source; tests/core/test_fields.py:96-170:
    explicit Cholesky, leading dimensions, errors), NumPy in -> NumPy out, CUDA in -> CUDA out."""
    import os

    import glass_b200

    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_vectors.npz"))
    cov = np.array([[1.0, 0.2, 0.1], [0.2, 0.5, 0.2], [0.1, 0.2, 0.3]])
    for k in (0, 1, 2):
        rows = [np.pad(cov[i, i::-1][: min(i, k) + 1], (0, k + 1 - min(i + 1, k + 1))) for i in range(3)]
        got = list(glass_b200.iternorm(rows))
        assert all(isinstance(g, np.ndarray) for g in got)
        assert np.abs(np.stack(got) - gold[f"iternorm_k{k}"]).max() <= 1e-14
    # the factor rebuilt from the weights reproduces the covariance (full band)
    wts = np.stack(list(glass_b200.iternorm([np.pad(cov[i, i::-1], (0, 2 - i)) for i in range(3)])))
    lmat = np.zeros((3, 3))
    for i in range(3):
        lmat[i, i] = wts[i, 2]
        for j in range(1, i + 1):
            lmat[i, i - j] = wts[i, 2 - j]
    assert np.allclose(lmat @ lmat.T, cov, rtol=0, atol=1e-15)
    # leading dimensions and device rows
    rng = np.random.default_rng(3)
    rows = [np.concatenate([2.0 + rng.random((3, 4, 1)), 0.3 * rng.random((3, 4, 2))], axis=-1) for _ in range(5)]
    want = G.iternorm(rows)
    got = list(glass_b200.iternorm(torch.as_tensor(r).to(cuda_device) for r in rows))
    assert all(g.is_cuda and g.shape == (3, 4, 3) for g in got)
    assert np.abs(np.stack([g.cpu().numpy() for g in got]) - np.stack(want)).max() <= 1e-13
    with pytest.raises(ValueError, match="empty covariance"):
        list(glass_b200.iternorm([np.ones(0)]))
    with pytest.raises(ValueError, match="shape mismatch"):
        list(glass_b200.iternorm([np.ones(1), np.ones((5, 2))]))
    with pytest.raises(ValueError, match="shape mismatch"):
        list(glass_b200.iternorm([np.ones((1, 1)), np.ones((1, 2))]))
    with pytest.raises(ValueError, match="not positive definite"):
        list(glass_b200.iternorm([np.array([1.0, 0.0]), np.array([0.1, 1.0])]))


def test_generate_not_positive_definite_is_deferred(cuda_device):
    """The reference's error when a covariance is not positive definite: shells before the
    failing one are still yielded (K1 records a flag per shell on its own stream)."""
    import glass_b200

    nshell, lmax, nside, ncorr = 12, 48, 16, 10
    gls = synthetic_gls(nshell, lmax, ncorr)
    flds = [glass_b200.grf.Lognormal()] * nshell
    assert len(list(glass_b200.generate(flds, gls, nside, ncorr=ncorr, rng=5))) == nshell
    # shell 2 is more strongly correlated with shell 1 than a positive definite matrix allows
    bad = [g.copy() for g in gls]
    bad[2 * 3 // 2 + 1] = 3.0 * gls[0]
    g = glass_b200.generate(flds, bad, nside, ncorr=ncorr, rng=5)
    got = []
    with pytest.raises(ValueError, match="covariance matrix is not positive definite"):
        for m in g:
            got.append(m)
    assert len(got) == 2


def test_philox_normals_statistics(cuda_device):
    """Random draws are validated statistically (north_star): recovered C_l within
    cosmic variance, seed reproducibility, shell independence."""
    import glass_b200

    nside, lmax = 64, 100
    l = np.arange(lmax + 1)
    cl = 1.0 / (l + 10.0) ** 2
    maps = [next(glass_b200.generate([glass_b200.grf.Normal()], [cl], nside, rng=s)) for s in (1, 1, 2)]
    assert np.array_equal(maps[0], maps[1])
    assert not np.array_equal(maps[0], maps[2])
    # alm-level check: draw z directly and test moments per l
    from glass_b200 import _lib
    import ctypes as C

    lib = _lib.load()
    n = (lmax + 1) * (lmax + 2) // 2
    z = torch.empty(n, dtype=torch.complex128, device=cuda_device)
    _lib.check(lib.glb_alm_draw(lmax, C.c_uint64(5), C.c_uint32(0), z.data_ptr(), torch.cuda.current_stream().cuda_stream))
    zz = z.cpu().numpy()
    assert abs(zz.real.mean()) < 5 / np.sqrt(n) and abs(zz.imag.mean()) < 5 / np.sqrt(n)
    assert abs(zz.real.var() - 1) < 5 * np.sqrt(2 / n) and abs(zz.imag.var() - 1) < 5 * np.sqrt(2 / n)
    from scipy import stats

    assert stats.kstest(zz.real, "norm").pvalue > 1e-4
    assert stats.kstest(zz.imag, "norm").pvalue > 1e-4
    assert abs(np.corrcoef(zz.real, zz.imag)[0, 1]) < 5 / np.sqrt(n)


def test_generate_shell_sharding(cuda_device):
    """rank r of W produces shells r::W; the union equals the unsharded run exactly."""
    import glass_b200

    nside, lmax, nshell, ncorr = 16, 32, 7, 2
    gls = synthetic_gls(nshell, lmax, ncorr)
    fields = [glass_b200.grf.Lognormal()] * nshell
    full = list(glass_b200.generate(fields, gls, nside, ncorr=ncorr, rng=11))
    for world in (2, 3):
        for rank in range(world):
            mine = list(glass_b200.generate(fields, gls, nside, ncorr=ncorr, rng=11, shells=range(rank, nshell, world)))
            assert len(mine) == len(range(rank, nshell, world))
            for k, j in enumerate(range(rank, nshell, world)):
                assert np.array_equal(mine[k], full[j])


def test_recovered_cl_within_cosmic_variance(cuda_device):
    """north_star: random draws validated statistically -- the C_l recovered from generated
    Gaussian maps (Philox normals -> synthesis -> analysis) agree with the input spectrum
    within cosmic variance; cross-shell correlation follows the input cross-spectrum."""
    import glass_b200
    from glass_b200 import healpix as hp

    nside, lmax = 128, 160
    l = np.arange(lmax + 1)
    cl = 1e-2 * (l + 1.0) ** -1.5
    cl[0] = 0.0
    gls = [cl, cl, 0.5 * cl]  # two fields, cross-spectrum 0.5 C_l
    maps = list(glass_b200.generate([glass_b200.grf.Normal()] * 2, [torch.as_tensor(g).to(cuda_device) for g in gls], nside, rng=2024))
    alms = [hp.map2alm(m, lmax=lmax, pol=False, niter=3).cpu().numpy() for m in maps]

    def spec(a, b):
        out = np.zeros(lmax + 1)
        for mm in range(lmax + 1):
            i0 = H.alm_index(lmax, mm, mm)
            seg_a, seg_b = a[i0 : i0 + lmax + 1 - mm], b[i0 : i0 + lmax + 1 - mm]
            out[mm:] += (1 if mm == 0 else 2) * (seg_a * np.conj(seg_b)).real
        return out / (2 * l + 1)

    for a in alms:
        c = spec(a, a)
        chi = (c[2:] / cl[2:] - 1) / np.sqrt(2.0 / (2 * l[2:] + 1))
        assert abs(chi.mean()) < 4 / np.sqrt(chi.size) and 0.8 < chi.std() < 1.2, (chi.mean(), chi.std())
    cx = spec(alms[0], alms[1])
    # var of the cross estimate: (C11 C22 + C12^2)/(2l+1) = 1.25 C^2/(2l+1)
    chi = (cx[2:] / cl[2:] - 0.5) / np.sqrt(1.25 / (2 * l[2:] + 1))
    assert abs(chi.mean()) < 4 / np.sqrt(chi.size) and 0.8 < chi.std() < 1.2, (chi.mean(), chi.std())
