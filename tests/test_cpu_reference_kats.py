"""Known-answer tests of the reference's own suite that need none of its third-party libraries,
ported one to one for the host-side mirrors (tests/core/test_fields.py:486-524, 637-840,
tests/core/test_algorithm.py:56-160, tests/core/test_points.py:26-48, 440-468 of the reference).
The expected values are the reference's."""
import types

import numpy as np
import pytest
import torch

import glass_b200 as glass
from glass_b200 import algorithm


@pytest.fixture(autouse=True)
def _device_agnostic_library_calls_on_cpu(monkeypatch):
    # cov_clip / nearcorr are batched torch.linalg calls; the product runs them on CUDA only
    import glass_b200.transformcl as tcl

    monkeypatch.setattr(tcl, "_compute_device", lambda *a: (torch.device("cpu"), False))


def test_enumerate_spectra():
    n = 100
    spectra = [np.asarray(x) for x in range(n * (n + 1) // 2)]
    indices = [(i, j) for i in range(n) for j in range(i, -1, -1)]
    it = glass.enumerate_spectra(spectra)
    for k, (i, j) in enumerate(indices):
        assert next(it) == (i, j, k)
    with pytest.raises(StopIteration):
        next(it)


def test_spectra_indices():
    assert np.array_equal(glass.spectra_indices(0), np.zeros((0, 2), dtype=np.int64))
    assert glass.spectra_indices(0).dtype == np.int64
    assert np.array_equal(glass.spectra_indices(1), [[0, 0]])
    assert np.array_equal(glass.spectra_indices(2), [[0, 0], [1, 1], [1, 0]])
    assert np.array_equal(glass.spectra_indices(3), [[0, 0], [1, 1], [1, 0], [2, 2], [2, 1], [2, 0]])
    assert torch.equal(glass.spectra_indices(2, xp=torch), torch.tensor([[0, 0], [1, 1], [1, 0]]))


def test_spectra_orders():
    assert glass.glass_to_healpix_spectra([11, 22, 21, 33, 32, 31, 44, 43, 42, 41]) == [11, 22, 33, 44, 21, 32, 43, 31, 42, 41]
    assert glass.healpix_to_glass_spectra([11, 22, 33, 44, 21, 32, 43, 31, 42, 41]) == [11, 22, 21, 33, 32, 31, 44, 43, 42, 41]


def test_triangle_numbers():
    for t in (2, 4, 5, 7, 8, 9, 11, 12, 13, 14, 16):
        with pytest.raises(ValueError, match=f"invalid number of spectra: {t}"):
            glass.nfields_from_nspectra(t)
    for n in range(10):
        assert glass.nfields_from_nspectra(n * (n + 1) // 2) == n


def test_lognormal_shift_hilbert2011():
    zs = [0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0]
    check = [0.0103031, 0.02975, 0.0538781, 0.0792, 0.103203, 0.12435, 0.142078, 0.1568]
    assert np.allclose([glass.lognormal_shift_hilbert2011(z) for z in zs], check, atol=1e-4, rtol=1e-4)


def test_cov_from_spectra():
    spectra = [np.asarray(x) for x in [[110, 111, 112, 113], [220, 221, 222, 223], [210, 211, 212, 213],
                                       [330, 331, 332, 333], [320, 321, 322, 323], [310, 311, 312, 313]]]
    full = np.array([[[110 + l, 210 + l, 310 + l], [210 + l, 220 + l, 320 + l], [310 + l, 320 + l, 330 + l]] for l in range(4)])
    assert np.array_equal(glass.cov_from_spectra(spectra), full)
    assert np.array_equal(glass.cov_from_spectra(spectra, lmax=1), full[:2])
    assert np.array_equal(glass.cov_from_spectra(spectra, lmax=4), np.concatenate([full, np.zeros((1, 3, 3))]))


def test_check_posdef_spectra():
    assert glass.check_posdef_spectra([np.asarray(x) for x in [[1.0, 1.0, 1.0], [1.0, 1.0, 1.0], [0.9, 0.9, 0.9]]])
    assert glass.check_posdef_spectra([np.asarray(x) for x in [[1.0, 1.0, 1.0], [1.0, 1.0, 0.0], [0.9, 1.0, 0.0]]])
    assert not glass.check_posdef_spectra([np.asarray(x) for x in [[1.0, 1.0, 1.0], [1.0, 1.0, 1.0], [1.1, 1.1, 1.1]]])


def test_regularized_spectra_methods(monkeypatch):
    rng = np.random.default_rng(42)
    spectra = [rng.random(101) for _ in range(6)]
    calls = []
    for name in ("cov_nearest", "cov_clip"):
        real = getattr(algorithm, name)
        monkeypatch.setattr(algorithm, name, lambda *a, _r=real, _n=name, **k: (calls.append(_n), _r(*a, **k))[1])
    with pytest.warns(UserWarning, match="Nearest correlation matrix not found"):
        glass.regularized_spectra(spectra, method="nearest")
    glass.regularized_spectra(spectra, method="clip")
    assert calls == ["cov_nearest", "cov_clip"]
    with pytest.raises(ValueError, match="unknown method"):
        glass.regularized_spectra(spectra, method="unknown")


def test_cov_clip():
    m = np.random.default_rng(42).random((4, 4))
    a = (m + m.T) / 2
    assert np.all(np.linalg.eigvalsh(algorithm.cov_clip(a)) >= 0)
    cov = algorithm.cov_clip(a, rtol=1.0)
    assert np.allclose(np.linalg.eigvalsh(cov), np.max(np.linalg.eigvalsh(a)))


def test_nearcorr():
    a = np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 1.0], [0.0, 1.0, 1.0]])  # Higham (2002)
    b = np.array([[1.0000, 0.7607, 0.1573], [0.7607, 1.0000, 0.7607], [0.1573, 0.7607, 1.0000]])
    assert np.allclose(algorithm.nearcorr(a), b, atol=1e-4)
    assert np.allclose(algorithm.nearcorr(a, tol=1e-10), b, atol=1e-4)
    with pytest.warns(UserWarning, match="Nearest correlation matrix not found in 0 iterations"):
        x = algorithm.nearcorr(a, niter=0)
    assert np.array_equal(x, a)
    with pytest.raises(ValueError, match="non-square matrix"):
        algorithm.nearcorr(np.zeros((4, 3)))
    with pytest.warns(UserWarning, match="Nearest correlation matrix not found in 1 iterations"):
        algorithm.nearcorr(a, niter=1)


def test_cov_nearest():
    m = np.random.default_rng(42).random((4, 4))
    a = np.eye(4) + (m + m.T) / 2
    assert np.all(np.linalg.eigvalsh(algorithm.cov_nearest(a)) >= -1e-15)
    with pytest.raises(ValueError, match="negative values"):
        algorithm.cov_nearest(np.array([[1, 0, 0], [0, 1, 0], [0, 0, -1]]))


def test_effective_bias():
    w = types.SimpleNamespace(za=np.linspace(0, 2, 100), wa=np.full(100, 2.0))
    assert glass.effective_bias(np.linspace(0, 1, 10), np.zeros(10), w) == 0.0
    assert glass.effective_bias(np.zeros(10), np.full(10, 0.5), w) == 0.0
    assert glass.effective_bias(np.linspace(0, 1, 10), np.full(10, 0.5), w) == 0.25


def test_position_weights():
    rng = np.random.default_rng(42)
    for bshape in None, (), (100,), (100, 1):
        for cshape in (100,), (100, 50), (100, 3, 2):
            counts = rng.random(cshape)
            bias = None if bshape is None else rng.random(bshape)
            weights = glass.position_weights(counts, bias)
            expected = counts / np.sum(counts, axis=0, keepdims=True)
            if bias is not None:
                if bias.ndim > expected.ndim:
                    expected = np.expand_dims(expected, axis=tuple(range(expected.ndim, bias.ndim)))
                else:
                    bias = np.expand_dims(bias, axis=tuple(range(bias.ndim, expected.ndim)))
                expected = bias * expected
            assert np.array_equal(weights, expected)


# ---- tests/core/grf/test_core.py, test_transformations.py of the reference -------------------


def _nulp(a, b):
    return np.max(np.abs(a - b) / np.spacing(np.maximum(np.abs(a), np.abs(b))))


def test_grf_corr_unknown():
    from glass_b200 import grf

    class Unknown:
        def corr(self, _other, _x):
            return NotImplemented

        icorr = dcorr = corr

    x = np.zeros(10)
    for fn in (grf.corr, grf.icorr, grf.dcorr):
        with pytest.raises(NotImplementedError, match="Unknown"):
            fn(grf.Normal(), Unknown(), x)


def test_grf_transformations_and_pairs():
    from glass_b200 import grf

    rng = np.random.default_rng(42)
    x = rng.standard_normal(10)
    assert np.array_equal(grf.Normal()(x, 1.0), x)
    for lam in 1.0, rng.uniform():
        var = rng.uniform()
        assert np.array_equal(grf.Lognormal(lam)(x, var), lam * np.expm1(x - var / 2))
        a = np.sqrt(1 - var)
        assert np.array_equal(grf.SquaredNormal(a, lam)(x, var), lam * ((x - a) ** 2 - 1))
    x = rng.random(10)
    n = grf.Normal()
    assert np.array_equal(grf.corr(n, n, x), x) and np.array_equal(grf.icorr(n, n, x), x)
    assert np.array_equal(grf.dcorr(n, n, x), np.ones_like(x))
    lam1, lam2 = rng.uniform(size=2)
    t1, t2 = grf.Lognormal(lam1), grf.Lognormal(lam2)
    y = lam1 * lam2 * np.expm1(x)
    assert np.array_equal(grf.corr(t1, t2, x), y) and _nulp(grf.icorr(t1, t2, y), x) <= 1
    assert np.array_equal(grf.dcorr(t1, t2, x), lam1 * lam2 * np.exp(x))
    y = lam1 * x
    assert np.array_equal(grf.corr(t1, n, x), y) and _nulp(grf.icorr(t1, n, y), x) <= 1
    assert np.array_equal(grf.dcorr(t1, n, x), lam1 * np.ones_like(x))
    assert np.array_equal(grf.corr(n, t1, x), y)  # the reflected pair dispatches to the Lognormal
    (lam1, var1), (lam2, var2) = rng.uniform(size=2), rng.uniform(size=2)
    a1, a2 = np.sqrt(1 - var1), np.sqrt(1 - var2)
    s1, s2 = grf.SquaredNormal(a1, lam1), grf.SquaredNormal(a2, lam2)
    y = 2 * lam1 * lam2 * x * (x + 2 * a1 * a2)  # arXiv:2408.16903 (E.7)
    assert np.array_equal(grf.corr(s1, s2, x), y) and _nulp(grf.icorr(s1, s2, y), x) <= 8
    assert np.array_equal(grf.dcorr(s1, s2, x), 4 * lam1 * lam2 * (x + a1 * a2))
