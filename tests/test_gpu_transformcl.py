"""GPU tests of the spectra side: C_l <-> C(theta) as FP64 DGEMMs and the
batched Gauss-Newton solver.  Same properties as the CPU suite checks on CPU tensors
(tests/test_cpu_host.py), here on the device the product uses.  Named to sort last: these
were written after the round's GPU minutes were spent and have not run on a GPU yet."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_transform_pair_on_device():
    import torch
    from scipy.special import eval_legendre

    from glass_b200 import transformcl as tcl

    for n in (3, 101, 1024):
        rng = np.random.default_rng(n)
        cl = rng.standard_normal(n) / (1 + np.arange(n)) ** 2
        c = tcl.cltocorr(cl)
        assert isinstance(c, np.ndarray)
        if n <= 101:
            x = np.cos(tcl.theta(n))
            direct = sum((2 * l + 1) / (4 * np.pi) * cl[l] * eval_legendre(l, x) for l in range(n))
            assert np.abs(c - direct).max() <= 1e-12 * np.abs(direct).max()
        assert np.abs(tcl.corrtocl(c) - cl).max() <= 1e-12 * np.abs(cl).max()
    d = torch.as_tensor(cl, device="cuda")
    back = tcl.corrtocl(tcl.cltocorr(d))
    assert back.is_cuda and torch.allclose(back, d, rtol=0, atol=1e-12 * float(d.abs().max()))
    tcl.clear_tables()


def test_solver_on_device():
    import torch

    import glass_b200 as glass
    from glass_b200 import grf

    lmax = 100
    cl = 1e-2 / (2 * np.arange(lmax + 1) + 1) ** 2
    t = grf.Lognormal(0.8)
    gl, cl_out, info = grf.solve(cl, t, pad=2 * cl.shape[0], cltol=1e-7)
    assert info > 0 and cl_out.shape[0] == 3 * cl.shape[0]
    assert np.allclose(cl_out[1 : cl.shape[0]], cl[1:], atol=0.0, rtol=1e-7)
    assert np.allclose(grf.solve(cl, t, maxiter=0)[0], grf.compute(cl, t), rtol=1e-12, atol=1e-18)
    fields = [grf.Lognormal(1.0), grf.Lognormal(0.7), grf.Normal()]
    spectra = [torch.as_tensor(0.5 ** (i - j) * cl * (1 + 0.1 * i), device="cuda") for i in range(3) for j in range(i, -1, -1)]
    gls = glass.solve_gaussian_spectra(fields, spectra)
    assert all(g.is_cuda and g.shape[0] == cl.shape[0] for g in gls)
    for k, (i, j, s) in enumerate(glass.enumerate_spectra(spectra)):
        g, _, info = grf.solve(s, fields[i], fields[j], pad=2 * s.shape[0])
        assert info > 0
        # same iteration path -> agreement at rounding level (1e-15 on CPU tensors); the bound leaves room for
        # a column stopping one Gauss-Newton step apart from its single run (cltol = 1e-5) on other GEMM kernels
        assert float((g - gls[k]).abs().max()) <= 2e-5 * float(g.abs().max())
    glass.transformcl.clear_tables()


def test_regularized_spectra_on_device():
    import os

    import torch

    import glass_b200 as glass
    from glass_b200 import algorithm

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "glass_reference_spectra.npz"))
    close = lambda a, b: np.allclose(a, b, rtol=1e-8, atol=1e-10 * np.abs(b).max())  # noqa: E731
    assert close(algorithm.cov_clip(g["reg_cov"]), g["reg_clip"])
    assert close(algorithm.cov_nearest(g["reg_cov"]), g["reg_nearest"])
    bad = np.split(g["reg_gls"], np.cumsum(g["reg_gls_len"])[:-1])
    assert close(np.stack(glass.regularized_spectra(bad, method="clip")), g["reg_spectra_clip"])
    reg = glass.regularized_spectra([torch.as_tensor(b, device="cuda") for b in bad], method="nearest")
    assert all(r.is_cuda for r in reg)
    assert close(torch.stack(reg).cpu().numpy(), g["reg_spectra_nearest"])


def test_points_cuts_on_device():
    """glb_points_cuts (one thread walking csrc/points_cuts.cuh) against the oracle's restatement of
    the reference's 1000-pixel stepping loop (glass/points.py:409-437), in chunks of 5 cuts."""
    import torch

    from glass_b200 import _lib
    from glass_b200.points import _Population
    from oracle import glass_ref as G

    rng = np.random.default_rng(9)
    dev = torch.device("cuda", 0)
    for npix, dens, batch in [(3072, 0.01, 2), (3072, 2.0, 1000), (12 * 64**2, 0.08, 157), (5000, 0.5, 1)]:
        counts = rng.poisson(dens, npix)
        counts[npix // 3] += 40  # an oversize pixel
        pop = object.__new__(_Population)
        pop.npix, pop.lib, pop.device = npix, _lib.load(), dev
        pop.off = torch.as_tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64), device=dev)
        pop.total = int(counts.sum())
        assert list(pop.cuts(batch, chunk=5)) == [tuple(int(v) for v in r) for r in G.batch_cuts(counts, batch)]
