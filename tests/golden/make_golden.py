#!/usr/bin/env python
"""
Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE (/root/reference/glass)
in the build container.  Nothing of the reference is copied into this repo: the package is
imported from where it lies, on top of small shim modules standing in for its third-party
dependencies that are not installed here (array_api_compat, array_api_extra, transformcl,
healpy, healpix).  The shims for healpy/healpix only *record* what the GLASS code passes to
the seam (alm, ipix, ...) or delegate to the oracle; the vectors therefore pin the GLASS-side
NumPy code, not healpy.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

The GPU box has no /root/reference: tests only read the committed .npz files.
"""
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import healpix_ref as H  # noqa: E402

REF = "/root/reference"


def install_shims():
    # ---- array_api_compat -------------------------------------------------
    aac = types.ModuleType("array_api_compat")

    def array_namespace(*xs, use_compat=None, api_version=None):
        return np

    aac.array_namespace = array_namespace
    aac.get_namespace = array_namespace
    aac.is_jax_array = lambda x: False
    aac.is_numpy_array = lambda x: isinstance(x, np.ndarray)
    aac.is_array_api_obj = lambda x: isinstance(x, np.ndarray)
    aac.is_numpy_namespace = lambda xp: xp is np
    aac.is_jax_namespace = lambda xp: False
    aac.is_array_api_strict_namespace = lambda xp: False
    aac.size = lambda x: x.size
    aac.device = lambda x: "cpu"
    sys.modules["array_api_compat"] = aac
    # numpy lacks a few array-API spellings used by the reference
    if not hasattr(np, "concat"):
        np.concat = np.concatenate

    # ---- array_api_extra ---------------------------------------------------
    xpx = types.ModuleType("array_api_extra")

    class _At:
        def __init__(self, x, idx=None):
            self.x, self.idx = x, idx

        def __getitem__(self, idx):
            return _At(self.x, idx)

        def set(self, v, copy=None):
            self.x[self.idx] = v
            return self.x

        def add(self, v, copy=None):
            self.x[self.idx] += v
            return self.x

        def multiply(self, v, copy=None):
            self.x[self.idx] *= v
            return self.x

    xpx.at = lambda x, idx=None: _At(x, idx)
    xpx.pad = lambda x, pad_width, mode="constant", constant_values=0, xp=None: np.pad(x, pad_width, constant_values=constant_values)

    def apply_where(cond, args, f1, f2=None, /, *, fill_value=None, xp=None):
        args = args if isinstance(args, tuple) else (args,)
        out = np.full(np.shape(cond), fill_value if fill_value is not None else 0.0, dtype=np.result_type(*args, float))
        with np.errstate(all="ignore"):
            out[cond] = f1(*[a[cond] for a in args])
            if f2 is not None:
                out[~cond] = f2(*[a[~cond] for a in args])
        return out

    xpx.apply_where = apply_where
    xpx.union1d = lambda a, b, /, *, xp=None: np.union1d(a, b)
    sys.modules["array_api_extra"] = xpx

    # ---- transformcl --------------------------------------------------------
    tcl = types.ModuleType("transformcl")

    def cltovar(cl):
        cl = np.asarray(cl)
        ell = np.arange(cl.shape[0])
        return np.sum((2 * ell + 1) / (4 * np.pi) * cl)

    tcl.cltovar = cltovar
    tcl.cltocorr = tcl.corrtocl = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    sys.modules["transformcl"] = tcl

    # ---- healpy / healpix: record + oracle ----------------------------------
    hpy = types.ModuleType("healpy")
    REC = {"alm": [], "ipix": []}
    hpy._rec = REC

    def alm2map(alms, nside, inplace=False, lmax=None, pixwin=False, pol=True, **kw):
        REC["alm"].append(np.array(alms, copy=True))
        return np.array(alms, copy=True)  # the GLASS code just yields it on

    hpy.alm2map = alm2map
    hpy.npix2nside = H.npix2nside
    hpy.nside2npix = H.nside2npix
    hpy.Rotator = object
    sys.modules["healpy"] = hpy

    hpx = types.ModuleType("healpix")
    hpx.npix2nside = H.npix2nside
    hpx.nside2npix = H.nside2npix

    def randang(nside, ipix, lonlat=False, rng=None):
        REC["ipix"].append(np.array(ipix, copy=True))
        n = np.asarray(ipix).size
        return H.ring2ang_uv(nside, ipix, np.full(n, 0.5), np.full(n, 0.5), lonlat=lonlat)

    hpx.randang = randang
    hpx.ang2pix = lambda nside, a, b, nest=False, lonlat=False: H.ang2pix(nside, a, b, lonlat=lonlat)
    sys.modules["healpix"] = hpx
    return REC


def synthetic_gls(nshell, lmax, ncorr, ragged=False):
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


class MockCosmology:  # tests/fixtures/domain.py:36-97
    Omega_m0 = 0.3
    hubble_distance = 4.4e3

    def H_over_H0(self, z):  # noqa: N802
        return (self.Omega_m0 * (1 + z) ** 3 + 1 - self.Omega_m0) ** 0.5

    def transverse_comoving_distance(self, z, z2=None):
        if z2 is None:
            return self.hubble_distance * np.asarray(z) * 1_000
        return self.hubble_distance * (np.asarray(z2) - np.asarray(z)) * 1_000


def main():
    REC = install_shims()
    sys.path.insert(0, REF)
    import glass  # the reference itself
    import glass.fields
    import glass.points

    out = {}

    # ---- fields: iternorm / cls2cov / alm of _generate_grf with seed-42 NumPy normals ----
    cov = np.array([[1.0, 0.2, 0.1], [0.2, 0.5, 0.2], [0.1, 0.2, 0.3]])
    for k in (0, 1, 2):
        rows = [cov[i, i::-1][: min(i, k) + 1] for i in range(3)]
        # iternorm needs a fixed row width
        rows = [np.pad(r, (0, k + 1 - r.size)) for r in rows]
        out[f"iternorm_k{k}"] = np.stack(list(glass.iternorm(rows)))
    for name, (nshell, lmax, ncorr, ragged) in {"a": (4, 12, 2, False), "b": (5, 9, None, False), "c": (4, 10, 1, True)}.items():
        nc = nshell - 1 if ncorr is None else ncorr
        gls = synthetic_gls(nshell, lmax, nc, ragged)
        out[f"cls2cov_{name}"] = np.stack([c.copy() for c in glass.cls2cov(gls, lmax + 1, nshell, nc)])
        out[f"iternorm_{name}"] = np.stack(list(glass.iternorm(glass.cls2cov(gls, lmax + 1, nshell, nc))))
        REC["alm"].clear()
        maps = list(glass.fields._generate_grf(gls, 4, ncorr=ncorr, rng=np.random.default_rng(42)))
        out[f"grf_alm_{name}"] = np.stack(REC["alm"])
        # lognormal transform through generate (alm passthrough "map" -> use real part)
    x = np.linspace(-2, 2, 41)
    out["lognormal_x"] = x
    out["lognormal_y"] = glass.grf.Lognormal(0.7)(x.copy(), 0.35)
    out["lognormal_y1"] = glass.grf.Lognormal()(x.copy(), 0.35)
    out["sqnormal_y"] = glass.grf.SquaredNormal(0.3, 1.5)(x.copy(), 0.35)
    alm = np.arange(10) + 1j * np.arange(10)[::-1]
    out["g2h_in"] = alm
    out["g2h_out"] = glass.fields._glass_to_healpix_alm(alm)
    out["multalm_out"] = glass.harmonics.multalm(alm[:6], np.array([2.0, 0.5, 1.0]))

    # ---- points: expected counts and batch cuts ----
    rng = np.random.default_rng(3)
    nside = 8
    npix = 12 * nside**2
    delta = np.expm1(0.5 * rng.standard_normal(npix) - 0.125)
    vis = rng.random(npix)
    out["pt_delta"], out["pt_vis"] = delta, vis
    ngal = np.asarray(1e-3)
    for tag, (bias, v, model, rm) in {
        "none": (None, None, glass.linear_bias, False),
        "lin_vis": (np.asarray(0.8), vis, glass.linear_bias, False),
        "loglin": (np.asarray(1.3), None, glass.loglinear_bias, False),
        "lin_vis_rm": (np.asarray(0.8), vis, glass.linear_bias, True),
    }.items():
        n = glass.points._compute_density_contrast(bias, model, delta, ())
        n = glass.points._compute_expected_count((), n, ngal, remove_monopole=rm)
        n = glass.points._apply_visibility((), n, v)
        out[f"pt_nbar_{tag}"] = n
    counts = rng.poisson(3.0 * (1 + np.clip(delta, -1, 3)))
    counts[10:40] = 0
    counts[100] = 60
    out["pt_counts"] = counts
    for batch in (1_000_000, 500, 37, 1):
        REC["ipix"].clear()
        sizes = [int(c) for _lo, _la, c in glass.points._sample_galaxies_per_pixel(batch, (), (), counts)]
        out[f"pt_batches_{batch}"] = np.array(sizes)
        out[f"pt_ipix_{batch}"] = np.concatenate(REC["ipix"]) if REC["ipix"] else np.zeros(0, dtype=np.int64)
    out["ARCMIN2_SPHERE"] = np.asarray(glass.points.ARCMIN2_SPHERE)

    # ---- lensing: multi-plane convergence with MockCosmology ----
    cosmo = MockCosmology()
    shells = [glass.RadialWindow(np.array([i, i + 1.0, i + 2.0]), np.array([0.0, 1.0, 0.0]), i + 1.0) for i in range(5)]
    deltas = np.random.default_rng(42).random((5, 48))
    mpc = glass.MultiPlaneConvergence(cosmo)
    kap = []
    for i, w in enumerate(shells):
        mpc.add_window(deltas[i].copy(), w)
        kap.append(np.array(mpc.kappa, copy=True))
    out["mpc_deltas"], out["mpc_kappas"] = deltas, np.stack(kap)
    out["mpc_matrix"] = glass.multi_plane_matrix(shells, cosmo)

    # ---- galaxies / shapes ----
    nside = 4
    npix = 12 * nside**2
    r = np.random.default_rng(1)
    kappa, g1, g2 = 0.1 * r.standard_normal((3, npix))
    n = 300
    ipix = r.integers(0, npix, n)
    lon, lat = H.ring2ang_uv(nside, ipix, r.random(n), r.random(n), lonlat=True)
    eps = 0.3 * (r.standard_normal(n) + 1j * r.standard_normal(n))
    out["gs_kappa"], out["gs_g1"], out["gs_g2"], out["gs_lon"], out["gs_lat"], out["gs_eps"] = kappa, g1, g2, lon, lat, eps
    out["gs_reduced"] = glass.galaxy_shear(lon, lat, eps, kappa, g1, g2, reduced_shear=True)
    out["gs_plain"] = glass.galaxy_shear(lon, lat, eps, kappa, g1, g2, reduced_shear=False)
    rr = np.random.default_rng(7)
    out["eps_intnorm"] = glass.ellipticity_intnorm(500, 0.256, rng=rr, xp=np)
    rr = np.random.default_rng(7)
    out["eps_normals"] = rr.standard_normal(500) + 1j * rr.standard_normal(500)
    z = np.linspace(0.0, 2.0, 101)
    nz = z**2 * np.exp(-((z / 0.5) ** 1.5))
    rr = np.random.default_rng(9)
    out["z_grid"], out["z_nz"] = z, nz
    out["z_samples"] = glass.redshifts_from_nz(400, z, nz, rng=rr, warn=False)
    out["z_uniform"] = np.random.default_rng(9).uniform(0.0, 1.0, size=400)

    # ---- discretized_cls / effective_cls (glass/fields.py:239-300, 607-694) ----
    gls_d = synthetic_gls(4, 12, 3)
    PW = 1.0 / (1.0 + 0.01 * np.arange(20.0) ** 2)
    import glass.healpix as _ghp

    _ghp.pixwin = lambda nside, lmax=None, pol=False, xp=None: PW[: (lmax + 1) if lmax is not None else None]
    out["dcl_pw"] = PW
    for tag, kw in {"lmax": {"lmax": 8}, "ncorr": {"ncorr": 1}, "all": {"lmax": 9, "ncorr": 2, "nside": 4}}.items():
        res = glass.discretized_cls(gls_d, **kw)
        out[f"dcl_{tag}_len"] = np.array([r.shape[0] for r in res])
        out[f"dcl_{tag}"] = np.concatenate(res)
    w1 = np.random.default_rng(4).random((4, 3))
    w2 = np.random.default_rng(5).random((4, 2))
    out["ecl_w1"], out["ecl_w2"] = w1, w2
    out["ecl_auto"] = glass.effective_cls(gls_d, w1)
    out["ecl_cross"] = glass.effective_cls(gls_d, w1, w2, lmax=7)

    # ---- gaussian_phz (glass/galaxies.py:350-455): array z, bounds that force several rounds ----
    zt = np.random.default_rng(21).uniform(0.0, 1.5, size=300)
    out["phz_z"] = zt
    out["phz_out"] = glass.gaussian_phz(zt, 0.2, lower=0.1, upper=1.2, rng=np.random.default_rng(22), xp=np)
    rr = np.random.default_rng(22)
    out["phz_normals"] = np.stack([rr.standard_normal(300) for _ in range(40)])

    # ---- uniform_positions (glass/points.py:543-607): two populations, seed 11 ----
    ngal_u = np.array([2e-6, 5e-6])
    ups = list(glass.uniform_positions(ngal_u, rng=np.random.default_rng(11), xp=np))
    out["up_ngal"] = ngal_u
    out["up_lon"] = np.concatenate([u[0] for u in ups])
    out["up_lat"] = np.concatenate([u[1] for u in ups])
    out["up_count"] = np.stack([u[2] for u in ups])
    # the same stream replayed: Poisson totals first, then per population lon deviates, lat deviates
    rr = np.random.default_rng(11)
    tot = rr.poisson(glass.points.ARCMIN2_SPHERE * ngal_u)
    ul, ub = [], []
    for n in tot:
        ul.append(rr.random(int(n)))
        ub.append(rr.random(int(n)))
    out["up_totals"], out["up_u_lon"], out["up_u_lat"] = tot, np.concatenate(ul), np.concatenate(ub)

    np.savez_compressed(os.path.join(HERE, "glass_reference_vectors.npz"), **out)
    print("wrote", len(out), "arrays,", sum(v.nbytes for v in out.values()) // 1024, "KiB")


def main_displace():
    """Second file (kept separate so that adding it did not rewrite the first):
    displace / displacement / deflect of glass/points.py:654-772 and glass/lensing.py:687-778."""
    install_shims()
    sys.path.insert(0, REF)
    import glass  # the reference itself
    import glass.lensing
    import glass.points

    rr = np.random.default_rng(77)
    n = 2000
    lon = rr.uniform(-180.0, 360.0, n)
    lat = np.degrees(np.arcsin(rr.uniform(-1.0, 1.0, n)))
    alpha = (rr.standard_normal(n) + 1j * rr.standard_normal(n)) * 10.0 ** rr.uniform(-6, -0.3, n)
    # edge cases: zero displacement, poles, due north / south / east
    lon[:6] = [0.0, 10.0, 20.0, 30.0, 40.0, 50.0]
    lat[:6] = [0.0, 90.0, -90.0, 45.0, -30.0, 89.999]
    alpha[:6] = [0.0, 0.1 + 0.0j, 0.2j, -0.3 + 0.0j, 0.0 + 0.25j, 0.01 + 0.01j]
    out = {"lon": lon, "lat": lat, "alpha": alpha}
    dlon, dlat = glass.points.displace(lon, lat, alpha)
    out["displace_lon"], out["displace_lat"] = dlon, dlat
    dlon2, dlat2 = glass.points.displace(lon, lat, np.stack([alpha.real, alpha.imag]))
    assert np.array_equal(dlon, dlon2) and np.array_equal(dlat, dlat2)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        flon, flat = glass.lensing.deflect(lon, lat, alpha)
    out["deflect_lon"], out["deflect_lat"] = flon, flat
    to_lon = rr.uniform(-180.0, 360.0, n)
    to_lat = np.degrees(np.arcsin(rr.uniform(-1.0, 1.0, n)))
    out["to_lon"], out["to_lat"] = to_lon, to_lat
    out["displacement"] = glass.points.displacement(lon, lat, to_lon, to_lat)
    np.savez_compressed(os.path.join(HERE, "glass_reference_displace.npz"), **out)
    print("wrote", len(out), "arrays to glass_reference_displace.npz")


def main_lensing_factors():
    """Fifth file: the l-dependent factors of glass/lensing.py:296-371, 403-428 as the reference's
    own source hands them to healpy.almxfl, and the alm it passes on to the spin transforms -- the
    transforms themselves are recording shims (healpy is absent), so this pins the GLASS-side
    arithmetic between them."""
    install_shims()
    sys.path.insert(0, REF)
    hpy = sys.modules["healpy"]
    rec = {"fl": [], "spin": [], "scalar": []}
    nside, lmax = 4, 9
    nalm = (lmax + 1) * (lmax + 2) // 2
    r = np.random.default_rng(17)
    alm0 = r.standard_normal(nalm) + 1j * r.standard_normal(nalm)
    lof = np.concatenate([np.arange(m, lmax + 1) for m in range(lmax + 1)])  # l of every m-major entry
    PW0 = 1.0 / (1.0 + 0.010 * np.arange(lmax + 1.0) ** 2)
    PW2 = 1.0 / (1.0 + 0.013 * np.arange(lmax + 1.0) ** 2)

    def map2alm(m, lmax=None, pol=True, use_pixel_weights=False, **kw):
        return alm0.copy()

    def almxfl(alm, fl, mmax=None, inplace=False):
        rec["fl"].append(np.array(fl, copy=True))
        out = alm if inplace else alm.copy()
        out *= np.asarray(fl)[lof]
        return out

    def alm2map(alm, nside, lmax=None, **kw):
        rec["scalar"].append(np.array(alm, copy=True))
        return np.zeros(12 * nside * nside)

    def alm2map_spin(alms, nside, spin, lmax, mmax=None):
        rec["spin"].append((spin, np.array(alms[0], copy=True), np.array(alms[1], copy=True)))
        return [np.zeros(12 * nside * nside), np.zeros(12 * nside * nside)]

    hpy.map2alm, hpy.almxfl, hpy.alm2map, hpy.alm2map_spin = map2alm, almxfl, alm2map, alm2map_spin
    hpy.get_nside = lambda m: H.npix2nside(np.asarray(m).shape[-1])
    hpy.pixwin = lambda nside, pol=False, lmax=None: (PW0[: lmax + 1], PW2[: lmax + 1]) if pol else PW0[: lmax + 1]
    import glass  # the reference itself
    import glass.healpix as ghp
    import glass.lensing

    ghp.pixwin = lambda nside, lmax=None, pol=False, xp=None: (PW0[: lmax + 1], PW2[: lmax + 1]) if pol else PW0[: lmax + 1]
    kappa = np.zeros(12 * nside * nside)
    out = {"alm0": alm0, "pw0": PW0, "pw2": PW2, "lmax": np.asarray(lmax)}
    for tag, disc in (("plain", False), ("disc", True)):
        for k in rec:
            rec[k].clear()
        glass.lensing.from_convergence(kappa, lmax, potential=True, deflection=True, shear=True, discretized=disc)
        out[f"fc_{tag}_fl"] = np.stack(rec["fl"])  # psi, alpha, gamma factors
        out[f"fc_{tag}_psi_alm"] = rec["scalar"][0]
        out[f"fc_{tag}_alpha_alm"] = rec["spin"][0][1]
        out[f"fc_{tag}_gamma_alm"] = rec["spin"][1][1]
        assert rec["spin"][0][0] == 1 and rec["spin"][1][0] == 2 and not rec["spin"][1][2].any()
        for k in rec:
            rec[k].clear()
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            glass.lensing.shear_from_convergence(kappa, lmax, discretized=disc)
        out[f"sfc_{tag}_fl"] = rec["fl"][0]
        out[f"sfc_{tag}_alm"] = rec["spin"][0][1]
    np.savez_compressed(os.path.join(HERE, "glass_reference_lensing_factors.npz"), **out)
    print("wrote", len(out), "arrays to glass_reference_lensing_factors.npz")


def main_positions():
    """Sixth file: the whole glass.positions_from_delta (glass/points.py:443-540) executed from the
    reference's source for populations with leading axes -- ngal (2,), delta (npix,), bias (3, 1),
    vis (npix,) -> dims (3, 2) -- with the Poisson count maps it drew recorded, so that the
    product can be replayed from the same counts; in-pixel positions are pixel centres
    (healpix.randang shim, u = v = 1/2)."""
    install_shims()
    sys.path.insert(0, REF)
    import glass  # the reference itself
    import glass.points

    nside = 4
    npix = 12 * nside**2
    r = np.random.default_rng(23)
    delta = np.expm1(0.4 * r.standard_normal(npix) - 0.08)
    vis = (r.random(npix) > 0.2) * r.random(npix)
    ngal = np.array([6e-4, 1.5e-3])
    bias = np.array([[0.5], [1.0], [1.7]])
    drawn = []
    real = glass.points._sample_number_galaxies

    def recording(n, *, rng=None):
        c = real(n, rng=rng)
        drawn.append(np.array(c, copy=True))
        return c

    glass.points._sample_number_galaxies = recording
    out = {"delta": delta, "vis": vis, "ngal": ngal, "bias": bias}
    for tag, kw in {"lin": {}, "loglin_rm": {"bias_model": glass.loglinear_bias, "remove_monopole": True}}.items():
        drawn.clear()
        res = list(glass.positions_from_delta(ngal, delta, bias, vis, batch=40, rng=np.random.default_rng(5), **kw))
        out[f"{tag}_counts"] = np.stack(drawn)  # one count map per population, in iteration order
        out[f"{tag}_batch_count"] = np.stack([c for _lo, _la, c in res])  # (nbatch, 3, 2) one-hot x size
        out[f"{tag}_lon"] = np.concatenate([lo for lo, _la, _c in res])
        out[f"{tag}_lat"] = np.concatenate([la for _lo, la, _c in res])
    # scalar inputs: count is a plain int
    drawn.clear()
    res = list(glass.positions_from_delta(2e-3, delta, None, None, batch=1000, rng=np.random.default_rng(6)))
    out["scalar_counts"] = drawn[0]
    out["scalar_batch_count"] = np.array([c for _lo, _la, c in res])
    assert all(isinstance(c, (int, np.integer)) for _lo, _la, c in res)
    out["scalar_lon"] = np.concatenate([lo for lo, _la, _c in res])
    np.savez_compressed(os.path.join(HERE, "glass_reference_positions.npz"), **out)
    print("wrote", len(out), "arrays to glass_reference_positions.npz")


def main_solver():
    """Fourth file: the reference's OWN spectra solver (glass/grf/_solver.py:27-148,
    glass/grf/_core.py:141-179, glass/fields.py:743-836) executed from source, with the oracle's
    restatement of the absent third-party transformcl as its transform pair."""
    install_shims()
    from oracle import transformcl_ref as tref

    tcl = sys.modules["transformcl"]
    tcl.cltocorr, tcl.corrtocl, tcl.cltovar = tref.cltocorr, tref.corrtocl, tref.cltovar
    sys.path.insert(0, REF)
    import glass  # the reference itself
    import glass.fields
    import glass.grf

    out = {}
    lmax = 40
    ell = np.arange(lmax + 1)
    cl = 1e-2 / (2 * ell + 1) ** 2
    out["cl"] = cl
    cases = {
        "ln": (glass.grf.Lognormal(0.8), None, {}),
        "ln_pad": (glass.grf.Lognormal(0.8), glass.grf.Lognormal(1.3), {"pad": 2 * (lmax + 1)}),
        "ln_mono": (glass.grf.Lognormal(), None, {"pad": 2 * (lmax + 1), "monopole": 0.0, "cltol": 1e-9, "gltol": 1e-9}),
        "ln_normal": (glass.grf.Lognormal(0.6), glass.grf.Normal(), {"pad": lmax + 1}),
        "sq": (glass.grf.SquaredNormal(0.9, 1.1), glass.grf.SquaredNormal(0.8, 0.7), {"pad": lmax + 1, "cltol": 1e-8}),
        "iter2": (glass.grf.Lognormal(0.5), None, {"pad": lmax + 1, "maxiter": 2, "cltol": 1e-14, "gltol": 1e-14}),
    }
    for tag, (t1, t2, kw) in cases.items():
        gl, rl, info = glass.grf.solve(cl.copy(), t1, t2, **kw)
        out[f"solve_{tag}_gl"], out[f"solve_{tag}_rl"], out[f"solve_{tag}_info"] = gl, rl, np.asarray(info)
    out["compute_ln"] = glass.grf.compute(cl.copy(), glass.grf.Lognormal(0.8))
    # solve_gaussian_spectra: 3 fields (two lognormal, one normal), a zero monopole, an empty spectrum
    fields = [glass.grf.Lognormal(1.0), glass.grf.Lognormal(0.7), glass.grf.Normal()]
    spectra = []
    for i in range(3):
        for j in range(i, -1, -1):
            c = 0.5 ** (i - j) * cl * (1 + 0.1 * i)
            if i == 1:
                c[0] = 0.0
            spectra.append(c if (i, j) != (2, 0) else np.zeros(0))
    gls = glass.fields.solve_gaussian_spectra(fields, spectra)
    out["sgs_len"] = np.array([g.shape[0] for g in gls])
    out["sgs_gls"] = np.concatenate(gls)
    out["sgs_spectra_len"] = np.array([c.shape[0] for c in spectra])
    out["sgs_spectra"] = np.concatenate(spectra)
    np.savez_compressed(os.path.join(HERE, "glass_reference_solver.npz"), **out)
    print("wrote", len(out), "arrays to glass_reference_solver.npz")


def main_spectra():
    """Third file: the spectra-order helpers of glass/fields.py:563-604, 897-1052 and
    position_weights of glass/points.py:610-651."""
    install_shims()
    sys.path.insert(0, REF)
    import glass  # the reference itself
    import glass.fields
    import glass.points

    # array_api_extra.tril_indices is reached through glass._array_api_utils.XPAdditions
    out = {}
    for n in (1, 2, 3, 5):
        out[f"indices_{n}"] = np.asarray(glass.fields.spectra_indices(n, xp=np))
    labels = np.arange(15)  # 5 fields, entries are their own positions
    out["enum_15"] = np.array([(i, j, int(c)) for i, j, c in glass.fields.enumerate_spectra(list(labels))])
    out["g2h_15"] = np.array(glass.fields.glass_to_healpix_spectra(list(labels)))
    out["h2g_15"] = np.array(glass.fields.healpix_to_glass_spectra(list(labels)))
    zz = np.linspace(0.0, 3.0, 13)
    out["hilbert_z"] = zz
    out["hilbert_shift"] = np.array([glass.fields.lognormal_shift_hilbert2011(float(z)) for z in zz])
    gls = synthetic_gls(4, 12, 2, ragged=True)
    out["cov_gls_len"] = np.array([g.shape[0] for g in gls])
    out["cov_gls"] = np.concatenate(gls)
    out["cov_full"] = glass.fields.cov_from_spectra(gls)
    out["cov_lmax5"] = glass.fields.cov_from_spectra(gls, lmax=5)
    out["cov_lmax20"] = glass.fields.cov_from_spectra(gls, lmax=20)
    out["posdef_good"] = np.asarray(bool(glass.fields.check_posdef_spectra(gls)))
    bad = [np.array([1.0, 1.0]), np.array([1.0, 1.0]), np.array([0.5, 1.5])]  # |rho| > 1 at l = 1
    out["posdef_bad"] = np.asarray(bool(glass.fields.check_posdef_spectra(bad)))
    r = np.random.default_rng(8)
    d1, b1 = r.random(5), r.random(5)
    d2, b2 = r.random((5, 3, 2)), r.random((5, 2))
    out["pw_d1"], out["pw_b1"], out["pw_d2"], out["pw_b2"] = d1, b1, d2, b2
    out["pw_1"] = glass.points.position_weights(d1)
    out["pw_1b"] = glass.points.position_weights(d1, b1)
    out["pw_1f"] = glass.points.position_weights(d1, 1.7)
    out["pw_2b"] = glass.points.position_weights(d2, b2)
    out["pw_2b1"] = glass.points.position_weights(d2, b1)
    # ---- redshifts_from_bins (glass/galaxies.py:122-185): three bins, seed 31 ----
    import glass.galaxies

    zg = np.linspace(0.0, 2.0, 51)
    nzd = {5: zg * np.exp(-zg), 2: zg**2 * np.exp(-2 * zg), 9: np.exp(-((zg - 1.0) ** 2) / 0.1)}
    binlab = np.random.default_rng(30).choice([2, 5, 9], size=200)  # 1-D: the reference's double argsort runs along the LAST axis,
    # so for N-D labels it ranks within rows and indexes only the first few draws
    out["rb_bins"], out["rb_z"] = binlab, zg
    out["rb_nz"] = np.stack([nzd[2], nzd[5], nzd[9]])
    out["rb_out"] = glass.galaxies.redshifts_from_bins(binlab, zg, nzd, rng=np.random.default_rng(31))
    out["rb_uniform"] = np.random.default_rng(31).uniform(0.0, 1.0, size=binlab.size)  # one stream: bin 2, 5, 9 runs in turn

    # ---- effective_bias (glass/points.py:75-112) ----
    zb = np.linspace(0.0, 2.0, 41)
    bzv = 1.0 + 0.5 * zb**2
    win = glass.RadialWindow(np.array([0.33, 0.5, 0.71, 0.9]), np.array([0.0, 1.0, 0.6, 0.0]), 0.6)
    out["eb_z"], out["eb_bz"], out["eb_za"], out["eb_wa"] = zb, bzv, np.asarray(win.za), np.asarray(win.wa)
    out["eb"] = np.asarray(glass.points.effective_bias(zb, bzv, win))
    wide = glass.RadialWindow(np.array([-0.5, 1.0, 2.5]), np.array([0.0, 1.0, 0.0]), 1.0)  # wider than b(z)
    out["eb_wide_za"], out["eb_wide_wa"] = np.asarray(wide.za), np.asarray(wide.wa)
    out["eb_wide"] = np.asarray(glass.points.effective_bias(zb, bzv, wide))

    # ---- regularisation (glass/algorithm.py:111-277, glass/fields.py:1055-1112) ----
    import glass.algorithm

    r = np.random.default_rng(12)
    a = r.standard_normal((6, 4, 4))
    covs = a @ np.swapaxes(a, -1, -2)
    covs[1, 0, 1] = covs[1, 1, 0] = 1.5 * np.sqrt(covs[1, 0, 0] * covs[1, 1, 1])  # |rho| > 1
    covs[4, 2, 3] = covs[4, 3, 2] = -1.2 * np.sqrt(covs[4, 2, 2] * covs[4, 3, 3])
    out["reg_cov"] = covs
    out["reg_clip"] = glass.algorithm.cov_clip(covs)
    out["reg_clip_rtol"] = glass.algorithm.cov_clip(covs, rtol=0.1)
    out["reg_nearest"] = glass.algorithm.cov_nearest(covs)
    corr = covs / np.sqrt(np.einsum("...ii,...jj->...ij", covs, covs))
    out["reg_nearcorr"] = glass.algorithm.nearcorr(corr)
    bad = synthetic_gls(3, 8, 2)
    bad[2] = 1.4 * bad[0]  # cross-spectrum (1, 0) larger than the auto-spectra allow
    out["reg_gls_len"] = np.array([g.shape[0] for g in bad])
    out["reg_gls"] = np.concatenate(bad)
    for m in ("nearest", "clip"):
        out[f"reg_spectra_{m}"] = np.stack(glass.fields.regularized_spectra(bad, method=m))
    out["reg_spectra_lmax5"] = np.stack(glass.fields.regularized_spectra(bad, lmax=5, method="clip"))
    np.savez_compressed(os.path.join(HERE, "glass_reference_spectra.npz"), **out)
    print("wrote", len(out), "arrays to glass_reference_spectra.npz")


if __name__ == "__main__":
    if "--positions" in sys.argv:
        main_positions()
    elif "--lensing-factors" in sys.argv:
        main_lensing_factors()
    elif "--solver" in sys.argv:
        main_solver()
    elif "--spectra" in sys.argv:
        main_spectra()
    elif "--displace" in sys.argv:
        main_displace()
    else:
        main()
