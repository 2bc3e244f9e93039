"""
glass_b200.lensing -- B200-native mirror of the hot-path part of ``glass/lensing.py``:
``MultiPlaneConvergence``, ``from_convergence``, ``shear_from_convergence``,
``multi_plane_matrix`` and ``multi_plane_weights``.

The multi-plane recurrence (glass/lensing.py:584-586) is one fused kernel instead of three
full-map NumPy passes; the kappa -> (psi, alpha, gamma) conversions run the forward and
spin-weighted transforms of libglassb200 instead of healpy (glass/healpix.py:107,270).
"""

from __future__ import annotations

import numpy as np
import torch

from . import _arrays as A
from . import _lib
from . import healpix as hp


class MultiPlaneConvergence:
    """Compute convergence fields iteratively from multiple matter planes
    (glass/lensing.py:431-606).  Same state machine, properties and error behaviour; the
    convergence planes live on the GPU and ``kappa`` is returned in the array kind of the
    ``delta`` that was added (CUDA tensor -> the internal buffer, which like the
    reference's is recycled two calls later; NumPy -> a host copy)."""

    def __init__(self, cosmo) -> None:
        self.cosmo = cosmo
        self.z2 = 0.0
        self.z3 = 0.0
        self.x3 = 0.0
        self.w3 = 0.0
        self.r23 = 1.0
        self.delta3 = None
        self.kappa2 = None
        self.kappa3 = None
        self._host = False
        self._host_delta = None  # the caller's own NumPy matter plane (host mode): what `delta` hands back
        self._host_kappa = None  # host copy of the current convergence plane, made once per plane on first access
        self._pin = [None, None]  # two page-locked staging buffers, recycled like the planes themselves
        self._npin = 0
        self._like = None  # map-shaped tensor sizing the hand-off of an empty block (dist.multi_plane_block)

    def _set_state_map(self, name: str, value) -> None:
        """Install a received map (delta3, kappa2, kappa3) of the recurrence state
        (glass_b200.dist.recv_multi_plane_state)."""
        setattr(self, name, value)
        self._host_kappa = None

    def add_window(self, delta, w) -> None:
        """Add a mass plane from a window function (glass/lensing.py:489-509)."""
        zsrc = w.zeff
        za, wa = A.to_np(w.za), A.to_np(w.wa)
        lens_weight = float(np.trapezoid(wa, za) / np.interp(zsrc, za, wa))
        self.add_plane(delta, zsrc, lens_weight)

    @A.nvtx("glass.MultiPlaneConvergence.add_plane")
    def add_plane(self, delta, zsrc, wlens: float = 1.0) -> None:
        """Add a mass plane at redshift ``zsrc`` (glass/lensing.py:511-586)."""
        if zsrc <= self.z3:
            msg = "source redshift must be increasing"
            raise ValueError(msg)
        device, on_device = A.pick_device(delta)
        self._host = not on_device
        self._host_delta = None if on_device else delta
        self._host_kappa = None
        delta_d = A.to_dev(delta, device)

        # cycle mass plane, redshifts and weights (lensing.py:544-549)
        delta2, self.delta3 = self.delta3, delta_d
        z1, self.z2, self.z3 = self.z2, self.z3, zsrc
        w2, self.w3 = self.w3, wlens

        # extrapolation law (lensing.py:551-563): five scalar cosmology calls on the host
        x2, self.x3 = (
            self.x3,
            self.cosmo.transverse_comoving_distance(self.z3) / self.cosmo.hubble_distance,
        )
        r12 = self.r23
        r13, self.r23 = np.asarray(
            self.cosmo.transverse_comoving_distance([z1, self.z2], self.z3)
        ) / (self.cosmo.hubble_distance * self.x3)
        t = r13 / r12

        # lensing weight of mass plane to be added (lensing.py:565-569)
        f = 3 * self.cosmo.Omega_m0 / 2
        f *= x2 * self.r23
        f *= (1 + self.z2) / self.cosmo.H_over_H0(self.z2)
        f *= w2

        if self.kappa2 is None:
            self.kappa2 = torch.zeros_like(delta_d)
            self.kappa3 = torch.zeros_like(delta_d)

        # cycle convergence planes and update in place of the oldest (lensing.py:580-586)
        self.kappa2, self.kappa3 = self.kappa3, self.kappa2
        lib = _lib.load()
        with torch.cuda.device(device):
            st = torch.cuda.current_stream(device).cuda_stream
            _lib.check(
                lib.glb_multiplane_update(
                    self.kappa3.data_ptr(),
                    self.kappa2.data_ptr(),
                    None if delta2 is None else delta2.data_ptr(),
                    0.0,
                    self.kappa3.numel(),
                    float(t),
                    float(f),
                    st,
                ),
                "glb_multiplane_update",
            )

    @property
    def zsrc(self):
        """The redshift of the current convergence plane."""
        return self.z3

    @property
    def kappa(self):
        """The current convergence plane.  Host mode (NumPy planes in): ONE device -> host copy per plane,
        through page-locked staging at the link rate, however often the property is read; the array is
        recycled two planes later, like the reference's own buffers (glass/lensing.py:580)."""
        if self.kappa3 is None:
            return None
        if not self._host:
            return self.kappa3
        if self._host_kappa is None:
            k = self.kappa3
            if not k.is_cuda:
                self._host_kappa = k.numpy()
            else:
                slot = self._npin & 1
                self._npin += 1
                if self._pin[slot] is None or self._pin[slot].shape != k.shape:
                    self._pin[slot] = torch.empty(k.shape, dtype=k.dtype, pin_memory=True)
                self._pin[slot].copy_(k, non_blocking=True)
                torch.cuda.current_stream(k.device).synchronize()
                self._host_kappa = self._pin[slot].numpy()
        return self._host_kappa

    @property
    def delta(self):
        """The current matter plane (host mode: the array that was added, as in the reference)."""
        if self.delta3 is None:
            return None
        if self._host:
            return self._host_delta if self._host_delta is not None else self.delta3.cpu().numpy()
        return self.delta3

    @property
    def wlens(self) -> float:
        """The weight of the current matter plane."""
        return self.w3


def multi_plane_matrix(shells, cosmo):
    """Compute the matrix of lensing contributions from each shell
    (glass/lensing.py:609-635): row i = kappa after adding shells 0..i, for unit deltas."""
    mpc = MultiPlaneConvergence(cosmo)
    n = len(shells)
    device = torch.device("cuda", hp._device_index())
    wmat = torch.eye(n, dtype=torch.float64, device=device)
    rows = []
    for i, w in enumerate(shells):
        mpc.add_window(wmat[i].clone(), w)
        rows.append(mpc.kappa.clone())
    return torch.stack(rows).cpu().numpy()


def multi_plane_weights(weights, shells, cosmo):
    """Compute effective weights for multi-plane convergence (glass/lensing.py:638-684)."""
    weights = A.to_np(weights)
    shape = weights.shape
    if not shape or shape[0] != len(shells):
        msg = "shape mismatch between weights and shells"
        raise ValueError(msg)
    weights = weights / np.sum(weights, axis=0)
    mat = multi_plane_matrix(shells, cosmo)
    return np.matmul(mat.T, weights)


def _kappa_alm(kappa, lmax, niter, ring_weights):
    device, on_device = A.pick_device(kappa)
    k = A.to_dev(kappa, device)
    nside = hp.get_nside(k)
    if lmax is None:
        lmax = 3 * nside - 1
    alm = hp.map2alm(k, lmax=lmax, pol=False, use_pixel_weights=True, niter=niter, ring_weights=ring_weights)
    return alm, nside, lmax, device, on_device


def _pixwin_ratio(nside, lmax, pixwin):
    """pw2/pw0 of glass/lensing.py:361-362,421-422: ``hp.pixwin(nside, lmax=lmax, pol=True)``,
    generated numerically (glass_b200.pixwin; healpy reads data files that are not available
    offline), or the caller's own tables ``pixwin=(pw0, pw2)``, e.g. healpy's.  w^P vanishes below
    l = 2, where the shear factor is zero anyway: the ratio is taken as 0 there."""
    if pixwin is None:
        pixwin = hp.pixwin(nside, lmax=lmax, pol=True)
    pw0, pw2 = (A.to_np(p)[: lmax + 1] for p in pixwin)
    return np.divide(pw2, pw0, out=np.zeros_like(pw2), where=pw0 != 0)


def _convergence_factors(nside, lmax, discretized, pixwin):
    """The three successive ``almxfl`` factors of glass/lensing.py:316-363, with the reference's
    arithmetic: kappa -> psi ``-2 / (l (l+1))``, psi -> alpha ``sqrt(l (l+1))``, alpha -> gamma
    ``sqrt((l-1)(l+2)) / 2 [* pw2 / pw0]`` (zero at l = 0)."""
    ell = np.arange(lmax + 1, dtype=np.float64)
    f_psi = np.zeros(lmax + 1)
    f_psi[1:] = -2.0 / (ell[1:] * (ell[1:] + 1))
    f_alpha = np.sqrt(ell * (ell + 1))
    f_gamma = np.zeros(lmax + 1)
    f_gamma[1:] = np.sqrt((ell[1:] - 1) * (ell[1:] + 2))
    f_gamma /= 2
    if discretized:
        f_gamma *= _pixwin_ratio(nside, lmax, pixwin)
    return f_psi, f_alpha, f_gamma


@A.nvtx("glass.from_convergence")
def from_convergence(  # noqa: PLR0913
    kappa,
    lmax: int | None = None,
    *,
    potential: bool = False,
    deflection: bool = False,
    shear: bool = False,
    discretized: bool = True,
    pixwin=None,
    niter: int = 3,
    ring_weights=None,
):
    r"""
    Compute other weak lensing maps from the convergence (glass/lensing.py:185-371).

    Returns the requested maps in the order potential, deflection, shear; deflection and
    shear are complex128 maps.  Extensions over the reference signature: ``pixwin`` (see
    :func:`_pixwin_ratio`), ``niter`` and ``ring_weights`` (see
    :func:`glass_b200.healpix.map2alm`).
    """
    if not (potential or deflection or shear):
        return ()
    alm, nside, lmax, device, on_device = _kappa_alm(kappa, lmax, niter, ring_weights)
    if shear:
        f_psi, f_alpha, f_gamma = _convergence_factors(nside, lmax, discretized, pixwin)
    else:
        f_psi, f_alpha, _ = _convergence_factors(nside, lmax, False, None)
    results = ()

    def out(t):
        return t if on_device else t.cpu().numpy()

    # convert convergence to potential (lensing.py:316-322)
    alm = hp.almxfl(alm, f_psi, inplace=True)
    if potential:
        psi = hp.alm2map_batch(alm[None], nside, lmax)[0]
        results += (out(psi),)
    if not (deflection or shear):
        return results
    # deflection alms (lensing.py:337-339)
    alm = hp.almxfl(alm, f_alpha, inplace=True)
    if deflection:
        a1, a2 = hp.alm2map_spin([alm, None], nside, 1, lmax)
        results += (out(torch.complex(a1, a2)),)
    if not shear:
        return results
    # shear alms (lensing.py:353-363)
    alm = hp.almxfl(alm, f_gamma, inplace=True)
    g1, g2 = hp.alm2map_spin([alm, None], nside, 2, lmax)
    results += (out(torch.complex(g1, g2)),)
    return results


@A.nvtx("glass.shear_from_convergence")
def shear_from_convergence(kappa, lmax: int | None = None, *, discretized: bool = True, pixwin=None, niter: int = 3, ring_weights=None):
    r"""
    Weak lensing shear from convergence (glass/lensing.py:374-428; deprecated in the
    reference in favour of :func:`from_convergence`, but what every example calls).
    Returns ``[gamma1, gamma2]``.
    """
    if getattr(kappa, "ndim", 1) == 2:
        return _shear_from_convergence_stack(kappa, lmax, discretized, pixwin, niter, ring_weights)
    alm, nside, lmax, device, on_device = _kappa_alm(kappa, lmax, niter, ring_weights)
    alm = hp.almxfl(alm, _shear_factor(nside, lmax, discretized, pixwin), inplace=True)
    g1, g2 = hp.alm2map_spin([alm, None], nside, 2, lmax)
    return [g1, g2] if on_device else [g1.cpu().numpy(), g2.cpu().numpy()]


def _shear_factor(nside, lmax, discretized, pixwin):
    """-sqrt((l+2)(l+1)l(l-1)) / max(l(l+1), 1) [* pw2/pw0]  (glass/lensing.py:414-422)."""
    ell = np.arange(lmax + 1)
    fl = np.sqrt((ell + 2) * (ell + 1) * ell * (ell - 1))
    fl /= np.clip(ell * (ell + 1), 1, None)
    fl *= -1
    if discretized:
        fl *= _pixwin_ratio(nside, lmax, pixwin)
    return fl


def _shear_from_convergence_stack(kappa, lmax, discretized, pixwin, niter, ring_weights):
    """Extension: ``kappa`` of shape (n, npix) -- several convergence planes at once (the planes
    of a rank's block of shells, ``glass_b200.dist.multi_plane_block``).  The map2alm refinement
    syntheses of up to four planes share one Legendre recurrence (34 ms per map at nside 4096
    against 50 ms alone), and so do their spin-2 syntheses (``hp.alm2map_spin_batch``); returns
    ``[gamma1, gamma2]`` of shape (n, npix)."""
    device, on_device = A.pick_device(kappa)
    k = A.to_dev(kappa, device)
    nside = hp.get_nside(k[0])
    if lmax is None:
        lmax = 3 * nside - 1
    alms = hp.map2alm(list(k), lmax=lmax, pol=False, use_pixel_weights=True, niter=niter, ring_weights=ring_weights)
    fl = _shear_factor(nside, lmax, discretized, pixwin)
    stack = torch.stack([hp.almxfl(alm, fl, inplace=True) for alm in alms])
    g1, g2 = hp.alm2map_spin_batch(stack, nside, 2, lmax)
    return [g1, g2] if on_device else [g1.cpu().numpy(), g2.cpu().numpy()]


def deflect(lon, lat, alpha, xp=None):
    """Apply deflections to positions (glass/lensing.py:687-778; deprecated in the reference in
    favour of ``glass.displace``, which moves the longitude the other way)."""
    from .points import _displace

    return _displace(lon, lat, alpha, deflect=True)
