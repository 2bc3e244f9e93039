"""
glass_b200.points -- B200-native mirror of the hot-path part of ``glass/points.py``:
``positions_from_delta`` with its helpers and the two built-in bias models.

The reference walks the map in 1000-pixel steps in a Python loop, repeats pixel indices
with ``np.repeat`` and calls ``healpix.randang`` per batch (glass/points.py:389-440).  Here
one fused kernel makes the per-pixel Poisson counts, a scan turns them into offsets, and one
kernel writes every galaxy's position; the generator then hands out slices that follow the
reference's batch-cut rule exactly (greedy pixel-aligned batches of at most ``batch``
points, "first pixel alone" rule, points.py:409-424).
"""

from __future__ import annotations

import ctypes as C
import itertools
import math
from typing import Callable

import numpy as np
import torch

from . import _arrays as A
from . import _lib
from . import healpix as hp
from . import rng as _rng

ARCMIN2_SPHERE = 60**6 // 100 / math.pi  # glass/points.py:72


def linear_bias(delta, b):
    r"""Linear bias model :math:`\delta_g = b \, \delta` (glass/points.py:115-134)."""
    return b * delta


def loglinear_bias(delta, b):
    r"""Log-linear bias model :math:`\ln(1+\delta_g) = b \ln(1+\delta)` (glass/points.py:137-160)."""
    if isinstance(delta, torch.Tensor):
        return torch.expm1(torch.log1p(delta) * b)
    delta_g = np.log1p(delta)
    delta_g *= b
    return np.expm1(delta_g)


def _trapezoid_product(f, *ff):
    """Trapezoidal integral of a product of piecewise-linear functions on the union of their
    grids, restricted to the common support (glass/arraytools.py:158-194)."""
    x, _ = f
    for x_, _y in ff:
        x = np.union1d(x[(x >= x_[0]) & (x <= x_[-1])], x_[(x_ >= x[0]) & (x_ <= x[-1])])
    y = np.interp(x, *f)
    for f_ in ff:
        y *= np.interp(x, *f_)
    return np.trapezoid(y, x)


def effective_bias(z, bz, w):
    r"""Effective bias :math:`\bar b = \int b(z) w(z) dz / \int w(z) dz` of a redshift-dependent
    bias for a radial window (glass/points.py:75-112).  A few hundred numbers: host arithmetic."""
    z, bz, za, wa = (np.asarray(A.to_np(a), dtype=np.float64) for a in (z, bz, w.za, w.wa))
    return _trapezoid_product((z, bz), (za, wa)) / np.trapezoid(wa, za)


def position_weights(densities, bias=None):
    """Relative weight of each shell for angular clustering: densities normalised over the first
    (shell) axis, times an optional linear bias per shell (glass/points.py:610-651).  A handful
    of numbers per shell: host arithmetic on NumPy arrays, torch arithmetic on tensors."""
    is_t = isinstance(densities, torch.Tensor) or isinstance(bias, torch.Tensor)
    xp_asarray = (lambda a: torch.as_tensor(a, dtype=torch.float64)) if is_t else (lambda a: np.asarray(a, dtype=np.float64))
    densities = xp_asarray(densities)
    if bias is not None:
        bias = xp_asarray(bias)
        # broadcast_first (glass/arraytools.py:24-44): the first axis is common, the others broadcast
        mv = torch.movedim if is_t else np.moveaxis
        bc = torch.broadcast_tensors if is_t else np.broadcast_arrays
        moved = [mv(a, 0, -1) if a.ndim else a for a in (densities, bias)]
        densities, bias = (mv(a, -1, 0) if a.ndim else a for a in bc(*moved))
    densities = densities / densities.sum(0)
    if bias is not None:
        densities = densities * bias
    return densities


_BIAS_NONE, _BIAS_LINEAR, _BIAS_LOGLINEAR = 0, 1, 2


def _bias_code(bias, bias_model):
    """(kernel code, needs_prepass): built-in models are fused; any other callable is
    applied as a separate pass on the device tensor (glass/points.py:243-249)."""
    if bias is None:
        return _BIAS_NONE, False
    name = getattr(bias_model, "__name__", "")
    mod = getattr(bias_model, "__module__", "") or ""
    if bias_model is linear_bias or (name == "linear_bias" and mod.startswith("glass")):
        return _BIAS_LINEAR, False
    if bias_model is loglinear_bias or (name == "loglinear_bias" and mod.startswith("glass")):
        return _BIAS_LOGLINEAR, False
    return _BIAS_NONE, True


#: expected galaxies per pixel up to which K6 emits the galaxy -> pixel LIST instead of per-pixel
#: counts and offsets (8 B per galaxy against 16 B per pixel; the cut rule and the position kernel
#: then work from the list alone)
LIST_MODE_MAX_DENSITY = 1.0


class _Population:
    """Counts and positions of one population on the device.

    ``mode``: "scan" -- per-pixel ``counts`` and exclusive offsets ``off`` (K6+K7), positions by
    walking pixels; "list" -- only the galaxy -> pixel list ``gpix`` (= np.repeat(arange, counts),
    glass/points.py:426), cuts and positions from it; "auto" -- the list for sparse maps (expected
    density <= LIST_MODE_MAX_DENSITY galaxies per pixel), the scan otherwise.  Same galaxies
    either way: counts are a function of (seed, stream, pixel), positions of (seed, stream,
    galaxy index)."""

    def __init__(self, delta_k, vis_k, ngal_k, bias_k, bias_model, remove_monopole, seed, stream_id, counts_in, device, want_nbar=False,
                 mode: str = "auto"):
        lib = _lib.load()
        self.lib, self.device = lib, device
        code, prepass = _bias_code(bias_k, bias_model)
        d = A.to_dev(delta_k, device)
        if prepass:
            d = A.to_dev(bias_model(d, bias_k), device)
        self.npix = d.numel()
        self.nside = hp.npix2nside(self.npix)
        v = None if vis_k is None else A.to_dev(vis_k, device)
        scale = ARCMIN2_SPHERE / self.npix * float(ngal_k)  # same order as points.py:287
        if mode == "auto":
            mode = "list" if scale <= LIST_MODE_MAX_DENSITY else "scan"
        self.mode = mode
        self.counts = self.off = self.gpix = None
        cap = 0
        if mode == "scan":
            self.counts = torch.empty(self.npix, dtype=torch.int64, device=device)
            self.off = torch.empty(self.npix + 1, dtype=torch.int64, device=device)
        else:
            # first guess of the list length: the mean expected count (delta averages to ~0, vis <= 1)
            cap = int(1.25 * scale * self.npix) + 65536
            self.gpix = torch.empty(cap, dtype=torch.int64, device=device)
        self.nbar = torch.empty(self.npix, dtype=torch.float64, device=device) if want_nbar else None
        ws = torch.empty(int(lib.glb_points_workspace_bytes(self.npix)), dtype=torch.uint8, device=device)
        cin = None if counts_in is None else A.to_dev(counts_in, device, torch.int64)
        total = torch.zeros(1, dtype=torch.int64, device=device)
        self.seed, self.stream_id = seed, stream_id
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731

        def launch():
            _lib.check(
                lib.glb_points_counts(
                    self.npix, d.data_ptr(), ptr(v), code, float(bias_k) if (bias_k is not None and not prepass) else 0.0, scale,
                    int(bool(remove_monopole)), ptr(cin), C.c_uint64(seed), C.c_uint32(stream_id), ptr(self.nbar), ptr(self.counts),
                    ptr(self.off), ptr(self.gpix), cap, total.data_ptr(), ws.data_ptr(), torch.cuda.current_stream(device).cuda_stream,
                ),
                "glb_points_counts",
            )
            return int(total.item())

        self.total = launch()
        if self.gpix is not None and self.total > cap:  # the guess was too small: same counts again, into a list that fits
            cap = self.total
            self.gpix = torch.empty(cap, dtype=torch.int64, device=device)
            self.total = launch()

    def cuts(self, batch: int, chunk: int = 4096):
        """Pixel ranges (start, stop, npoints) of glass/points.py:409-437, walked on the device by
        ``glb_points_cuts`` (csrc/points_cuts.cuh, host-tested against the reference's loop): one
        device->host copy per ``chunk`` cuts.

        Closed form of the reference's 1000-pixel stepping loop.  The loop advances in groups of
        1000 pixels (counted from ``start``) until the group in which the running total reaches
        min(batch, remaining) -- at pixel q* -- and cuts inside that group with
        searchsorted(side="right"): after the last pixel whose running total is still <= batch,
        but not beyond the end of the group.  Hence stop = min(p, end of the group of q*) with p the
        largest index whose offset is <= off[start] + batch.  Consequences the reference shares: a
        batch may be EMPTY when zero-count pixels precede a pixel that alone exceeds ``batch``; a
        first pixel that alone exceeds ``batch`` is taken by itself; on an exact fit, and for the
        last batch, trailing empty pixels are included up to the end of the group."""
        start, remaining = 0, self.total
        cuts = torch.empty((chunk, 3), dtype=torch.int64, device=self.device)
        state = torch.empty(3, dtype=torch.int64, device=self.device)
        gpix = getattr(self, "gpix", None)
        while remaining > 0:
            st = torch.cuda.current_stream(self.device).cuda_stream
            if gpix is not None:
                _lib.check(
                    self.lib.glb_points_cuts_list(gpix.data_ptr(), self.total, self.npix, int(batch), start, remaining, chunk, cuts.data_ptr(),
                                                  state.data_ptr(), st),
                    "glb_points_cuts_list",
                )
            else:
                _lib.check(
                    self.lib.glb_points_cuts(self.off.data_ptr(), self.npix, int(batch), start, remaining, chunk, cuts.data_ptr(), state.data_ptr(), st),
                    "glb_points_cuts",
                )
            k, start, remaining = (int(v) for v in state.cpu().numpy())
            for a, b, n in cuts[:k].cpu().numpy():
                yield int(a), int(b), int(n)

    def batches(self, batch: int, group: int = 1 << 26):
        """(lon, lat, npoints) of every batch, as slices of positions written by ONE launch of the
        fill kernel per ``group`` galaxies (a draw depends on the galaxy's global index only, so
        the result does not depend on how the launches are cut)."""
        pending, first_pix, ngal, done = [], None, 0, 0

        def flush():
            nonlocal pending, first_pix, ngal, done
            if pending:
                lon, lat, _ = self.fill(first_pix, pending[-1][0], ngal, first_galaxy=done)
                pos = 0
                for _stop, n in pending:
                    yield lon[pos : pos + n], lat[pos : pos + n], n
                    pos += n
                done += ngal
            pending, first_pix, ngal = [], None, 0

        for start, stop, n in self.cuts(batch):
            if first_pix is None:
                first_pix = start
            pending.append((stop, n))
            ngal += n
            if ngal >= group:
                yield from flush()
        yield from flush()

    def fill(self, start: int, stop: int, n: int, uv=None, want_ipix=False, first_galaxy: int | None = None):
        """Positions of the ``n`` galaxies in ring pixels [start, stop).  List mode needs the index of
        the range's first galaxy: ``first_galaxy`` if the caller tracks it (batches are consecutive),
        else it is looked up in the list."""
        lon = torch.empty(n, dtype=torch.float64, device=self.device)
        lat = torch.empty(n, dtype=torch.float64, device=self.device)
        ipix = torch.empty(n, dtype=torch.int64, device=self.device) if (want_ipix and self.gpix is None) else None
        if n == 0:
            return lon, lat, (torch.empty(0, dtype=torch.int64, device=self.device) if want_ipix else None)
        u = v = None
        if uv is not None:
            uu, vv = uv(n) if callable(uv) else uv
            u, v = A.to_dev(uu, self.device), A.to_dev(vv, self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        if self.gpix is not None:
            if first_galaxy is None:
                first_galaxy = int(torch.searchsorted(self.gpix[: self.total], torch.tensor(start, device=self.device)).item())
            g0 = int(first_galaxy)
            _lib.check(
                self.lib.glb_points_fill_list(self.nside, self.gpix.data_ptr(), g0, g0 + n, None if u is None else u.data_ptr(),
                                              None if v is None else v.data_ptr(), C.c_uint64(self.seed), C.c_uint32(self.stream_id),
                                              lon.data_ptr(), lat.data_ptr(), st),
                "glb_points_fill_list",
            )
            return lon, lat, (self.gpix[g0 : g0 + n] if want_ipix else None)
        _lib.check(
            self.lib.glb_points_fill(
                self.nside,
                self.counts.data_ptr(),
                self.off.data_ptr(),
                start,
                stop,
                None if u is None else u.data_ptr(),
                None if v is None else v.data_ptr(),
                C.c_uint64(self.seed),
                C.c_uint32(self.stream_id),
                lon.data_ptr(),
                lat.data_ptr(),
                None if ipix is None else ipix.data_ptr(),
                st,
            ),
            "glb_points_fill",
        )
        return lon, lat, ipix


def positions_from_delta(  # noqa: PLR0913
    ngal,
    delta,
    bias=None,
    vis=None,
    *,
    bias_model: Callable = linear_bias,
    remove_monopole: bool = False,
    batch: int = 1_000_000,
    rng=None,
):
    """
    Generate positions tracing a density contrast (glass/points.py:443-540).

    Yields ``(lon, lat, count)`` batches: longitudes/latitudes in degrees and the number
    of points (an int, or an int64 array of shape ``dims`` with the count in the
    population's slot when the inputs have leading "population" axes).  Arrays are CUDA
    tensors when ``delta`` (or ``vis``) is a CUDA tensor, NumPy arrays otherwise.
    """
    if not callable(bias_model):
        raise TypeError("bias_model must be callable")

    device, on_device = A.pick_device(delta, vis, ngal, bias)
    deviates = rng if isinstance(rng, _rng.Deviates) else None
    seed = _rng.seed_from(rng)

    inputs = [(ngal, 0), (delta, 1)]
    if bias is not None:
        inputs.append((bias, 0))
    if vis is not None:
        inputs.append((vis, 1))
    dims, leads, _trails = A.broadcast_leading_axes(*inputs)
    lead = dict(zip(["ngal", "delta"] + (["bias"] if bias is not None else []) + (["vis"] if vis is not None else []), leads))

    with torch.cuda.device(device):
        for ipop, k in enumerate(itertools.product(*map(range, dims))):
            delta_k = A.take_leading(delta, lead["delta"], dims, k)
            ngal_k = A.take_leading(ngal, lead["ngal"], dims, k) if lead["ngal"] else ngal
            bias_k = None if bias is None else (A.take_leading(bias, lead["bias"], dims, k) if lead["bias"] else bias)
            vis_k = None if vis is None else A.take_leading(vis, lead["vis"], dims, k)
            counts_in = deviates.next_poisson() if (deviates is not None and deviates.poisson is not None) else None
            pop = _Population(
                delta_k, vis_k, float(A.to_np(ngal_k)), None if bias_k is None else float(A.to_np(bias_k)),
                bias_model, remove_monopole, seed, ipop, counts_in, device,
            )
            if pop.total == 0:
                continue
            if dims:
                cmask = np.zeros(dims, dtype=np.int64)
                cmask[k] = 1
            else:
                cmask = 1
            uv = deviates.uv if deviates is not None else None
            if uv is not None:  # parity mode: the supplied (u, v) are consumed batch by batch
                def per_cut(pop=pop):
                    g0 = 0
                    for start, stop, n in pop.cuts(batch):
                        lon, lat, _ = pop.fill(start, stop, n, uv, first_galaxy=g0)
                        g0 += n
                        yield lon, lat, n

                it = per_cut()
            else:
                it = pop.batches(batch)
            if on_device:
                for lon, lat, n in it:
                    yield lon, lat, n * cmask
            else:
                yield from ((lo, la, n * cmask) for lo, la, n in A.host_slices(it))


def uniform_positions(ngal, *, rng=None, xp=None):
    """
    Generate positions uniformly over the sphere (glass/points.py:543-607).

    ``ngal`` (scalar or array): expected number of points per arcmin2.  Yields, per entry
    of ``ngal`` in C order, ``(lon, lat, count)`` with ``lon = uniform(-180, 180)``,
    ``lat = degrees(asin(uniform(-1, 1)))`` and ``count`` an int (scalar ``ngal``) or an
    int64 array of ``ngal``'s shape holding the count in that entry's slot.  Arrays are CUDA
    tensors when ``ngal`` is a CUDA tensor or ``xp is torch``, NumPy arrays otherwise.

    The total counts are Poisson(ARCMIN2_SPHERE * ngal) drawn on the host (a handful of
    scalars); the two uniforms per point come from Philox keyed by (seed, population, point
    index), or from ``rng=Deviates(poisson=[counts], uniform=...)`` in parity mode
    (``uniform``: callable n -> (u1, u2), the deviates of lon and lat).
    """
    device, on_device = A.pick_device(ngal)
    on_device = on_device or (xp is torch)
    deviates = rng if isinstance(rng, _rng.Deviates) else None
    seed = _rng.seed_from(rng)
    lam = ARCMIN2_SPHERE * np.asarray(A.to_np(ngal), dtype=np.float64)
    if deviates is not None and deviates.poisson is not None:
        ngal_sphere = np.asarray(deviates.next_poisson(), dtype=np.int64).reshape(lam.shape)
    else:
        ngal_sphere = np.asarray(np.random.default_rng(seed).poisson(lam), dtype=np.int64)
    dims = ngal_sphere.shape
    lib = _lib.load()
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        for ipop, k in enumerate(np.ndindex(dims)):
            n = int(ngal_sphere[k])
            lon = torch.empty(n, dtype=torch.float64, device=device)
            lat = torch.empty(n, dtype=torch.float64, device=device)
            u1 = u2 = None
            if deviates is not None and deviates.uniform is not None:
                a, b = deviates.uniform(n)
                u1, u2 = A.to_dev(a, device), A.to_dev(b, device)
            _lib.check(
                lib.glb_uniform_positions(
                    n, None if u1 is None else u1.data_ptr(), None if u2 is None else u2.data_ptr(),
                    C.c_uint64(seed), C.c_uint32(ipop), lon.data_ptr(), lat.data_ptr(), st,
                ),
                "glb_uniform_positions",
            )
            if dims:
                count = np.zeros(dims, dtype=np.int64)
                count[k] = n
            else:
                count = n
            if on_device:
                yield lon, lat, count
            else:
                yield lon.cpu().numpy(), lat.cpu().numpy(), count


def _alpha_parts(alpha, device):
    """(alpha1, alpha2, stride) device views of a displacement given as a complex array or with a
    leading axis of size 2 (glass/points.py:692-696)."""
    is_complex = (isinstance(alpha, torch.Tensor) and alpha.is_complex()) or (
        not isinstance(alpha, torch.Tensor) and np.iscomplexobj(alpha)
    )
    if is_complex:
        a = A.to_dev(alpha, device, torch.complex128)
        r = torch.view_as_real(a.contiguous())  # [..., 2] float64, interleaved
        return a.shape, r, r.data_ptr(), r.data_ptr() + 8, 2
    a = A.to_dev(alpha, device)
    if a.shape[0] != 2:
        raise ValueError("alpha must be complex-valued or have a leading axis of size 2")
    a = a.contiguous()
    return a.shape[1:], a, a[0].data_ptr(), a[1].data_ptr(), 1


def _displace(lon, lat, alpha, deflect: bool):
    device, on_device = A.pick_device(lon, lat, alpha)
    shape_a, keep, p1, p2, stride = _alpha_parts(alpha, device)
    lon_d, lat_d = A.to_dev(lon, device), A.to_dev(lat, device)
    shape = torch.broadcast_shapes(lon_d.shape, lat_d.shape, tuple(shape_a))
    if tuple(shape_a) != tuple(shape):
        # broadcast the displacement (rare: scalar alpha): materialise it as complex128
        if stride == 2:
            full = torch.view_as_complex(keep).expand(shape).contiguous()
        else:
            full = torch.complex(keep[0], keep[1]).expand(shape).contiguous()
        keep = torch.view_as_real(full)
        p1, p2, stride = keep.data_ptr(), keep.data_ptr() + 8, 2
    lon_d = lon_d.expand(shape).contiguous()
    lat_d = lat_d.expand(shape).contiguous()
    n = lon_d.numel()
    out_lon = torch.empty(shape, dtype=torch.float64, device=device)
    out_lat = torch.empty(shape, dtype=torch.float64, device=device)
    lib = _lib.load()
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        _lib.check(
            lib.glb_displace(lon_d.data_ptr(), lat_d.data_ptr(), p1, p2, stride, int(deflect), n, out_lon.data_ptr(), out_lat.data_ptr(), st),
            "glb_displace",
        )
    if on_device:
        return out_lon, out_lat
    return out_lon.cpu().numpy(), out_lat.cpu().numpy()


def displace(lon, lat, alpha):
    """Displace positions on the sphere (glass/points.py:654-716): the exponential map.  ``alpha``
    complex, or real with a leading axis of size 2.  One kernel, 48 bytes per point."""
    return _displace(lon, lat, alpha, deflect=False)


def displacement(from_lon, from_lat, to_lon, to_lat):
    """Complex displacement between two sets of positions in degrees (glass/points.py:719-772)."""
    device, on_device = A.pick_device(from_lon, from_lat, to_lon, to_lat)
    ts = torch.broadcast_tensors(*(A.to_dev(x, device) for x in (from_lon, from_lat, to_lon, to_lat)))
    ts = [t.contiguous() for t in ts]
    out = torch.empty(ts[0].shape, dtype=torch.complex128, device=device)
    lib = _lib.load()
    with torch.cuda.device(device):
        st = torch.cuda.current_stream(device).cuda_stream
        _lib.check(
            lib.glb_displacement(ts[0].data_ptr(), ts[1].data_ptr(), ts[2].data_ptr(), ts[3].data_ptr(), out.numel(), out.data_ptr(), st),
            "glb_displacement",
        )
    return out if on_device else out.cpu().numpy()
