"""
glass_b200.fields -- B200-native mirror of the hot-path part of ``glass/fields.py``:
``iternorm``, ``cls2cov``, ``getcl``, ``generate`` (+ the deprecated
``generate_gaussian`` / ``generate_lognormal``), ``lognormal_fields``,
``gaussian_fields``.

Same signatures, argument meaning and error messages as the reference.  Per shell the
reference does (glass/fields.py:404-429, 884-894): draw N_lm complex normals on one CPU
thread, combine with the iternorm weights through full-size temporaries, re-order
l-major -> m-major in a Python loop, call healpy.alm2map, then apply the transformation
as one more pass.  Here: Philox normals are drawn directly in m-major order on the GPU
(glb_alm_draw), combined in one fused kernel (glb_alm_combine), several shells are
synthesised together so they share one Legendre recurrence (glb_alm2map with a batch),
and the transformation is fused into the ring-FFT store.

Array rule: if any ``gls`` entry is a CUDA tensor the maps are yielded as CUDA tensors,
otherwise as NumPy arrays (device->host copies overlap the next shells' compute).
"""

from __future__ import annotations

import collections
import contextlib
import ctypes as C
import functools
import math
import warnings
import weakref
from typing import Callable, Iterable, Iterator, Sequence

import numpy as np
import torch

from . import _lib, grf
from . import healpix as hp
from . import rng as _rng

# how many shells are synthesised together (up to 8): eight share one recurrence on the INT8 tensor-core
# Legendre path (nside >= 1024), smaller groups share one on the FP64 pipe in fours, twos or alone
SHT_BATCH = 8


class _PinnedPool:
    """Page-locked host buffers for the NumPy-out path.

    cudaHostAlloc of a 1.6 GB map costs 0.2-0.9 s, far more than computing the map, so
    buffers are recycled: a buffer is handed out again only when the NumPy array that was
    yielded from it (and every view of it) has been garbage collected -- the caller owns
    what it was given for as long as it keeps it."""

    def __init__(self):
        self._bufs: dict[int, list] = {}

    def take(self, n: int) -> torch.Tensor:
        for ent in self._bufs.setdefault(n, []):
            if ent[1] is None or ent[1]() is None:
                ent[1] = _BUSY
                return ent[0]
        t = torch.empty(n, dtype=torch.float64, pin_memory=True)
        self._bufs[n].append([t, _BUSY])
        return t

    def lend(self, t: torch.Tensor) -> np.ndarray:
        arr = t.numpy()
        for ent in self._bufs.get(t.numel(), []):
            if ent[0] is t:
                ent[1] = weakref.ref(arr)
        return arr

    def clear(self):
        self._bufs.clear()


def _BUSY():  # sentinel "weakref" that is always alive
    return True


_PINNED = _PinnedPool()


def deprecated(msg: str, /):
    def decorator(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            warnings.warn(msg, category=DeprecationWarning, stacklevel=2)
            return func(*args, **kwargs)

        return wrapper

    return decorator


def _inv_triangle_number(triangle_number: int) -> int:
    """glass/fields.py:69-80."""
    n = math.floor(math.sqrt(2 * triangle_number))
    if n * (n + 1) // 2 != triangle_number:
        msg = f"not a triangle number: {triangle_number}"
        raise ValueError(msg)
    return n


def nfields_from_nspectra(nspectra: int) -> int:
    """glass/fields.py:83-98."""
    try:
        n = _inv_triangle_number(nspectra)
    except ValueError:
        msg = f"invalid number of spectra: {nspectra}"
        raise ValueError(msg) from None
    return n


def _np(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def _gls_to_host(gls):
    """Host (NumPy) copies of the spectra with ONE device->host transfer for all CUDA tensors.

    The spectra are tiny (<= lmax+1 doubles each) and only feed host-side recursions
    (cls2cov / iternorm / cltovar); fetching them one by one would synchronise the stream
    once per spectrum and drain the kernel queue between shells.
    """
    idx = [i for i, g in enumerate(gls) if isinstance(g, torch.Tensor) and g.is_cuda and g.numel() > 0]
    if not idx:
        return [_np(g) for g in gls]
    flat = torch.cat([gls[i].detach().reshape(-1).to(torch.float64) for i in idx]).cpu().numpy()
    out = [None] * len(gls)
    pos = 0
    for i in idx:
        n = gls[i].numel()
        out[i] = flat[pos : pos + n].reshape(tuple(gls[i].shape))
        pos += n
    for i, g in enumerate(gls):
        if out[i] is None:
            out[i] = _np(g)
    return out


def iternorm(cov: Iterable) -> Iterator:
    """
    Scaling vectors for iterative normal sampling (glass/fields.py:101-188): for every row of
    shape (..., k+1) of ``cov`` yield ``[a_1..a_k, s]`` of the same shape, so that
    ``x_i = sum_j a_j z_{i-j} + s z_i`` has the prescribed covariance with the k previous items.

    The recursion runs on the GPU (K1, ``glb_iternorm_step``: one thread per leading-dimension
    element, state resident in HBM); rows come back as NumPy arrays for NumPy rows and as CUDA
    tensors for CUDA rows.  Errors and their order as in the reference.
    """
    state = None
    lead = None
    for row in cov:
        on_device = isinstance(row, torch.Tensor) and row.is_cuda
        shape = tuple(row.shape)
        if shape[-1] == 0:
            raise ValueError("empty covariance matrix")
        if state is None:
            lead = shape[:-1]
            device = row.device if on_device else torch.device("cuda", hp._device_index())
            state = _DeviceIterNorm(int(np.prod(lead, dtype=np.int64)), shape[-1] - 1, 1, device)
        elif shape[:-1] != lead or shape[-1] - 1 != state.k:
            raise ValueError("shape mismatch in covariance")
        w = state.step(row.reshape(state.n, state.k + 1) if on_device else np.reshape(_np(row), (state.n, state.k + 1)), flag_slot=0)
        if state.failed(0):
            raise ValueError("covariance matrix is not positive definite")
        w = w.reshape(shape)
        yield w if on_device else w.cpu().numpy()


def cls2cov(cls, nl: int, nf: int, nc: int):
    """
    Rows of the banded covariance for iterative sampling (glass/fields.py:191-236): for shell j
    an (nl, nc+1) array whose column i holds half the spectrum of shells (j, j-i), zero-padded;
    columns beyond the shell's own history keep zeros.  ONE array is filled again for every shell
    and yielded each time, as the reference does with the NumPy backend -- copy it to keep it.
    """
    half = np.zeros((nl, nc + 1))
    first = 0
    for j in range(nf):
        spectra = [_np(c) for c in cls[first : first + min(j, nc) + 1]]
        first += j + 1
        if spectra and (spectra[0] < 0).any():
            msg = "negative values in cl"
            raise ValueError(msg)
        for col, c in enumerate(spectra):
            n = c.shape[0]
            np.multiply(c, 0.5, out=half[:n, col])
            half[n:, col] = 0.0
        yield half


def getcl(cls, i: int, j: int, lmax: int | None = None):
    """glass/fields.py:525-560."""
    if j > i:
        i, j = j, i
    cl = cls[i * (i + 1) // 2 + i - j]
    if lmax is not None:
        if cl.shape[0] > lmax + 1:
            cl = cl[: lmax + 1]
        elif isinstance(cl, torch.Tensor):
            cl = torch.nn.functional.pad(cl, (0, lmax + 1 - cl.shape[0]))
        else:
            cl = np.pad(cl, (0, lmax + 1 - cl.shape[0]))
    return cl


def enumerate_spectra(entries):
    """Iterate over two-point functions in standard ("Christmas tree") order, yielding
    ``(i, j, entry)`` (glass/fields.py:563-581)."""
    for k, cl in enumerate(entries):
        i = int((2 * k + 0.25) ** 0.5 - 0.5)
        j = i * (i + 3) // 2 - k
        yield i, j, cl


def spectra_indices(n: int, *, xp=None) -> np.ndarray:
    """Index pairs ``(i, j)`` in standard order for ``n`` fields, one per row
    (glass/fields.py:584-604)."""
    i, j = np.tril_indices(n)
    out = np.stack([i, i - j]).T
    return torch.as_tensor(out) if xp is torch else out


def glass_to_healpix_spectra(spectra):
    """Reorder spectra from GLASS order to (new) HEALPix order: all auto-spectra, then all
    first off-diagonals, ... (glass/fields.py:897-917)."""
    n = nfields_from_nspectra(len(spectra))
    comb = {(int(i), int(j)): pos for pos, (i, j) in enumerate(spectra_indices(n))}
    return [spectra[comb[(i + k, i)]] for k in range(n) for i in range(n - k)]


def healpix_to_glass_spectra(spectra):
    """Reorder spectra from (new) HEALPix order to GLASS order (glass/fields.py:920-940)."""
    n = nfields_from_nspectra(len(spectra))
    comb = {(i + k, i): pos for pos, (k, i) in enumerate((k, i) for k in range(n) for i in range(n - k))}
    return [spectra[comb[(int(i), int(j))]] for i, j in spectra_indices(n)]


def lognormal_shift_hilbert2011(z: float) -> float:
    """Lognormal shift of Hilbert et al. (2011) for convergence fields (glass/fields.py:965-982)."""
    return z * (8e-3 + z * (2.9e-2 + z * (-7.9e-3 + z * 6.5e-4)))


def cov_from_spectra(spectra, *, lmax: int | None = None) -> np.ndarray:
    """Covariance matrices ``cov[l, i, j]`` from spectra in standard order; ragged or empty
    spectra leave zeros (glass/fields.py:985-1032).  Host-side: (lmax+1) x n x n doubles."""
    n = nfields_from_nspectra(len(spectra))
    spectra = _gls_to_host(spectra)
    k = max((cl.shape[0] for cl in spectra), default=0) if lmax is None else lmax + 1
    cov = np.zeros((k, n, n), dtype=spectra[0].dtype if spectra else np.float64)
    for i, j, cl in enumerate_spectra(spectra):
        size = min(k, cl.shape[0])
        flat = np.reshape(cl, (-1,))
        cov[:size, i, j] = flat[:size]
        cov[:size, j, i] = flat[:size]
    return cov


def check_posdef_spectra(spectra) -> bool:
    """Whether the spectra form positive semi-definite matrices at every l
    (glass/fields.py:1035-1052)."""
    cov = cov_from_spectra(spectra)
    return bool(np.all(np.linalg.eigvalsh(cov) >= 0))


def regularized_spectra(spectra, *, lmax: int | None = None, method: str = "nearest", **method_kwargs):
    """
    Regularise a complete set of spectra so that at every l the matrix C_l^{ij} is a valid
    positive semi-definite covariance (glass/fields.py:1055-1112).  ``method``: "nearest"
    (``algorithm.cov_nearest``) or "clip" (``algorithm.cov_clip``); the (lmax+1) matrices are
    processed as one batch on the device.
    """
    from . import algorithm

    if method == "clip":
        cov_method = algorithm.cov_clip
    elif method == "nearest":
        cov_method = algorithm.cov_nearest
    else:
        msg = f"unknown method '{method}'"
        raise ValueError(msg)
    on_device = any(isinstance(cl, torch.Tensor) and cl.is_cuda for cl in spectra)
    cov = cov_from_spectra(spectra, lmax=lmax)
    if on_device:
        cov = torch.as_tensor(cov, device=next(cl.device for cl in spectra if isinstance(cl, torch.Tensor) and cl.is_cuda))
    cov = cov_method(cov, **method_kwargs)
    return [cov[:, i, j] for i, j in spectra_indices(cov.shape[-1])]


def cltovar(cl) -> float:
    """transformcl.cltovar as used at glass/fields.py:890: sum_l (2l+1)/(4 pi) C_l."""
    cl = _np(cl)
    ell = np.arange(cl.shape[0])
    return float(np.sum((2 * ell + 1) / (4 * np.pi) * cl))


def _pack_spectra(cls, lmax, device):
    """Spectra as zero-padded rows [nspec][ld] on the device (one copy each; empty spectra are zero
    rows) plus their lengths after truncation to lmax."""
    lens = [min(cl.shape[0], lmax + 1) if lmax is not None else cl.shape[0] for cl in cls]
    ld = max(max(lens, default=0), 1)
    rows = torch.zeros((len(cls), ld), dtype=torch.float64, device=device)
    for s, (cl, n) in enumerate(zip(cls, lens)):
        if n > 0:
            rows[s, :n] = torch.as_tensor(cl[:n]).to(device=device, dtype=torch.float64)
    return rows, lens


def discretized_cls(cls, *, lmax: int | None = None, ncorr: int | None = None, nside: int | None = None, pixwin=None):
    """
    Apply discretisation effects to angular power spectra (glass/fields.py:239-300): truncate
    to ``lmax``, keep ``ncorr`` correlations, multiply by the squared HEALPix pixel window.

    The window is ``hp.pixwin(nside, lmax=lmax)`` -- generated numerically here
    (glass_b200.pixwin), read from healpy's data files in the reference (glass/healpix.py:313-356)
    -- or the caller's own table ``pixwin=`` (extension).  Spectra given as CUDA tensors stay on the
    device: they are packed as rows and windowed by ONE launch (``glb_cls_window``; a real shell set
    has S(S+1)/2 = 1830 spectra), the results are row views.  NumPy spectra are handled on the host.
    """
    if len(cls) == 0:
        return []
    if ncorr is not None:
        n = nfields_from_nspectra(len(cls))
        empty = cls[0][:0]
        cls = [cls[i * (i + 1) // 2 + j] if j <= ncorr else empty for i in range(n) for j in range(i + 1)]
    pw = None
    if nside is not None:
        if pixwin is None:
            pixwin = hp.pixwin(nside, lmax=lmax)
        pw = pixwin[: lmax + 1] if lmax is not None else pixwin
    dev = next((cl.device for cl in cls if isinstance(cl, torch.Tensor) and cl.is_cuda), None)
    if dev is not None and pw is not None:
        rows, lens = _pack_spectra(cls, lmax, dev)
        w = torch.as_tensor(np.asarray(pw) if not isinstance(pw, torch.Tensor) else pw).to(device=dev, dtype=torch.float64).contiguous()
        n = min(rows.shape[1], w.shape[0])
        out = torch.empty_like(rows)
        lib = _lib.load()
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.glb_cls_window(rows.shape[0], n, rows.shape[1], out.shape[1], rows.data_ptr(), w.data_ptr(), out.data_ptr(), st), "glb_cls_window")
        return [out[s, : min(k, n)] if k > 0 else cl for s, (cl, k) in enumerate(zip(cls, lens))]
    gls = []
    for cl in cls:
        if cl.shape[0] > 0:
            if lmax is not None:
                cl = cl[: lmax + 1]
            if pw is not None:
                n = min(cl.shape[0], pw.shape[0])
                w = pw[:n]
                if isinstance(cl, torch.Tensor) and not isinstance(w, torch.Tensor):
                    w = torch.as_tensor(np.asarray(w), dtype=cl.dtype, device=cl.device)
                cl = cl[:n] * w**2
        gls.append(cl)
    return gls


def effective_cls(cls, weights1, weights2=None, *, lmax: int | None = None):
    """
    Effective angular power spectra from weights (glass/fields.py:607-694):
    ``out[j1 + j2] = sum_{i1, i2} w1[i1, j1] w2[i2, j2] C_l^{i1 i2}``, accumulated in the
    reference's order (i1 outer, i2 inner, transposed elements copied when ``weights2`` is not
    given) so that the result is bit-identical.  With CUDA tensors (spectra or weights) all output
    spectra come from ONE launch of ``glb_effective_cls`` and stay on the device; NumPy inputs -- a
    handful of short arrays -- are summed on the host.
    """
    n = nfields_from_nspectra(len(cls))
    if lmax is None:
        lmax = max((cl.shape[0] for cl in cls), default=0) - 1
    same = weights2 is None
    dev = next((a.device for a in (*cls, weights1, weights2) if isinstance(a, torch.Tensor) and a.is_cuda), None)
    if dev is None:
        return _effective_cls_host(cls, n, weights1, weights2, lmax)
    w1 = torch.as_tensor(weights1).to(device=dev, dtype=torch.float64)
    w2 = w1 if same else torch.as_tensor(weights2).to(device=dev, dtype=torch.float64)
    shape1, shape2 = tuple(w1.shape), tuple(w2.shape)
    for i, shape in enumerate((shape1, shape2)):
        if not shape or shape[0] != n:
            msg = f"shape mismatch between fields and weights{i + 1}"
            raise ValueError(msg)
    w1 = w1.reshape(n, -1).contiguous()
    w2 = w1 if same else w2.reshape(n, -1).contiguous()
    J1, J2, L = w1.shape[1], w2.shape[1], lmax + 1
    rows, _ = _pack_spectra(cls, lmax, dev)
    if rows.shape[1] < L:
        rows = torch.nn.functional.pad(rows, (0, L - rows.shape[1]))
    out = torch.empty((J1, J2, L), dtype=torch.float64, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(
            lib.glb_effective_cls(n, J1, J2, L, rows.shape[1], int(same), rows.data_ptr(), w1.data_ptr(), w2.data_ptr(), out.data_ptr(), st),
            "glb_effective_cls",
        )
    return out.reshape(shape1[1:] + shape2[1:] + (L,))


def _effective_cls_host(cls, n, weights1, weights2, lmax):
    """NumPy spectra and weights (a handful of short arrays): the same sums on the host."""
    import itertools

    cls = [_np(cl) for cl in cls]
    weights1 = _np(weights1)
    same = weights2 is None
    weights2 = weights1 if same else _np(weights2)
    shape1, shape2 = weights1.shape, weights2.shape
    for i, shape in enumerate((shape1, shape2)):
        if not shape or shape[0] != n:
            msg = f"shape mismatch between fields and weights{i + 1}"
            raise ValueError(msg)
    if same:
        pairs = itertools.combinations_with_replacement(np.ndindex(shape1[1:]), 2)
    else:
        pairs = itertools.product(np.ndindex(shape1[1:]), np.ndindex(shape2[1:]))
    out = np.empty(shape1[1:] + shape2[1:] + (lmax + 1,))
    c = (slice(None),)
    for j1, j2 in pairs:
        w1, w2 = weights1[c + j1], weights2[c + j2]
        cl = sum(w1[i1] * w2[i2] * getcl(cls, i1, i2, lmax=lmax) for i1 in range(n) for i2 in range(n))
        out[j1 + j2 + (...,)] = cl
        if same and j1 != j2:
            out[j2 + j1 + (...,)] = cl
    return out


def compute_gaussian_spectra(fields, spectra):
    """
    Band-limited Gaussian angular power spectra for the target ``spectra`` after transformation
    by ``fields`` (glass/fields.py:743-775): ``grf.compute`` per non-empty spectrum.
    """
    n = len(fields)
    if len(spectra) != n * (n + 1) // 2:
        msg = "mismatch between number of fields and spectra"
        raise ValueError(msg)
    return [grf.compute(cl, fields[i], fields[j]) if cl.shape[0] > 0 else 0 * cl for i, j, cl in enumerate_spectra(spectra)]


def _stack_transformations(ts, device):
    """One transformation object whose parameters are per-column tensors, for a group of
    transformations of the same built-in class; None for anything else."""
    cls = type(ts[0])
    if any(type(t) is not cls for t in ts) or cls not in (grf.Normal, grf.Lognormal, grf.SquaredNormal):
        return None
    if cls is grf.Normal:
        return grf.Normal()
    col = lambda name: torch.tensor([float(getattr(t, name)) for t in ts], dtype=torch.float64, device=device)  # noqa: E731
    if cls is grf.Lognormal:
        return grf.Lognormal(col("lamda"))
    return grf.SquaredNormal(col("a"), col("lamda"))


def solve_gaussian_spectra(fields, spectra):
    """
    Solve a sequence of Gaussian angular power spectra (glass/fields.py:778-836): after
    transformation by ``fields`` the two-point statistics recover ``spectra`` for a
    non-band-limited transform.  Per spectrum the reference's choices are kept -- zero padding
    ``2 n``, monopole pinned to zero when the target monopole is zero, a warning when the
    solver does not converge.

    The reference runs S(S+1)/2 independent solves one after the other on the CPU.  Here the
    spectra that share a length and a pair of built-in transformation classes become the
    columns of ONE batched Gauss-Newton run (``grf.solve_columns``: every C_l <-> C(theta)
    transform is one FP64 DGEMM over all columns, step halving and stopping are per column);
    user-defined transformations are solved one by one.
    """
    from . import transformcl as tcl

    n = len(fields)
    if len(spectra) != n * (n + 1) // 2:
        msg = "mismatch between number of fields and spectra"
        raise ValueError(msg)
    if len(spectra) == 0:
        return []
    device, _ = tcl._compute_device(*spectra)
    out: list = [None] * len(spectra)
    groups: dict[tuple, list[int]] = {}
    pairs = [(i, j) for i, j, _cl in enumerate_spectra(spectra)]
    for k, cl in enumerate(spectra):
        if cl.shape[0] == 0:
            out[k] = 0 * cl  # a copy of the empty array
            continue
        i, j = pairs[k]
        groups.setdefault((cl.shape[0], type(fields[i]), type(fields[j])), []).append(k)
    for (length, _c1, _c2), members in groups.items():
        t1 = _stack_transformations([fields[pairs[k][0]] for k in members], device)
        t2 = _stack_transformations([fields[pairs[k][1]] for k in members], device)
        if t1 is None or t2 is None:  # user-defined transformations: the reference's own loop
            for k in members:
                i, j = pairs[k]
                cl = spectra[k]
                monopole = 0.0 if float(cl[0]) == 0 else None
                gl, _cl_out, info = grf.solve(cl, fields[i], fields[j], pad=2 * length, monopole=monopole)
                if info == 0:
                    warnings.warn(f"Gaussian spectrum for fields ({i}, {j}) did not converge", stacklevel=2)
                out[k] = gl
            continue
        cols = torch.stack([_as_device_f64(spectra[k], device) for k in members], dim=1)
        mono = torch.where(cols[0] == 0, torch.zeros_like(cols[0]), torch.full_like(cols[0], float("nan")))
        gl, _rl, info = grf.solve_columns(cols, t1, t2, pad=2 * length, fix_monopole=mono)
        info = info.cpu().numpy()
        for c, k in enumerate(members):
            if info[c] == 0:
                i, j = pairs[k]
                warnings.warn(f"Gaussian spectrum for fields ({i}, {j}) did not converge", stacklevel=2)
            g = gl[:, c].contiguous()
            out[k] = g if (isinstance(spectra[k], torch.Tensor) and spectra[k].is_cuda) else g.cpu().numpy()
    return out


def _as_device_f64(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64)
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device)


@deprecated("use glass.lognormal_fields() and glass.solve_gaussian_spectra() instead")
def lognormal_gls(cls, shift: float = 1.0):
    """Gaussian spectra for lognormal fields of one shift (glass/fields.py:304-331)."""
    n = nfields_from_nspectra(len(cls))
    return solve_gaussian_spectra([grf.Lognormal(shift) for _ in range(n)], cls)


def _glass_to_healpix_alm(alm):
    """l-major -> m-major (glass/fields.py:943-962)."""
    if isinstance(alm, torch.Tensor) and alm.is_cuda:
        n = _inv_triangle_number(alm.numel())
        a = alm.to(torch.complex128).contiguous()
        out = torch.empty_like(a)
        if n:
            lib = _lib.load()
            with torch.cuda.device(a.device):
                st = torch.cuda.current_stream().cuda_stream
                _lib.check(lib.glb_alm_glass_to_healpix(n - 1, a.data_ptr(), out.data_ptr(), st), "glb_alm_glass_to_healpix")
        return out
    alm = _np(alm)
    n = _inv_triangle_number(alm.size)
    ell = np.arange(n)
    out = [alm[ell[m:] * (ell[m:] + 1) // 2 + m] for m in ell]
    return np.concatenate(out) if out else alm


# --------------------------------------------------------------------------------------
# the generator chain
# --------------------------------------------------------------------------------------


class _DeviceIterNorm:
    """glass/fields.py:101-188 with the state (m, a, s) resident on the GPU (K1,
    ``glb_iternorm_step``, one thread per multipole); ``step`` consumes one row of ``cls2cov``
    and returns the weights ``[a, s]`` as a device tensor (n, k+1).

    The recursion depends on nothing but the spectra, so it runs on its OWN high-priority stream:
    reading its flags ("covariance matrix is not positive definite", one per step) waits for a
    few microsecond-sized kernels and never for the transforms queued on the caller's stream.
    The caller's stream waits for the event recorded after each step before it reads the weights.
    """

    def __init__(self, n: int, k: int, nsteps: int, device):
        self.lib = _lib.load()
        self.n, self.k, self.device = int(n), int(k), device
        kk = max(self.k, 1)
        self.stream = torch.cuda.Stream(device, priority=-1)
        with torch.cuda.stream(self.stream):
            self.m = torch.zeros((kk, kk, self.n), dtype=torch.float64, device=device)
            self.a = torch.zeros((kk, self.n), dtype=torch.float64, device=device)
            self.s = torch.ones(self.n, dtype=torch.float64, device=device)
            self.tmp = torch.empty((kk, self.n), dtype=torch.float64, device=device)
            self.flags = torch.zeros(max(int(nsteps), 1), dtype=torch.int32, device=device)
        self.i = 0
        self.checked = 0

    def step(self, row, flag_slot: int | None = None) -> torch.Tensor:
        """One shell.  ``row``: (n, k+1) NumPy array or CUDA tensor.  The flag of this step goes
        to ``flags[flag_slot]`` (default: the step index)."""
        if tuple(row.shape) != (self.n, self.k + 1):
            raise ValueError("shape mismatch in covariance")
        dev = self.device
        slot = self.i if flag_slot is None else flag_slot
        caller = torch.cuda.current_stream(dev)
        with torch.cuda.stream(self.stream):
            if isinstance(row, torch.Tensor):
                self.stream.wait_stream(caller)
                rd = row.to(device=dev, dtype=torch.float64).contiguous()
            else:
                # pinned copy of the row: cls2cov re-yields one buffer, and a pageable H2D would block
                rd = torch.from_numpy(np.array(row, dtype=np.float64, order="C")).pin_memory().to(dev, non_blocking=True)
            if flag_slot is not None:
                self.flags[slot].zero_()
            w = torch.empty((self.n, self.k + 1), dtype=torch.float64, device=dev)
            _lib.check(
                self.lib.glb_iternorm_step(self.n, self.k, 1 if self.i == 0 else 0, rd.data_ptr(), self.m.data_ptr(), self.a.data_ptr(),
                                           self.s.data_ptr(), self.tmp.data_ptr(), w.data_ptr(), self.flags[slot:].data_ptr(),
                                           self.stream.cuda_stream),
                "glb_iternorm_step",
            )
        caller.wait_stream(self.stream)
        w.record_stream(caller)
        self.i += 1
        return w

    def failed(self, slot: int) -> bool:
        with torch.cuda.stream(self.stream):
            return bool(self.flags[slot].item())

    def first_failure(self):
        """Index of the first step whose covariance was not positive definite, or None
        (waits for the recursion's own stream only)."""
        if self.i == self.checked:
            return None
        with torch.cuda.stream(self.stream):
            f = self.flags[self.checked : self.i].cpu().numpy()
        bad = np.nonzero(f)[0]
        if bad.size:
            return self.checked + int(bad[0])
        self.checked = self.i
        return None


class _ShellSampler:
    """Device-side state of _generate_grf: z history, iternorm weights, alm of each shell.

    ``wanted`` (optional predicate on the shell index) implements shell sharding across
    GPUs: the iternorm recursion (host, tiny) still walks every shell, but normals are
    only drawn for the shells a wanted shell is correlated with -- they are a pure
    function of (seed, shell, index), so every rank regenerates the neighbours it needs
    and no data is exchanged.
    """

    def __init__(self, gls, nside, ncorr, rng, device, wanted=None):
        self.lib = _lib.load()
        self.device = device
        self.nside = nside
        gls = _gls_to_host(gls)
        ngrf = nfields_from_nspectra(len(gls))
        self.ngrf = ngrf
        self.ncorr = ngrf - 1 if ncorr is None else ncorr
        self.n = max((gl.shape[0] for gl in gls), default=0)
        if self.n == 0:
            raise ValueError("all gls are empty")
        self.lmax = self.n - 1
        self.nalm = self.n * (self.n + 1) // 2
        self.deviates = rng if isinstance(rng, _rng.Deviates) else None
        self.seed = _rng.seed_from(rng)
        self.dnorm = _DeviceIterNorm(self.n, self.ncorr, ngrf, device)
        self.witer = cls2cov(gls, self.n, ngrf, self.ncorr)  # rows; the recursion itself is K1
        self.wanted = wanted
        self.zcache: dict[int, torch.Tensor] = {}
        self.shell = 0
        self.h2d_bytes = 0

    def _z(self, j: int) -> torch.Tensor:
        z = self.zcache.get(j)
        if z is None:
            lib, dev = self.lib, self.device
            st = torch.cuda.current_stream(dev).cuda_stream
            z = torch.empty(self.nalm, dtype=torch.complex128, device=dev)
            if self.deviates is not None:
                zh = np.ascontiguousarray(self.deviates.normal_alm[j], dtype=np.complex128)
                zg = torch.as_tensor(zh).to(dev)
                self.h2d_bytes += zh.nbytes
                _lib.check(lib.glb_alm_glass_to_healpix(self.lmax, zg.data_ptr(), z.data_ptr(), st), "glb_alm_glass_to_healpix")
            else:
                _lib.check(lib.glb_alm_draw(self.lmax, C.c_uint64(self.seed), C.c_uint32(j), z.data_ptr(), st), "glb_alm_draw")
            self.zcache[j] = z
        return z

    def next_alm(self, out: torch.Tensor):
        """Fill ``out`` (nalm complex128, m-major) with the next wanted shell's alm
        (glass/fields.py:404-425); returns its shell index, or None when exhausted."""
        while True:
            try:
                w = next(self.witer)
            except StopIteration:
                return None
            self.h2d_bytes += w.nbytes
            w = self.dnorm.step(w)  # device tensor (n, ncorr + 1)
            j = self.shell
            self.shell += 1
            if self.wanted is None or self.wanted(j):
                break
        lib, dev = self.lib, self.device
        st = torch.cuda.current_stream(dev).cuda_stream
        nterms = min(j + 1, w.shape[-1])  # len(y) after the deque trimming of fields.py:410-414
        mis = w.shape[-1] - nterms
        zs = [self._z(s) for s in range(j - nterms + 1, j + 1)]
        for s in [s for s in self.zcache if s < j - self.ncorr]:
            del self.zcache[s]
        stride, wptr = w.shape[-1], w.data_ptr() + 8 * mis
        zptrs = (C.c_void_p * nterms)(*[t.data_ptr() for t in zs])
        _lib.check(
            lib.glb_alm_combine(self.lmax, nterms, zptrs, wptr, stride, out.data_ptr(), st),
            "glb_alm_combine",
        )
        return j

    def first_failure(self):
        """Shell index at which the recursion found a non positive definite covariance (None if
        none so far)."""
        return self.dnorm.first_failure()


@contextlib.contextmanager
def _nvtx_range(name: str):
    """NVTX range (timelines show the stages by name); nothing without a device."""
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


def _pick_device(gls) -> tuple[torch.device, bool]:
    """(device, on_device): the device of the first CUDA spectrum, else the current CUDA device
    (raises without one); on_device decides the array rule of the module docstring."""
    for gl in gls:
        if isinstance(gl, torch.Tensor) and gl.is_cuda:
            return gl.device, True
    return torch.device("cuda", hp._device_index()), False


def _generate_maps(gls, nside, ncorr, rng, transforms_for, shells=None, stats=None):
    """
    Core of _generate_grf / generate.  ``transforms_for(i)`` returns the fused
    (kind, p0, p1) descriptor for shell i or None for "no fused transform".
    Yields (i, map) with map a CUDA tensor or a NumPy array (array rule above).
    """
    device, on_device = _pick_device(gls)
    wanted = None
    if shells is not None:
        wanted = shells if callable(shells) else (lambda j, _s=frozenset(int(i) for i in shells): j in _s)
    with torch.cuda.device(device):
        sampler = _ShellSampler(gls, nside, ncorr, rng, device, wanted)
        # eight maps at once need eight phase maps and eight output maps in HBM: 30 GB at nside 4096, 120 GB at 8192
        B = max(1, min(int(SHT_BATCH), 8 if nside <= 4096 else 4))
        npix = hp.nside2npix(nside)
        copy_stream = None if on_device else torch.cuda.Stream(device)
        state = {"error": None}

        main = torch.cuda.current_stream(device)
        # the next batch's a_lm, records and digit planes are made here (a high-priority stream moves the same few ms from
        # the ring FFT, which the look-ahead otherwise overlaps with, into the Legendre kernel: no difference in the step)
        side = torch.cuda.Stream(device)
        lib = _lib.load()
        pipe = {"split": B == 8, "slot": 0, "freed": [None, None], "last": False}

        def start():
            """Draw and combine the next batch (<= B shells) on the side stream and, for a full batch of
            eight on the INT8 Legendre path, prepare its transform there as well (glb_alm2map_prepare)
            while the caller's stream is busy with the previous batch.  None when there is nothing left."""
            if pipe["last"] or state["error"] is not None:
                return None
            side.wait_stream(main)  # (order after what the caller queued before the call)
            with torch.cuda.stream(side), _nvtx_range("glass.generate: draw, combine, prepare (look-ahead batch)"):
                alms = torch.empty((B, sampler.nalm), dtype=torch.complex128, device=device)
                idx = []
                while len(idx) < B:
                    try:
                        j = sampler.next_alm(alms[len(idx)])
                    except ValueError as e:  # raise only once the earlier shells were yielded
                        state["error"] = e
                        break
                    if j is None:
                        pipe["last"] = True
                        break
                    idx.append(j)
                bad = sampler.first_failure()
                if bad is not None:
                    # the reference raises when it reaches shell `bad`: yield the earlier ones first
                    idx = [j for j in idx if j < bad]
                    state["error"] = ValueError("covariance matrix is not positive definite")
                    pipe["last"] = True
                if not idx:
                    return None
                rec = {"alms": alms, "idx": idx, "slot": None}
                if pipe["split"] and len(idx) == 8:
                    slot = pipe["slot"]
                    if pipe["freed"][slot] is not None:
                        side.wait_event(pipe["freed"][slot])  # the Legendre kernel that read this set of tile blocks
                    pl = hp.get_plan(nside, sampler.lmax, max_batch=4, device=device)
                    rc = lib.glb_alm2map_prepare(pl.handle, alms.data_ptr(), 8, slot, side.cuda_stream)
                    if rc == _lib.GLB_ERR_UNSUPPORTED:
                        pipe["split"] = False  # not a size the INT8 path takes: the plain call below
                    else:
                        _lib.check(rc, "glb_alm2map_prepare")
                        rec["slot"], rec["plan"] = slot, pl
                        pipe["slot"] = slot ^ 1
                rec["ready"] = torch.cuda.Event()
                rec["ready"].record(side)
            return rec

        def batches():
            """One batch per iteration: the look-ahead batch is started (side stream) before the current
            one is synthesised on the caller's stream -- batched transform with fused pixel
            transformations and, in host mode, the async D2H copies."""
            cur = start()
            while cur is not None:
                nxt = start()
                alms, idx = cur["alms"], cur["idx"]
                nb = len(idx)
                main.wait_event(cur["ready"])
                alms.record_stream(main)
                tr = [transforms_for(j) or (_lib.T_NORMAL, 0.0, 1.0) for j in idx]
                with _nvtx_range("glass.generate: alm2map of a batch"):
                    if cur["slot"] is not None:
                        pl = cur["plan"]
                        maps = torch.empty((nb, npix), dtype=torch.float64, device=device)
                        kinds, params, _keep = hp._transform_args(tr)
                        _lib.check(lib.glb_alm2map_finish(pl.handle, 8, cur["slot"], maps.data_ptr(), kinds, params, main.cuda_stream), "glb_alm2map_finish")
                        done_leg = torch.cuda.Event()
                        done_leg.record(main)
                        pipe["freed"][cur["slot"]] = done_leg
                    else:
                        maps = hp.alm2map_batch(alms[:nb], nside, sampler.lmax, transforms=tr)
                if nxt is not None:
                    # the look-ahead batch used the plan's record buffer on the side stream: whatever the consumer runs on
                    # this stream next (other transforms of the same plan) comes after it
                    main.wait_event(nxt["ready"])
                if stats is not None:
                    stats["h2d_bytes"] = sampler.h2d_bytes
                if on_device:
                    yield [(idx[b], maps[b], None) for b in range(nb)]
                    cur = nxt
                    continue
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(device))
                out = []
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    for b in range(nb):
                        h = _PINNED.take(npix)
                        h.copy_(maps[b], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                        out.append((idx[b], h, ev))
                maps.record_stream(copy_stream)
                yield out
                cur = nxt

        def drain(batch):
            for j, m, ev in batch:
                if ev is None:
                    yield j, m
                else:
                    ev.synchronize()
                    if stats is not None:
                        stats["d2h_bytes"] = stats.get("d2h_bytes", 0) + m.numel() * 8
                    yield j, _PINNED.lend(m)

        if on_device:
            for batch in batches():
                yield from drain(batch)
        else:
            # one batch of look-ahead: the next batch's kernels are queued before the
            # previous batch's device->host copies are waited for, so they overlap
            prev = None
            for batch in batches():
                if prev is not None:
                    yield from drain(prev)
                prev = batch
            if prev is not None:
                yield from drain(prev)
        if state["error"] is not None:
            raise state["error"]


def _generate_grf(gls, nside: int, *, ncorr: int | None = None, rng=None):
    """Iteratively sample Gaussian random fields (glass/fields.py:334-429)."""
    for _i, m in _generate_maps(gls, nside, ncorr, rng, lambda i: None):
        yield m


def generate(fields: Sequence, gls, nside: int, *, ncorr: int | None = None, rng=None, shells=None, stats=None) -> Iterator:
    """
    Sample random fields from Gaussian angular power spectra (glass/fields.py:839-894).

    Extension over the reference signature (keyword-only, default = reference behaviour):
    ``shells`` -- iterable of shell indices (or predicate) to produce; the others are
    skipped.  This is the multi-GPU sharding of the path: rank r of W passes
    ``shells=range(r, n, W)`` and no data is exchanged between ranks.
    """
    n = len(fields)
    if len(gls) != n * (n + 1) // 2:
        msg = "mismatch between number of fields and gls"
        raise ValueError(msg)

    variances: dict[int, float] = {}
    host_autos: dict[int, np.ndarray] = {}
    if any(isinstance(g, torch.Tensor) and g.is_cuda for g in gls):
        # one transfer for all auto-spectra instead of a stream synchronisation per shell
        ii = [i * (i + 1) // 2 for i in range(n)]  # getcl(gls, i, i)
        host_autos = dict(zip(range(n), _gls_to_host([gls[k] for k in ii])))

    def var_of(i: int) -> float:
        if i not in variances:
            variances[i] = cltovar(host_autos[i] if i in host_autos else getcl(gls, i, i))
        return variances[i]

    def transforms_for(i: int):
        if i >= n:
            return None
        return grf.fused_descriptor(fields[i], var_of(i))

    for i, x in _generate_maps(gls, nside, ncorr, rng, transforms_for, shells, stats):
        if i >= n:
            break
        t = fields[i]
        if grf.fused_descriptor(t, 0.0) is None:
            x = t(x, var_of(i))  # user-defined transformation: separate pass
        yield x


@deprecated("use glass.generate() instead")
def generate_gaussian(gls, nside: int, *, ncorr: int | None = None, rng=None):
    """glass/fields.py:432-483."""
    n = nfields_from_nspectra(len(gls))
    fields = [grf.Normal() for _ in range(n)]
    yield from generate(fields, gls, nside, ncorr=ncorr, rng=rng)


@deprecated("use glass.generate() instead")
def generate_lognormal(gls, nside: int, shift: float = 1.0, *, ncorr: int | None = None, rng=None):
    """glass/fields.py:485-522."""
    n = nfields_from_nspectra(len(gls))
    fields = [grf.Lognormal(shift) for _ in range(n)]
    yield from generate(fields, gls, nside, ncorr=ncorr, rng=rng)


def gaussian_fields(shells: Sequence) -> Sequence[grf.Normal]:
    """glass/fields.py:697-713."""
    return [grf.Normal() for _shell in shells]


def lognormal_fields(shells: Sequence, shift: Callable[[float], float] | None = None) -> Sequence[grf.Lognormal]:
    """glass/fields.py:716-740."""
    if shift is None:
        shift = lambda _z: 1.0  # noqa: E731
    return [grf.Lognormal(shift(shell.zeff)) for shell in shells]
