"""
glass_b200.algorithm -- the covariance regularisation of ``glass/algorithm.py:111-277``
(``cov_clip``, ``nearcorr``, ``cov_nearest``) behind ``regularized_spectra``
(glass/fields.py:1055-1112): at every multipole the n x n matrix of spectra is made a valid
covariance.  The reference calls LAPACK ``eigh`` on the stack of (lmax+1) matrices on the CPU;
here the stack is one batched ``torch.linalg.eigh`` (cuSOLVER) and batched matmuls on the
device -- library calls, like the DGEMMs of :mod:`glass_b200.transformcl`; up to 100 alternating
projections of Higham's algorithm for ``nearest``.  NumPy in -> NumPy out, CUDA tensors stay.
"""

from __future__ import annotations

import warnings

import numpy as np
import torch

from . import transformcl as _tcl


def _to_dev(a):
    device, on_device = _tcl._compute_device(a)
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64), on_device
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(device), on_device


def _out(t: torch.Tensor, on_device: bool):
    return t if on_device else t.cpu().numpy()


def _cov_clip(cov: torch.Tensor, rtol: float | None) -> torch.Tensor:
    w, v = torch.linalg.eigh(cov)
    if rtol is None:
        rtol = max(v.shape[-2], v.shape[-1]) * torch.finfo(w.dtype).eps
    w = torch.maximum(w, rtol * w.amax(dim=-1, keepdim=True))
    v = torch.sqrt(w[..., None, :]) * v
    return v @ v.mT


def cov_clip(cov, rtol: float | None = None):
    """Covariance matrix from clipping non-positive eigenvalues (glass/algorithm.py:111-149):
    eigenvalues below ``rtol`` times the largest one are raised to that value."""
    c, on_device = _to_dev(cov)
    return _out(_cov_clip(c, rtol), on_device)


def _nearcorr(a: torch.Tensor, tol: float | None, niter: int) -> torch.Tensor:
    *dim, m, n = a.shape
    if m != n:
        msg = "non-square matrix"
        raise ValueError(msg)
    if tol is None:
        tol = n * torch.finfo(a.dtype).eps
    frob = torch.linalg.matrix_norm
    y = a.reshape(-1, n, n)
    ds = torch.zeros_like(y)
    diag = torch.eye(n, dtype=a.dtype, device=a.device)
    for _ in range(niter):
        r = y - ds
        x = _cov_clip(r, None)
        ds = x - r
        y = (1 - diag) * x + diag
        if bool(torch.all(frob(y - x) <= tol * frob(y))):
            break
    else:
        warnings.warn(
            f"Nearest correlation matrix not found in {niter} iterations. "
            "The result may be invalid. Please run with a larger `niter` value, "
            "or run the function again on the returned result.",
            stacklevel=3,
        )
    return y.reshape(*dim, n, n)


def nearcorr(a, *, tol: float | None = None, niter: int = 100):
    """Nearest correlation matrix by Higham's alternating projections
    (glass/algorithm.py:152-231), batched over the leading axes."""
    t, on_device = _to_dev(a)
    return _out(_nearcorr(t, tol, niter), on_device)


def cov_nearest(cov, tol: float | None = None, niter: int = 100):
    """Nearest covariance matrix: normalise to a correlation matrix, :func:`nearcorr`, scale
    back (glass/algorithm.py:234-277)."""
    c, on_device = _to_dev(cov)
    d = torch.diagonal(c, dim1=-2, dim2=-1)
    if bool(torch.any(d < 0)):
        msg = "negative values on the diagonal"
        raise ValueError(msg)
    norm = torch.sqrt(d)
    norm = norm[..., None, :] * norm[..., :, None]
    corr = c / torch.where(norm > 0, norm, torch.ones_like(norm))
    return _out(_nearcorr(corr, tol, niter) * norm, on_device)
