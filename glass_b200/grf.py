"""
glass_b200.grf -- the transformations of ``glass/grf/_transformations.py`` and the spectra
side of ``glass.grf`` (``corr`` / ``icorr`` / ``dcorr`` / ``compute`` / ``solve``,
glass/grf/_core.py:63-179, glass/grf/_solver.py:27-148).

``__call__`` works on NumPy arrays and torch tensors; in ``glass_b200.generate`` these
transformations are recognised and fused into the ring-FFT epilogue of the synthesis kernel
instead of being applied as a separate pass.

The solver is batched: all spectra of a simulation that share a length
and a transformation pair are the columns of one matrix, every C_l <-> C(theta) transform of
the Gauss-Newton iteration is one FP64 DGEMM against the tables of
:mod:`glass_b200.transformcl`, and step halving / convergence are tracked per column.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


def _expm1(x):
    return torch.expm1(x) if isinstance(x, torch.Tensor) else np.expm1(x)


def _xp(x):
    return torch if isinstance(x, torch.Tensor) else np


@dataclass
class Normal:
    """glass/grf/_transformations.py:14-27: t(X) = X."""

    def __call__(self, x, _var, /):
        return x

    def corr(self, other, x, /):
        """glass/grf/_transformations.py:27-35."""
        return x if type(other).__name__ == "Normal" else NotImplemented

    def icorr(self, other, x, /):
        return x if type(other).__name__ == "Normal" else NotImplemented

    def dcorr(self, other, x, /):
        return 1.0 + (0 * x) if type(other).__name__ == "Normal" else NotImplemented

    def _fused(self, var):
        return (_lib.T_NORMAL, 0.0, 1.0)


@dataclass
class Lognormal:
    r"""glass/grf/_transformations.py:60-89: t(X) = lamda [exp(X - var/2) - 1]."""

    lamda: float = 1.0

    def __call__(self, x, var, /):
        x = _expm1(x - var / 2)
        if self.lamda != 1.0:
            x = self.lamda * x
        return x

    def corr(self, other, x, /):
        """glass/grf/_transformations.py:86-100."""
        name = type(other).__name__
        if name == "Lognormal":
            return self.lamda * other.lamda * _xp(x).expm1(x)
        if name == "Normal":
            return self.lamda * x
        return NotImplemented

    def icorr(self, other, x, /):
        name = type(other).__name__
        if name == "Lognormal":
            return _xp(x).log1p(x / (self.lamda * other.lamda))
        if name == "Normal":
            return x / self.lamda
        return NotImplemented

    def dcorr(self, other, x, /):
        name = type(other).__name__
        if name == "Lognormal":
            return self.lamda * other.lamda * _xp(x).exp(x)
        if name == "Normal":
            return self.lamda + (0.0 * x)
        return NotImplemented

    def _fused(self, var):
        return (_lib.T_LOGNORMAL, float(var) / 2, float(self.lamda))


@dataclass
class SquaredNormal:
    r"""glass/grf/_transformations.py:141-176: t(X) = lamda [(X - a)^2 - 1]."""

    a: float
    lamda: float = 1.0

    def __call__(self, x, _var, /):
        x = (x - self.a) ** 2 - 1
        if self.lamda != 1.0:
            x = self.lamda * x
        return x

    def corr(self, other, x, /):
        """glass/grf/_transformations.py:170-181."""
        if type(other).__name__ == "SquaredNormal":
            aa, ll = self.a * other.a, self.lamda * other.lamda
            return 2 * ll * x * (x + 2 * aa)
        return NotImplemented

    def icorr(self, other, x, /):
        if type(other).__name__ == "SquaredNormal":
            aa, ll = self.a * other.a, self.lamda * other.lamda
            return _xp(x).sqrt(x / (2 * ll) + aa**2) - aa
        return NotImplemented

    def dcorr(self, other, x, /):
        if type(other).__name__ == "SquaredNormal":
            aa, ll = self.a * other.a, self.lamda * other.lamda
            return 4 * ll * (x + aa)
        return NotImplemented

    def _fused(self, var):
        return (_lib.T_SQUARED_NORMAL, float(self.a), float(self.lamda))


def fused_descriptor(t, var):
    """(kind, p0, p1) if ``t`` is one of the known transformations (ours or the
    reference's own dataclasses of the same name), else None."""
    if hasattr(t, "_fused"):
        return t._fused(var)
    name = type(t).__name__
    if name == "Normal" and not hasattr(t, "lamda"):
        return (_lib.T_NORMAL, 0.0, 1.0)
    if name == "Lognormal" and hasattr(t, "lamda"):
        return (_lib.T_LOGNORMAL, float(var) / 2, float(t.lamda))
    if name == "SquaredNormal" and hasattr(t, "lamda") and hasattr(t, "a"):
        return (_lib.T_SQUARED_NORMAL, float(t.a), float(t.lamda))
    return None


# --------------------------------------------------------------------------------------
# correlation-function transforms, compute, solve  (glass/grf/_core.py, glass/grf/_solver.py)
# --------------------------------------------------------------------------------------


def _dispatch(method: str, t1, t2, x):
    for a, b in ((t1, t2), (t2, t1)):
        fn = getattr(a, method, None)
        if fn is not None:
            result = fn(b, x)
            if result is not NotImplemented:
                return result
    msg = f"{t1.__class__.__name__} x {t2.__class__.__name__}"
    raise NotImplementedError(msg)


def corr(t1, t2, x, /):
    """Transform a Gaussian angular correlation function (glass/grf/_core.py:63-86)."""
    return _dispatch("corr", t1, t2, x)


def icorr(t1, t2, x, /):
    """Inverse of :func:`corr` (glass/grf/_core.py:89-112)."""
    return _dispatch("icorr", t1, t2, x)


def dcorr(t1, t2, x, /):
    """Derivative of :func:`corr` (glass/grf/_core.py:115-138)."""
    return _dispatch("dcorr", t1, t2, x)


def compute(cl, t1, t2=None):
    """Band-limited Gaussian angular power spectrum for the target ``cl`` and a pair of
    transformations (glass/grf/_core.py:141-179): ``corrtocl(icorr(cltocorr(cl)))``."""
    from . import transformcl as tcl

    if t2 is None:
        t2 = t1
    # one device pipeline (the same operations, hence the same bits, as the solver's initial guess)
    return tcl._wrap(lambda c: tcl.corrtocl_dev(icorr(t1, t2, tcl.cltocorr_dev(c))), cl)


def _take(t, idx):
    """The transformation ``t`` restricted to the columns ``idx``: stacked transformations
    (``fields._stack_transformations``) carry their parameters as per-column tensors [S],
    which must be subset together with the data columns; scalar parameters pass through."""
    pars = {k: v for k, v in vars(t).items() if isinstance(v, torch.Tensor) and v.ndim == 1}
    if not pars:
        return t
    return type(t)(**{**vars(t), **{k: v[idx] for k, v in pars.items()}})


def _relerr_cols(dx: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """max |dx / x| per column, entries with dx == 0 ignored (glass/grf/_solver.py:21-24)."""
    q = torch.where(dx != 0, dx / x, torch.zeros_like(dx))
    return q.abs().amax(dim=0) if q.shape[0] else q.new_zeros(q.shape[1])


def solve_columns(cl: torch.Tensor, t1, t2, *, pad: int, initial=None, cltol=1e-5, gltol=1e-5, maxiter=20, fix_monopole=None):
    """
    The Gauss-Newton solver of glass/grf/_solver.py:27-148 for MANY spectra at once.

    ``cl``: device tensor [n, S] (S target spectra of one length, one transformation pair);
    ``fix_monopole``: None, or a tensor [S] with NaN where the monopole is free and the value it
    is pinned to elsewhere.  Returns ``(gl [n, S], rl [n + pad, S], info [S] int32)`` with the
    reference's meaning; every column follows exactly the iteration the reference would run on
    it alone (its own step halving, its own stopping point) -- converged columns are frozen
    while the others go on.
    """
    from . import transformcl as tcl

    n, S = cl.shape
    if pad < 0:
        msg = "pad must be a positive integer"
        raise ValueError(msg)
    fixed = None if fix_monopole is None else ~torch.isnan(fix_monopole)

    def padded(x):
        return torch.nn.functional.pad(x, (0, 0, 0, pad)) if pad else x

    def pin0(x):  # residuals / steps ignore a pinned monopole
        if fixed is not None and n:
            x[0] = torch.where(fixed[: x.shape[1]] if x.shape[1] == S else fixed, torch.zeros_like(x[0]), x[0])
        return x

    if initial is None:
        gl = tcl.corrtocl_dev(icorr(t1, t2, tcl.cltocorr_dev(cl)))
    else:
        gl = torch.zeros_like(cl)
        k = min(n, initial.shape[0])
        gl[:k] = initial[:k]
    if fixed is not None and n:
        gl[0] = torch.where(fixed, fix_monopole, gl[0])

    gt = tcl.cltocorr_dev(padded(gl))
    rl = tcl.corrtocl_dev(corr(t1, t2, gt))
    fl = pin0(rl[:n] - cl)
    clerr = _relerr_cols(fl, cl)
    info = torch.zeros(S, dtype=torch.int32, device=cl.device)

    for _ in range(maxiter):
        # the reference tests clerr at the top of every iteration, also for a column whose step
        # already converged (info can end as 3); frozen columns keep their clerr
        running = info == 0
        info = torch.where(clerr <= cltol, info | 1, info)
        act = torch.nonzero(running & (info == 0)).flatten()  # columns still iterating
        if act.numel() == 0:
            break
        sub = (lambda x: x) if act.numel() == S else (lambda x: x[:, act])
        cl_a, gl_a, gt_a, fl_a, err_a = sub(cl), sub(gl), sub(gt), sub(fl), clerr[act]
        fixed_a = None if fixed is None else fixed[act]
        t1_a, t2_a = (t1, t2) if act.numel() == S else (_take(t1, act), _take(t2, act))
        ft = tcl.cltocorr_dev(padded(fl_a))
        xl = -tcl.corrtocl_dev(ft / dcorr(t1_a, t2_a, gt_a))[:n]
        if fixed_a is not None and n:
            xl[0] = torch.where(fixed_a, torch.zeros_like(xl[0]), xl[0])
        # halve the step of every column whose residual did not improve, until all did
        new_gl, new_gt, new_rl, new_fl, new_err = (torch.empty_like(v) for v in (gl_a, gt_a, sub(rl), fl_a, err_a))
        todo = torch.arange(act.numel(), device=cl.device)
        while todo.numel():
            g_ = gl_a[:, todo] + xl[:, todo]
            gt_ = tcl.cltocorr_dev(padded(g_))
            whole = todo.numel() == act.numel()
            rl_ = tcl.corrtocl_dev(corr(t1_a if whole else _take(t1_a, todo), t2_a if whole else _take(t2_a, todo), gt_))
            fl_ = rl_[:n] - cl_a[:, todo]
            if fixed_a is not None and n:
                fl_[0] = torch.where(fixed_a[todo], torch.zeros_like(fl_[0]), fl_[0])
            e_ = _relerr_cols(fl_, cl_a[:, todo])
            ok = e_ <= err_a[todo]
            done = todo[ok]
            new_gl[:, done], new_gt[:, done], new_rl[:, done], new_fl[:, done], new_err[done] = g_[:, ok], gt_[:, ok], rl_[:, ok], fl_[:, ok], e_[ok]
            todo = todo[~ok]
            xl[:, todo] /= 2
        small = _relerr_cols(xl, gl_a) <= gltol
        info[act] = torch.where(small, info[act] | 2, info[act])
        gl[:, act], gt[:, act], rl[:, act], fl[:, act], clerr[act] = new_gl, new_gt, new_rl, new_fl, new_err
    return gl, rl, info


def solve(cl, t1, t2=None, *, pad: int = 0, initial=None, cltol: float = 1e-5, gltol: float = 1e-5, maxiter: int = 20, monopole: float | None = None):
    """
    Solve for a Gaussian angular power spectrum (glass/grf/_solver.py:27-148): returns
    ``(gl, cl_out, info)``, info bit 0 = converged in cl, bit 1 = converged in gl, 0 = not
    converged in ``maxiter`` iterations.  One spectrum; see :func:`solve_columns` for the
    batched form that ``solve_gaussian_spectra`` uses.
    """
    from . import transformcl as tcl

    if t2 is None:
        t2 = t1
    if pad < 0:
        msg = "pad must be a positive integer"
        raise ValueError(msg)
    device, on_device = tcl._compute_device(cl)
    to = lambda a: (a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))).to(device=device, dtype=torch.float64)  # noqa: E731
    cld = to(cl).reshape(-1, 1)
    init = None if initial is None else to(initial).reshape(-1, 1)
    mono = None if monopole is None else torch.full((1,), float(monopole), dtype=torch.float64, device=device)
    gl, rl, info = solve_columns(cld, t1, t2, pad=pad, initial=init, cltol=cltol, gltol=gltol, maxiter=maxiter, fix_monopole=mono)
    gl, rl = gl[:, 0], rl[:, 0]
    if not on_device:
        gl, rl = gl.cpu().numpy(), rl.cpu().numpy()
    return gl, rl, int(info[0].item())
