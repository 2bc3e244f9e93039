"""
oracle/transformcl_ref.py -- CPU restatement of the third-party ``transformcl`` interface that
the reference's spectra solver calls (glass/grf/_solver.py:11,100-130; glass/grf/_core.py:179).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

transformcl / flt are un-vendored and absent here (``pyproject.toml``: transformcl>=2026.1, no lock
file), so this follows their published definition -- correlation function on the open nodes
theta_k = pi (k + 1/2) / n, ``corrtocl`` the exact inverse of ``cltocorr`` on those nodes -- with
constructions that share nothing with the product's (glass_b200/transformcl.py builds P by a
recurrence and P^-1 = L.D in closed form): here the forward direction is NumPy's Clenshaw
evaluation of the Legendre series and the inverse a dense solve against SciPy's Legendre
polynomials.  O(n^3): for the small sizes of the tests only.  Parity with transformcl itself:
unpinned; it is used (a) as the checker of the product's transform pair and (b) as the
``transformcl`` module under which tests/golden/make_golden.py executes the reference's OWN solver
source, so that the solver's iteration logic is pinned by the reference.
"""

from __future__ import annotations

import numpy as np
from scipy.special import eval_legendre


def theta(n: int) -> np.ndarray:
    return (np.arange(n) + 0.5) * (np.pi / n)


def _factors(n: int) -> np.ndarray:
    return (2 * np.arange(n) + 1) / (4 * np.pi)


def cltocorr(cl, closed: bool = False) -> np.ndarray:
    assert not closed
    cl = np.asarray(cl, dtype=float)
    n = cl.shape[0]
    if n == 0:
        return cl.copy()
    return np.polynomial.legendre.legval(np.cos(theta(n)), _factors(n) * cl)


def corrtocl(corr, closed: bool = False) -> np.ndarray:
    assert not closed
    corr = np.asarray(corr, dtype=float)
    n = corr.shape[0]
    if n == 0:
        return corr.copy()
    x = np.cos(theta(n))
    P = np.stack([eval_legendre(l, x) for l in range(n)], axis=1)
    return np.linalg.solve(P, corr) / _factors(n)


def cltovar(cl) -> float:
    cl = np.asarray(cl, dtype=float)
    return float(np.sum(_factors(cl.shape[0]) * cl))
