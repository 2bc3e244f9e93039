"""Shared test fixtures: the reference's MockCosmology (tests/fixtures/domain.py:36-97) and
its 5-shell triangular windows (tests/fixtures/domain.py:116-144), restated."""
import numpy as np


class MockCosmology:
    Omega_m0 = 0.3
    critical_density0 = 3e4
    hubble_distance = 4.4e3

    def H_over_H0(self, z):  # noqa: N802
        return (self.Omega_m0 * (1 + z) ** 3 + 1 - self.Omega_m0) ** 0.5

    def xm(self, z, z2=None):
        if z2 is None:
            return np.asarray(z) * 1_000
        return (np.asarray(z2) - np.asarray(z)) * 1_000

    def transverse_comoving_distance(self, z, z2=None):
        return self.hubble_distance * self.xm(z, z2)


def triangular_shells(n=5, dz=1.0):
    from glass_b200 import RadialWindow

    return [RadialWindow(np.array([i, i + 1.0, i + 2.0]) * dz, np.array([0.0, 1.0, 0.0]), (i + 1.0) * dz) for i in range(n)]


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
