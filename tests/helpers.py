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


def native_iternorm():
    """ctypes handle of tests/native/iternorm_host.cpp: the per-multipole body of K1
    (glass_b200/csrc/iternorm_core.cuh) compiled for the host, with glb_iternorm_step's
    arguments minus the stream.  CPU test infrastructure (known-answer check of the product's
    recursion; test double of the C-ABI call in the host-flow tests)."""
    import ctypes as C
    import os
    import shutil
    import subprocess
    import tempfile

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which("g++")
    assert gxx, "g++ needed for the native host tests"
    out = os.path.join(tempfile.gettempdir(), f"glb_iternorm_host_{os.getpid()}.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(root, "tests", "native", "iternorm_host.cpp")], check=True)
    lib = C.CDLL(out)
    dp, ip = C.c_void_p, C.c_void_p
    lib.iternorm_step_host.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, ip]
    lib.iternorm_step_host.restype = C.c_int
    return lib


def native_iternorm_rows(rows):
    """The weights [a, s] of every row through the host build of K1's body; raises like the
    product where a flag comes back."""
    lib = native_iternorm()
    out = []
    state = None
    for i, row in enumerate(rows):
        row = np.array(row, dtype=np.float64)
        lead, k = row.shape[:-1], row.shape[-1] - 1
        n = int(np.prod(lead, dtype=np.int64))
        r = np.ascontiguousarray(row.reshape(n, k + 1))
        if state is None:
            kk = max(k, 1)
            state = [np.zeros((kk, kk, n)), np.zeros((kk, n)), np.ones(n), np.zeros((kk, n))]
        w = np.empty((n, k + 1))
        flag = np.zeros(1, dtype=np.int32)
        lib.iternorm_step_host(n, k, int(i == 0), r.ctypes.data, *(x.ctypes.data for x in state), w.ctypes.data, flag.ctypes.data)
        if flag[0]:
            raise ValueError("covariance matrix is not positive definite")
        out.append(w.reshape(*lead, k + 1))
    return out


def native_points_cuts():
    """Test double of glb_points_cuts: csrc/points_cuts.cuh (the header the kernel walks) compiled
    for the host behind the C-ABI call's arguments.  Returns an object with ``glb_points_cuts``."""
    import ctypes as C
    import os
    import shutil
    import subprocess
    import tempfile

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which("g++")
    assert gxx, "g++ needed for the native host tests"
    tmp = tempfile.mkdtemp(prefix="glb_cuts_")
    src = os.path.join(tmp, "chain.cpp")
    with open(src, "w") as f:
        f.write(
            '#include "%s"\nextern "C" void chain(const int64_t* off, int64_t npix, int64_t batch, int64_t start, int64_t remaining,'
            " int max_cuts, int64_t* cuts, int64_t* state) { glb::cuts_chain(off, npix, batch, start, remaining, max_cuts, cuts, state); }\n"
            'extern "C" void chain_list(const int64_t* gpix, int64_t total, int64_t npix, int64_t batch, int64_t start, int64_t remaining,'
            " int max_cuts, int64_t* cuts, int64_t* state) { glb::cuts_list_chain(gpix, total, npix, batch, start, remaining, max_cuts, cuts, state); }\n"
            % os.path.join(root, "glass_b200", "csrc", "points_cuts.cuh")
        )
    so = os.path.join(tmp, "chain.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True, capture_output=True, timeout=300)
    host = C.CDLL(so)
    host.chain.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]

    host.chain_list.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]

    class FakeLib:
        @staticmethod
        def glb_points_cuts(off, npix, batch, start, remaining, max_cuts, cuts, state, st):
            host.chain(off, npix, batch, start, remaining, max_cuts, cuts, state)
            return 0

        @staticmethod
        def glb_points_cuts_list(gpix, total, npix, batch, start, remaining, max_cuts, cuts, state, st):
            host.chain_list(gpix, total, npix, batch, start, remaining, max_cuts, cuts, state)
            return 0

    return FakeLib()
