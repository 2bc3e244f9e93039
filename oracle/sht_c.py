"""ctypes wrapper of the C oracle (oracle/sht_ref.c, the checker) and of the timed CPU arm
(oracle/sht_fast.cpp).  TEST INFRASTRUCTURE ONLY."""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIBS: dict[bool, C.CDLL] = {}


def build(quiet: bool = True) -> None:
    """make -C oracle (march=native: built on the machine that runs it)."""
    r = subprocess.run(["make", "-C", os.fspath(_HERE)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
    if not quiet:
        print(r.stdout)


def _lib(long_double: bool) -> C.CDLL:
    if long_double not in _LIBS:
        name = "liboracle_sht_ld.so" if long_double else "liboracle_sht.so"
        path = _HERE / "_build" / name
        stamp = _HERE / "_build" / "host.txt"
        host = os.uname().nodename + ":" + str(os.cpu_count())
        # -march=native objects must be rebuilt when the snapshot lands on another host
        if not stamp.exists() or stamp.read_text() != host:
            subprocess.run(["make", "-C", os.fspath(_HERE), "clean"], capture_output=True)
        build()  # make: a no-op when the libraries are newer than their sources
        stamp.write_text(host)
        lib = C.CDLL(os.fspath(path))
        for fn in (lib.ref_alm2map, lib.ref_alm2phase):
            fn.restype = C.c_int
            fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ref_max_threads.restype = C.c_int
        lib.ref_map2alm_pass.restype = C.c_int
        lib.ref_map2alm_pass.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_alm2map_spin.restype = C.c_int
        lib.ref_alm2map_spin.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIBS[long_double] = lib
    return _LIBS[long_double]


_FAST: list = []


def _fast() -> C.CDLL:
    if not _FAST:
        _lib(False)  # (re)builds everything for this host
        lib = C.CDLL(os.fspath(_HERE / "_build" / "liboracle_sht_fast.so"))
        lib.fast_alm2map.restype = C.c_int
        lib.fast_alm2map.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        lib.fast_alm2map_timed.restype = C.c_int
        lib.fast_alm2map_timed.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _FAST.append(lib)
    return _FAST[0]


def alm2map_fast(alm, nside: int, lmax: int, *, nthreads: int = 0, timings: dict | None = None):
    """The timed CPU arm: same result as ``alm2map(..., use_mlim=True)`` from the SIMD / register
    blocked implementation in sht_fast.cpp.  ``timings`` receives the seconds of the two stages."""
    alm = np.ascontiguousarray(alm, dtype=np.complex128)
    assert alm.size == (lmax + 1) * (lmax + 2) // 2
    out = np.empty(12 * nside * nside)
    tl, tf = C.c_double(0), C.c_double(0)
    rc = _fast().fast_alm2map_timed(nside, lmax, alm.ctypes.data, out.ctypes.data, nthreads, C.byref(tl), C.byref(tf))
    if rc != 0:
        raise MemoryError("fast_alm2map") if rc == -1 else ValueError("fast_alm2map: nside must be a power of two >= 2")
    if timings is not None:
        timings["legendre_s"], timings["fft_s"] = tl.value, tf.value
    return out


def max_threads() -> int:
    return int(_lib(False).ref_max_threads())


def alm2map(alm, nside: int, lmax: int, *, long_double: bool = False, use_mlim: bool = False, nthreads: int = 0):
    """Scalar synthesis, healpy.alm2map(pol=False) semantics (glass/healpix.py:71)."""
    alm = np.ascontiguousarray(alm, dtype=np.complex128)
    assert alm.size == (lmax + 1) * (lmax + 2) // 2
    out = np.empty(12 * nside * nside)
    rc = _lib(long_double).ref_alm2map(nside, lmax, alm.ctypes.data, out.ctypes.data, int(use_mlim), nthreads)
    if rc != 0:
        raise MemoryError("ref_alm2map")
    return out


def alm2phase(alm, nside: int, lmax: int, *, long_double: bool = False, use_mlim: bool = False, nthreads: int = 0):
    alm = np.ascontiguousarray(alm, dtype=np.complex128)
    out = np.empty((4 * nside - 1, lmax + 1), dtype=np.complex128)
    _lib(long_double).ref_alm2phase(nside, lmax, alm.ctypes.data, out.ctypes.data, int(use_mlim), nthreads)
    return out


def map2alm(mp, lmax: int, *, niter: int = 3, ring_w=None, long_double: bool = False, nthreads: int = 0):
    """Scalar analysis, healpy.map2alm(pol=False) semantics (glass/healpix.py:270): one analysis pass
    and ``niter`` Jacobi refinements alm += A(map - S(alm)) (healpy's ``iter``); ``ring_w`` stands in
    for healpy's pixel weights (default uniform).  Same definition as oracle/healpix_ref.py::map2alm."""
    mp = np.ascontiguousarray(mp, dtype=np.float64)
    nside = int(round((mp.size / 12) ** 0.5))
    nalm = (lmax + 1) * (lmax + 2) // 2
    w = None if ring_w is None else np.ascontiguousarray(ring_w, dtype=np.float64)
    lib = _lib(long_double)

    def analysis(m):
        out = np.zeros(nalm, dtype=np.complex128)
        rc = lib.ref_map2alm_pass(nside, lmax, m.ctypes.data, None if w is None else w.ctypes.data, out.ctypes.data, nthreads)
        if rc != 0:
            raise MemoryError("ref_map2alm_pass")
        return out

    alm = analysis(mp)
    for _ in range(niter):
        res = np.ascontiguousarray(mp - alm2map(alm, nside, lmax, long_double=long_double, nthreads=nthreads))
        alm = alm + analysis(res)
    return alm


def alm2map_spin(alm1, alm2, nside: int, spin: int, lmax: int, *, long_double: bool = False, nthreads: int = 0):
    """healpy.alm2map_spin semantics (glass/healpix.py:107); ``alm2`` may be None (E-only)."""
    a1 = np.ascontiguousarray(alm1, dtype=np.complex128)
    a2 = None if alm2 is None else np.ascontiguousarray(alm2, dtype=np.complex128)
    assert a1.size == (lmax + 1) * (lmax + 2) // 2
    m1, m2 = np.empty(12 * nside * nside), np.empty(12 * nside * nside)
    rc = _lib(long_double).ref_alm2map_spin(nside, lmax, spin, a1.ctypes.data, None if a2 is None else a2.ctypes.data, m1.ctypes.data, m2.ctypes.data, nthreads)
    if rc != 0:
        raise MemoryError("ref_alm2map_spin")
    return m1, m2
