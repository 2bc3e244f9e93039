"""ctypes wrapper of the C oracle (oracle/sht_ref.c).  TEST INFRASTRUCTURE ONLY."""

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
        if not path.exists() or not stamp.exists() or stamp.read_text() != host:
            subprocess.run(["make", "-C", os.fspath(_HERE), "clean"], capture_output=True)
            build()
            stamp.write_text(host)
        lib = C.CDLL(os.fspath(path))
        for fn in (lib.ref_alm2map, lib.ref_alm2phase):
            fn.restype = C.c_int
            fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ref_max_threads.restype = C.c_int
        _LIBS[long_double] = lib
    return _LIBS[long_double]


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
