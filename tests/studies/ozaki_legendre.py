"""
Numerical feasibility of moving the Legendre contraction  F[ring, c] = sum_l lambda_lm(ring) A[l, c]
from the FP64 pipe to INT8 tensor cores (Ozaki-type splitting), CPU emulation with NumPy only.

Both operands are cut into signed 7-bit slices relative to a power-of-two scale per (ring, l-tile)
and per (column, l-tile); slice products are accumulated exactly in integers (what
tcgen05.mma kind::i8 with int32 accumulators does) and the partial sums of the tiles are added in
FP64.  Reported: relative error of F against the FP64 contraction as a function of the number of
slices d (all pairs i + j <= d + 1 are used: d (d + 1) / 2 integer products).

    python tests/studies/ozaki_legendre.py [nside] [tile]

This is design evidence for DESIGN.md section 7, not part of the product.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import healpix_ref as H  # noqa: E402


def slices(x, scale_exp, d):
    """x / 2^scale_exp in (-1, 1) -> d signed 7-bit integer slices s_i with
    x ~ 2^scale_exp * sum_i s_i 2^(-7 i)."""
    y = np.ldexp(x, -scale_exp)
    out = []
    for _ in range(d):
        y = y * 128.0
        s = np.rint(y)
        s = np.clip(s, -127, 127)
        out.append(s.astype(np.int64))
        y = y - s
    return out


def contraction(lam, a, d, tile):
    """lam [nring, nl], a [nl, ncol] -> sum over l in tiles, sliced arithmetic."""
    nring, nl = lam.shape
    f = np.zeros((nring, a.shape[1]))
    for t0 in range(0, nl, tile):
        lt, at = lam[:, t0 : t0 + tile], a[t0 : t0 + tile]
        el = np.frexp(np.abs(lt).max(axis=1))[1] + 1  # per ring
        ea = np.frexp(np.abs(at).max(axis=0))[1] + 1  # per column
        el = np.where(np.abs(lt).max(axis=1) == 0, 0, el)
        ea = np.where(np.abs(at).max(axis=0) == 0, 0, ea)
        sl = slices(lt, el[:, None], d)
        sa = slices(at, ea[None, :], d)
        acc = np.zeros((nring, a.shape[1]))
        for i in range(d):
            for j in range(d):
                if i + j + 2 <= d + 1:
                    prod = sl[i] @ sa[j]  # exact integer matrix product
                    acc += np.ldexp(prod.astype(np.float64), -7 * (i + 1) - 7 * (j + 1))
        f += np.ldexp(acc, (el[:, None] + ea[None, :]))
    return f


def main():
    nside = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    tile = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    lmax = 2 * nside - 1
    ri = H.ring_info(nside)
    z, sth = ri["z"][: 2 * nside], ri["sth"][: 2 * nside]
    rng = np.random.default_rng(1)
    print(f"nside {nside} lmax {lmax}, l-tile {tile}: max |F_sliced - F_fp64| / max |F_fp64| over rings and 16 columns")
    print("   m    " + "  ".join(f"d={d} ({d*(d+1)//2:2d} products)" for d in (4, 5, 6, 7)))
    for m in (0, 1, lmax // 4, lmax // 2, (3 * lmax) // 4):
        lam = H.lam_lm(lmax, m, z, sth).T  # [ring, l - m]
        l = np.arange(m, lmax + 1)
        cl = 1e-2 * (l + 1.0) ** -1.5
        a = rng.standard_normal((l.size, 16)) * np.sqrt(cl)[:, None]
        ref = lam @ a
        errs = []
        for d in (4, 5, 6, 7):
            f = contraction(lam, a, d, tile)
            errs.append(np.abs(f - ref).max() / np.abs(ref).max())
        print(f"{m:5d}   " + "  ".join(f"{e:18.2e}" for e in errs))


if __name__ == "__main__":
    main()
