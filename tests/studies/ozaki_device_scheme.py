"""
CPU emulation (NumPy, exact integers) of the DEVICE arithmetic of the INT8 Legendre contraction
(glass_b200/csrc/sht_ozaki.cu): what each thread and the tensor core compute, step by step, so that
the formats are fixed before any GPU time is spent.

  p (FP64, one ring, 64 consecutive l-pairs)  ->  V = rint(p * s) + BIAS  by ONE FMA with the magic
      constant 2^52 + BIAS (the low 48 mantissa bits ARE V);  s = 0.99 * 2^(47 - e), 2^e > max |p| of the
      tile (from the exponent field alone);  BIAS = sum_j 128 * 256^j, so that the BYTES u_j of V are the
      balanced digits d_j = u_j - 128 in [-128, 127] (as int8: u_j ^ 0x80)
  A (FP64 coefficients, one column, the same l-pairs) -> the same digits relative to the column's bound
  D_g = sum_{i + j = 5 + g} sum_k dP_i[k] dA_j[k]      g = 0..5, exact in int32 (tcgen05.mma kind::i8)
  F  += (1 / sP) (256^5 / sA) * (P0 + 2^16 P1 + 2^32 P2),   P_i = D_2i + 256 D_2i+1  in int32

    python tests/studies/ozaki_device_scheme.py [nside] [KT]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import healpix_ref as H  # noqa: E402

ND = 6
BIAS = sum(128 * 256**j for j in range(ND))
MAGIC = float(2**52 + BIAS)
HEADROOM = 0.99


def scales(maxabs):
    """oz_scales: from the biased exponent field of the largest |x| (tiny maxima count as zero)."""
    eb = (np.asarray(maxabs, dtype=np.float64).view(np.int64) >> 52) & 0x7FF
    ok = eb >= 64
    s = np.where(ok, np.ldexp(HEADROOM, np.where(ok, 2092 - eb - 1023, 0)), 0.0)
    inv = np.where(ok, np.ldexp(1.0 / HEADROOM, np.where(ok, eb - 46 - 1023, 0)), 0.0)
    return s, inv


def digits(x, s):
    """x [..] and scale s -> int64 balanced digits [ND, ..] in [-128, 127] (bytes of the FMA result ^ 0x80)."""
    t = x * s + MAGIC  # exact product here (s is 0.99 * 2^n: one rounding in the product, one in the sum; the device
    # fuses them -- the digits can differ by one unit of the last digit, which is the quantisation anyway)
    v = t.view(np.int64) & ((1 << 52) - 1)
    assert np.all(v >> 48 == 0), "fixed-point overflow"
    u = np.stack([(v >> (8 * j)) & 255 for j in range(ND)])
    return (u ^ 0x80).astype(np.uint8).view(np.int8).astype(np.int64)


def contraction(P, A, KT):
    nring, K = P.shape
    F = np.zeros((nring, A.shape[1]))
    worst = worst_pair = 0
    for k0 in range(0, K, KT):
        p, a = P[:, k0 : k0 + KT], A[k0 : k0 + KT]
        sP, iP = scales(np.abs(p).max(axis=1))
        sA, iA = scales(np.abs(a).max(axis=0))
        dP = digits(p, sP[:, None])  # [6, ring, k]
        dA = digits(a, sA[None, :])  # [6, k, col]
        rec = sum(dP[j].astype(np.float64) * 256.0**j for j in range(ND)) * iP[:, None]
        assert np.abs(rec - p).max() <= 1.01 * iP.max()
        D = []
        for g in range(ND):
            acc = np.zeros((nring, A.shape[1]), dtype=np.int64)
            for i in range(ND):
                j = 5 + g - i
                if 0 <= j < ND:
                    acc += dP[i] @ dA[j]
            worst = max(worst, int(np.abs(acc).max()))
            D.append(acc)
        P0, P1, P2 = D[0] + 256 * D[1], D[2] + 256 * D[3], D[4] + 256 * D[5]
        worst_pair = max(worst_pair, int(max(np.abs(P0).max(), np.abs(P1).max(), np.abs(P2).max())))
        val = (P2.astype(np.float64) * 65536.0 + P1) * 65536.0 + P0
        F += val * (iP[:, None] * (iA[None, :] * 2.0**40))
    assert worst < 6 * KT * 128 * 128 + 1 and worst_pair < 2**31
    return F, worst


def main():
    nside = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    KT = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    lmax = 2 * nside - 1
    ri = H.ring_info(nside)
    z, sth = ri["z"][: 2 * nside], ri["sth"][: 2 * nside]
    rng = np.random.default_rng(1)
    print(f"nside {nside} lmax {lmax}, {KT} l-pairs per tile, 6 x 6 base-256 digits, products with i + j >= 5 (21 of 36)")
    for m in (0, 1, lmax // 4, lmax // 2, (3 * lmax) // 4):
        lam = H.lam_lm(lmax, m, z, sth).T[:, ::2]  # even-offset functions lambda_{m + 2k}: what p_k is up to alpha_k
        l = np.arange(m, lmax + 1, 2)
        a = rng.standard_normal((l.size, 16)) * np.sqrt(1e-2 * (l + 1.0) ** -1.5)[:, None]
        ref = lam @ a
        F, worst = contraction(lam, a, KT)
        print(f"m {m:5d}: max |F - F_fp64| / max |F_fp64| = {np.abs(F - ref).max() / np.abs(ref).max():.2e}   max |D_g| = 2^{np.log2(max(worst, 1)):.1f}")


if __name__ == "__main__":
    main()
