// sht_seed.cuh -- range handling shared by the Legendre synthesis kernels (sht_legendre.cu, sht_ozaki.cu):
// values carry an integer scale (true = v * 2^(512 * scale)); the seed lambda_mm as (value, scale).
#pragma once
#include "common.cuh"

namespace glb {

constexpr int SCALE_BITS = 512;
constexpr int BEXP_BIG = 1023 + 256;  // rescale when |p| >= 2^256
constexpr int BEXP_SIG = 1023 - 70;   // "significant" when scale==0 and |p| >= 2^-70

__device__ __forceinline__ int bexp(double v) { return (__double2hiint(v) >> 20) & 0x7ff; }

// -------------------------------------------------------------------------------------
// lambda_mm(theta) = (-1)^m c_m sin^m(theta) as (value, scale), true = value*2^(512*scale)
// -------------------------------------------------------------------------------------
__device__ __forceinline__ void lam_mm_scaled(int m, double sth, double cm_mant, int cm_exp, double& val,
                                              int& scale) {
  int e;
  double bv = frexp(sth, &e);  // sth = bv * 2^e, bv in [0.5, 1)
  int be = e;
  double rv = 1.0;
  int re = 0;
  int mm = m;
  while (mm) {
    if (mm & 1) {
      rv *= bv;
      re += be;
      if (rv < 0.5) {
        rv *= 2.0;
        re -= 1;
      }
    }
    bv *= bv;
    be *= 2;
    if (bv < 0.5) {
      bv *= 2.0;
      be -= 1;
    }
    mm >>= 1;
  }
  double mant = rv * cm_mant;  // in [0.25, 1)
  int E = re + cm_exp;
  if (m & 1) mant = -mant;
  if (E >= 0) {
    scale = 0;
    val = scalbn(mant, E);
  } else {
    const int s = (-E) / SCALE_BITS;  // truncation
    scale = -s;
    val = scalbn(mant, E + s * SCALE_BITS);  // exponent in (-512, 0]
  }
}

}  // namespace glb
