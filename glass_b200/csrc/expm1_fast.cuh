// expm1_fast.cuh -- e^x - 1 for the lognormal pixel transform (K5, glass/grf/_transformations.py:83-89).
//
// One call per pixel sits in the store of the ring FFT; the CUDA library expm1 costs ~58 SASS
// instructions there (ncu: 21 % of the kernel's instructions).  This one is ~25:
//   n = rint(x log2 e), r = x - n ln2 (two FMAs, |r| <= 0.3466),
//   e^r - 1 = r + r^2 (1/2! + r/3! + ... + r^11/13!)   (truncation 1.2e-17 relative)
//   e^x - 1 = 2^n (e^r - 1) + (2^n - 1)                 (one FMA; exact for n = 0)
// Measured against 80-bit expm1l on 4e6 points in [-40, 40] and on tiny arguments: <= 2 ulp
// (tests/native/expm1_host.cpp).  Arguments with |x| >= 700, Inf and NaN go to the library.
#pragma once
#include <math.h>
#include <string.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define GLB_EXPM1_HD __host__ __device__ __forceinline__
#else
#define GLB_EXPM1_HD inline
#endif

namespace glb {

GLB_EXPM1_HD double expm1_fast(double x) {
  if (!(fabs(x) < 700.0)) return expm1(x);
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: the sum's low word is n
  const double tm = fma(x, 1.4426950408889634074, MAGIC);
#ifdef __CUDA_ARCH__
  const int n = __double2loint(tm);
#else
  long long bits;
  memcpy(&bits, &tm, 8);
  const int n = (int)(unsigned)(bits & 0xffffffffLL);
#endif
  const double nf = tm - MAGIC;
  double r = fma(nf, -6.93147180559945286227e-01, x);
  r = fma(nf, -2.31904681384629955842e-17, r);
  double q = 1.6059043836821614599e-10;      // 1/13!
  q = fma(q, r, 2.0876756987868098979e-09);  // 1/12!
  q = fma(q, r, 2.5052108385441718775e-08);  // 1/11!
  q = fma(q, r, 2.7557319223985890653e-07);  // 1/10!
  q = fma(q, r, 2.7557319223985892511e-06);  // 1/9!
  q = fma(q, r, 2.4801587301587301566e-05);  // 1/8!
  q = fma(q, r, 1.9841269841269841253e-04);  // 1/7!
  q = fma(q, r, 1.3888888888888889419e-03);  // 1/6!
  q = fma(q, r, 8.3333333333333332177e-03);  // 1/5!
  q = fma(q, r, 4.1666666666666664354e-02);  // 1/4!
  q = fma(q, r, 1.6666666666666665741e-01);  // 1/3!
  q = fma(q, r, 0.5);
  const double p = fma(r * r, q, r);
#ifdef __CUDA_ARCH__
  const double s = __hiloint2double((n + 1023) << 20, 0);
#else
  const long long sb = (long long)(n + 1023) << 52;
  double s;
  memcpy(&s, &sb, 8);
#endif
  return fma(s, p, s - 1.0);
}

}  // namespace glb
