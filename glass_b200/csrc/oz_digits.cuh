// oz_digits.cuh -- the digit arithmetic of the INT8 tensor-core Legendre kernel (sht_ozaki.cu), host/device so that the
// CPU suite can run it bit for bit (tests/native/oz_digits_host.cpp):
//   * t = fma(x, s, 2^52 + BIAS): the low 48 mantissa bits of t are V = rint(x s) + BIAS, and with BIAS = sum_j 128 256^j the
//     bytes u_j of V are the balanced base-256 digits d_j = u_j - 128 of rint(x s) (as an int8: u_j ^ 0x80);
//   * oz_planes: the six digit planes of four consecutive values, i.e. a 4 x 4 (low words) and a 4 x 2 (high words) byte
//     transpose with byte permutes, in the order the tensor core's K-major operand wants them (byte q of plane j = digit
//     j of value q);
//   * oz_in_range: every t must lie in [2^52, 2^52 + 2^48), i.e. its high word is 0x4330xxxx -- checked on the OR and the
//     AND of the high words of a whole half tile;
//   * oz_i2d: int32 -> double through the mantissa of 2^52 + 2^31 + x (no conversion instruction).
#pragma once
#include <stdint.h>
#include <string.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define GLB_OZ_HD __host__ __device__ __forceinline__
#else
#define GLB_OZ_HD inline
#endif

namespace glb {

constexpr int OZ_ND = 6;  // base-256 digits per operand

// the three bit-level intrinsics, with host equivalents
GLB_OZ_HD int oz_hi(double x) {
#ifdef __CUDA_ARCH__
  return __double2hiint(x);
#else
  int64_t b;
  memcpy(&b, &x, 8);
  return (int)(b >> 32);
#endif
}
GLB_OZ_HD int oz_lo(double x) {
#ifdef __CUDA_ARCH__
  return __double2loint(x);
#else
  int64_t b;
  memcpy(&b, &x, 8);
  return (int)(uint32_t)b;
#endif
}
GLB_OZ_HD double oz_hilo(int hi, int lo) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(hi, lo);
#else
  const int64_t b = ((int64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}
GLB_OZ_HD uint32_t oz_prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, sel);
#else
  const uint64_t v = ((uint64_t)b << 32) | a;  // bytes 0-3 of a, 4-7 of b
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
#endif
}

constexpr double OZ_HEADROOM = 0.99;  // |V| <= 0.99 * 2^47 keeps V + BIAS inside 48 bits
constexpr int64_t oz_bias() {
  int64_t b = 0;
  for (int j = 0; j < OZ_ND; ++j) b += (int64_t)128 << (8 * j);
  return b;
}
constexpr double OZ_MAGIC = 4503599627370496.0 + (double)oz_bias();  // 2^52 + BIAS

// scale s = 0.99 * 2^(47 - e) and its inverse for |x| < 2^e, from the biased exponent field eb of the
// largest |x| (e = eb - 1022); tiny maxima (eb < 64) count as zero
GLB_OZ_HD void oz_scales(int eb, double& s, double& inv) {
  if (eb < 64) {
    s = 0.0;
    inv = 0.0;
  } else {
    s = oz_hilo((2092 - eb) << 20, 0) * OZ_HEADROOM;
    inv = oz_hilo((eb - 46) << 20, 0) * (1.0 / OZ_HEADROOM);
  }
}

// the six digit planes of four consecutive values: a 4 x 4 byte transpose of the low words, a 4 x 2
// one of the high words (__byte_perm), and the sign flip of the balanced digits
GLB_OZ_HD void oz_planes(const double (&tt)[4], uint32_t (&w)[OZ_ND]) {
  const uint32_t l0 = (uint32_t)oz_lo(tt[0]), l1 = (uint32_t)oz_lo(tt[1]);
  const uint32_t l2 = (uint32_t)oz_lo(tt[2]), l3 = (uint32_t)oz_lo(tt[3]);
  const uint32_t h0 = (uint32_t)oz_hi(tt[0]), h1 = (uint32_t)oz_hi(tt[1]);
  const uint32_t h2 = (uint32_t)oz_hi(tt[2]), h3 = (uint32_t)oz_hi(tt[3]);
  const uint32_t t0 = oz_prmt(l0, l1, 0x5140), t1 = oz_prmt(l0, l1, 0x7362);
  const uint32_t t2 = oz_prmt(l2, l3, 0x5140), t3 = oz_prmt(l2, l3, 0x7362);
  const uint32_t s0 = oz_prmt(h0, h1, 0x5140), s1 = oz_prmt(h2, h3, 0x5140);
  w[0] = oz_prmt(t0, t2, 0x5410) ^ 0x80808080u;
  w[1] = oz_prmt(t0, t2, 0x7632) ^ 0x80808080u;
  w[2] = oz_prmt(t1, t3, 0x5410) ^ 0x80808080u;
  w[3] = oz_prmt(t1, t3, 0x7632) ^ 0x80808080u;
  w[4] = oz_prmt(s0, s1, 0x5410) ^ 0x80808080u;
  w[5] = oz_prmt(s0, s1, 0x7632) ^ 0x80808080u;
}

// int32 -> double without a conversion instruction: the mantissa of 2^52 + 2^31 + x
GLB_OZ_HD double oz_i2d(int x) {
  return oz_hilo(0x43300000, (int)((uint32_t)x ^ 0x80000000u)) - 4503601774854144.0;
}

GLB_OZ_HD bool oz_in_range(uint32_t hor, uint32_t hand) {
  return (hor & 0xffff0000u) == 0x43300000u && (hand & 0xfff00000u) == 0x43300000u;
}

}  // namespace glb
