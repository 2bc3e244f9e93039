// rng.cuh -- counter-based Philox4x32-10 (Salmon et al. 2011) and the deviate transforms
// used in place of NumPy's sequential PCG64 streams (glass/rng.py:58-247).  Every draw is
// a pure function of (seed, stream id, element index) so any shell / pixel / galaxy can be
// regenerated independently on any GPU.
#pragma once
#include <stdint.h>

namespace glb {

struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = mulhilo32(M0, c0, &hi0);
    const uint32_t lo1 = mulhilo32(M1, c2, &hi1);
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n1 = lo1;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    const uint32_t n3 = lo0;
    c0 = n0;
    c1 = n1;
    c2 = n2;
    c3 = n3;
    k0 += W0;
    k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0;
  o.v[1] = c1;
  o.v[2] = c2;
  o.v[3] = c3;
  return o;
}

// 53-bit uniform in [0, 1)  (same construction as NumPy's random_sample: (a>>5, b>>6))
__host__ __device__ __forceinline__ double u01_closed_open(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// uniform in (0, 1]
__host__ __device__ __forceinline__ double u01_open_closed(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 1.0) * (1.0 / 9007199254740992.0);
}

// stream tags (third/fourth counter words) keep the deviates of different uses disjoint
enum : uint32_t {
  RNG_TAG_ALM = 0x414c4d00u,      // standard normals for a_lm
  RNG_TAG_POISSON = 0x504f4900u,  // per-pixel galaxy counts
  RNG_TAG_POS = 0x504f5300u,      // in-pixel positions (u, v)
  RNG_TAG_EPS = 0x45505300u,      // ellipticities
  RNG_TAG_REDSHIFT = 0x5a5a5a00u  // redshifts
};

}  // namespace glb
