// Host check of glass_b200/csrc/oz_digits.cuh: the digit arithmetic of the INT8 tensor-core Legendre kernel, bit for bit.
//   * the bytes of fma(x, s, 2^52 + BIAS), sign-flipped, are the balanced base-256 digits of rint(x s);
//   * oz_planes puts digit j of value q into byte q of plane j (what the tensor core's K-major operand rows hold);
//   * oz_in_range accepts exactly the values with |x s| < 2^47 (about) and rejects overflow on either side, Inf and NaN;
//   * oz_i2d is exact for every int32; oz_scales gives 0.99 * 2^(47 - e) and its inverse from the exponent field.
// Built and run by tests/test_cpu_host.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../glass_b200/csrc/oz_digits.cuh"

using namespace glb;

static int fails = 0;
#define CHECK(c)                                              \
  do {                                                        \
    if (!(c)) {                                               \
      if (++fails < 10) printf("FAILED line %d: %s\n", __LINE__, #c); \
    }                                                         \
  } while (0)

int main() {
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(-1.0, 1.0);
  CHECK(oz_bias() == 0x808080808080ll);
  CHECK(OZ_MAGIC == 4503599627370496.0 + 141289400074368.0);
  // byte permute shim against its definition
  CHECK(oz_prmt(0x33221100u, 0x77665544u, 0x5140) == 0x55114400u);
  CHECK(oz_prmt(0x33221100u, 0x77665544u, 0x7362) == 0x77336622u);
  CHECK(oz_prmt(0x33221100u, 0x77665544u, 0x5410) == 0x55441100u);
  CHECK(oz_prmt(0x33221100u, 0x77665544u, 0x7632) == 0x77663322u);
  for (int eb : {1023 - 300, 1023 - 70, 1000, 1023, 1030, 1500}) {
    double s, inv;
    oz_scales(eb, s, inv);
    CHECK(s == std::ldexp(0.99, 47 - (eb - 1022)));
    CHECK(std::fabs(s * inv - 1.0) < 1e-15);
    const double bound = std::ldexp(1.0, eb - 1022);  // |x| < bound
    for (int rep = 0; rep < 2000; ++rep) {
      double x[4], tt[4];
      uint32_t hor = 0u, hand = 0xffffffffu;
      for (int q = 0; q < 4; ++q) {
        x[q] = (rep == 0 && q == 0) ? 0.0 : (rep == 1 ? (q & 1 ? 1 : -1) * std::nextafter(bound, 0.0) : bound * U(rng) * std::ldexp(1.0, -(int)(rng() % 40)));
        tt[q] = std::fma(x[q], s, OZ_MAGIC);
        hor |= (uint32_t)oz_hi(tt[q]);
        hand &= (uint32_t)oz_hi(tt[q]);
      }
      CHECK(oz_in_range(hor, hand));
      uint32_t w[OZ_ND];
      oz_planes(tt, w);
      for (int q = 0; q < 4; ++q) {
        // digits of value q: byte q of every plane, as signed bytes
        long double v = 0;
        for (int j = OZ_ND - 1; j >= 0; --j) v = v * 256 + (int8_t)((w[j] >> (8 * q)) & 0xff);
        const long double want = std::nearbyint((long double)x[q] * (long double)s);
        CHECK(std::fabs((double)(v - want)) <= 1.0);  // (x s is rounded once in the FMA, twice here)
        CHECK(std::fabs((double)v * inv - x[q]) <= 0.51 * inv * 1.0000001 + 1e-300 + std::fabs(x[q]) * 2e-16);
      }
    }
    // out of range: beyond the bound on either side, Inf, NaN
    for (double bad : {bound * 1.02, -bound * 1.02, bound * 64.0, -bound * 1e30, (double)INFINITY, -(double)INFINITY, (double)NAN}) {
      const double t = std::fma(bad, s, OZ_MAGIC);
      const double ok = std::fma(0.0, s, OZ_MAGIC);
      CHECK(!oz_in_range((uint32_t)oz_hi(t) | (uint32_t)oz_hi(ok), (uint32_t)oz_hi(t) & (uint32_t)oz_hi(ok)));
    }
  }
  {
    double s, inv;
    oz_scales(10, s, inv);  // tiny maxima count as zero
    CHECK(s == 0.0 && inv == 0.0);
  }
  for (int rep = 0; rep < 200000; ++rep) {
    const int v = (int)(uint32_t)rng();
    CHECK(oz_i2d(v) == (double)v);
  }
  CHECK(oz_i2d(0) == 0.0 && oz_i2d(INT32_MIN) == -2147483648.0 && oz_i2d(INT32_MAX) == 2147483647.0);
  if (fails) {
    printf("%d failures\n", fails);
    return 1;
  }
  printf("oz_digits ok\n");
  return 0;
}
