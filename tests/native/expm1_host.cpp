// Host check of glass_b200/csrc/expm1_fast.cuh against 80-bit expm1l.  Built and run by
// tests/test_cpu_host.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../glass_b200/csrc/expm1_fast.cuh"

static double ulp_err(double got, long double want) {
  if (want == 0.0L) return got == 0.0 ? 0.0 : 1e300;
  const double w = (double)want;
  const double ulp = std::fabs(std::nextafter(w, INFINITY) - w);
  return (double)(fabsl((long double)got - want) / ulp);
}

int main() {
  double worst = 0.0, worst_x = 0.0;
  unsigned long long s = 88172645463325252ULL;
  auto rnd = [&]() {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    return (double)(s >> 11) / 9007199254740992.0;
  };
  for (int i = 0; i < 4000000; ++i) {
    const double x = (rnd() - 0.5) * 80.0;
    const double e = ulp_err(glb::expm1_fast(x), expm1l((long double)x));
    if (e > worst) { worst = e; worst_x = x; }
  }
  for (int i = 0; i < 1000000; ++i) {  // small and tiny arguments, both signs
    const double x = (rnd() - 0.5) * std::pow(2.0, -60.0 * rnd());
    const double e = ulp_err(glb::expm1_fast(x), expm1l((long double)x));
    if (e > worst) { worst = e; worst_x = x; }
  }
  for (int i = 0; i < 1000000; ++i) {  // around the reduction boundaries (n + 1/2) ln 2
    const int n = (int)(rnd() * 40) - 20;
    const double x = (n + 0.5) * 0.6931471805599453 + (rnd() - 0.5) * 1e-6;
    const double e = ulp_err(glb::expm1_fast(x), expm1l((long double)x));
    if (e > worst) { worst = e; worst_x = x; }
  }
  const double specials[] = {0.0, -0.0, 699.9, -699.9, 700.0, -800.0, 710.0, INFINITY, -INFINITY};
  for (double x : specials) {
    const double g = glb::expm1_fast(x), w = std::expm1(x);
    if (!(g == w || ulp_err(g, expm1l((long double)x)) <= 2.0)) {
      std::printf("special %g: got %.17g want %.17g\n", x, g, w);
      return 1;
    }
  }
  if (!std::isnan(glb::expm1_fast(NAN))) { std::printf("nan\n"); return 1; }
  std::printf("worst %.3f ulp at x = %.17g\n", worst, worst_x);
  if (worst > 2.0) return 1;
  std::printf("expm1_fast ok\n");
  return 0;
}
