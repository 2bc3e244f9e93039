// Host check of glass_b200/csrc/sht_tables.cuh (double-double coefficient tables of the
// Legendre stages) against 80-bit long double: every entry within 0.51 ulp of the extended
// value, and the product a_k a_{k-1} (e_l e_{l-1})^2 - 1 without a systematic offset.
// Built and run by tests/test_cpu_host.py.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../glass_b200/csrc/sht_tables.cuh"

typedef long double ld;

static ld eps_ld(int l, int m) {
  if (l <= m) return 0.0L;
  return sqrtl(((ld)(l - m) * (ld)(l + m)) / (4.0L * l * l - 1.0L));
}

static double ulps(double got, ld want) {
  const double w = (double)want;
  const double u = std::fabs(std::nextafter(w, INFINITY) - w);
  return (double)(fabsl((ld)got - want) / u);
}

int main() {
  const int cases[][2] = {{8191, 0}, {8191, 1}, {8191, 17}, {8191, 4000}, {8191, 8190}, {8191, 8191}, {383, 0}, {383, 200}, {7, 3}};
  double worst = 0.0;
  for (auto& cs : cases) {
    const int lmax = cs[0], m = cs[1];
    const int K = (lmax - m) / 2 + 1;
    std::vector<double> t((size_t)K * glb::PREP_TAB);
    glb::prep_tables_for_m(lmax, m, t.data());
    ld akm1 = 0, ak = 1;
    ld bias = 0;
    for (int k = 0; k < K; ++k) {
      const int l = m + 2 * k;
      const ld e_lm1 = eps_ld(l - 1, m), e_l = eps_ld(l, m), e1 = eps_ld(l + 1, m), e2 = eps_ld(l + 2, m), e3 = eps_ld(l + 3, m);
      const ld akp1 = (k == 0) ? 1.0L : akm1 * ((e_l * e_lm1) / (e1 * e2));
      const ld a = ak / (e1 * e2 * akp1);
      const ld want[6] = {a, -(e1 * e1 + e_l * e_l) * a, (1.0L - (e1 * e1 + e_l * e_l)) * a, ak, ak / e1, e2 / e3};
      for (int j = 0; j < 6; ++j) {
        const double u = ulps(t[(size_t)k * 6 + j], want[j]);
        // the 80-bit recursion for alpha itself drifts by ~sqrt(k) 2^-64: allow for it
        const double tol = 0.51 + 1e-3 * std::sqrt((double)k + 1.0);
        if (u > tol) {
          std::printf("lmax %d m %d k %d entry %d: %.3f ulp\n", lmax, m, k, j, u);
          return 1;
        }
        if (u > worst) worst = u;
      }
      if (k > 0) bias += (ld)t[(size_t)k * 6] * (ld)t[(size_t)(k - 1) * 6] * (e_l * e_lm1) * (e_l * e_lm1) - 1.0L;
      akm1 = ak;
      ak = akp1;
    }
    if (K > 100) {
      const double mean_ulp = (double)(bias / (K - 1)) / 1.11e-16;
      std::printf("lmax %d m %d: mean of a_k a_{k-1} (e_l e_{l-1})^2 - 1 = %+.3f ulp\n", lmax, m, mean_ulp);
      if (std::fabs(mean_ulp) > 0.25) return 1;
    }
  }
  std::printf("worst entry %.3f ulp\ntables ok\n", worst);
  return 0;
}
