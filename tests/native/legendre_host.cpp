// Host replay of the arithmetic of sht_legendre_synth_kernel's recurrence (csrc/sht_legendre.cu)
// at nside 4096, lmax 8191: the double-double coefficient table (csrc/sht_tables.cuh), the
// per-ring choice of the recurrence variable (y = z^2 where z^2 < 1/2, u = sin^2(theta)
// otherwise) and the FMA chain  p_{k+1} = fma(fma(A, v, C), p_k, -p_{k-1}),  compared with the
// standard three-term recurrence in 80-bit arithmetic on exact ring geometry.  Also replays the
// two formulations that were measured and discarded on the GPU (y everywhere, u everywhere) and
// checks that they really are worse, so the test documents why the kernel does what it does.
// Built and run by tests/test_cpu_host.py.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../glass_b200/csrc/sht_tables.cuh"

typedef long double ld;

static const int N = 4096, LMAX = 8191;

// exact geometry of northern ring pair r (0-based), SURVEY.md appendix A.1
static void ring_ld(int r, ld& z, ld& sth) {
  const int i = r + 1;
  if (i < N) {
    const ld t = (ld)i * i / (3.0L * N * N);
    z = 1.0L - t;
    sth = sqrtl(t * (2.0L - t));
  } else {
    z = (2.0L * N - i) * 2.0L / (3.0L * N);
    sth = sqrtl((1.0L - z) * (1.0L + z));
  }
}
static void ring_d(int r, double& z, double& sth) {  // as glass_b200/csrc/plan.cu computes them
  const int i = r + 1;
  if (i < N) {
    const double t = (double)i * i / (3.0 * N * N);
    z = 1.0 - t;
    sth = std::sqrt(t * (2.0 - t));
  } else {
    z = (2.0 * N - i) * 2.0 / (3.0 * N);
    sth = std::sqrt((1.0 - z) * (1.0 + z));
  }
}

// lambda_{l,m} / lambda_{m,m} by the standard recurrence in long double
static ld ratio_ref(int l, int m, ld z) {
  ld pp = 0.0L, p = 1.0L;
  for (int ll = m + 1; ll <= l; ++ll) {
    const ld a = sqrtl((4.0L * ll * ll - 1.0L) / ((ld)ll * ll - (ld)m * m));
    const ld b = (ll > m + 1) ? sqrtl((((ld)ll - 1) * ((ld)ll - 1) - (ld)m * m) / (4.0L * ((ld)ll - 1) * ((ld)ll - 1) - 1.0L)) : 0.0L;
    const ld t = a * (z * p - b * pp);
    pp = p;
    p = t;
  }
  return p;
}

// the kernel's chain for l = m + 2k: alpha_k p_k with p_0 = 1; mode 0 = per-ring choice, 1 = y, 2 = u
static double ratio_kernel(const std::vector<double>& t, int k, double z, double sth, int mode) {
  const bool use_u = (mode == 0) ? (z * z >= 0.5) : (mode == 2);
  const double v = use_u ? sth * sth : z * z;
  double p1 = 0.0, p2 = 1.0;
  for (int kk = 0; kk < k; ++kk) {
    const double* tk = &t[(size_t)kk * glb::PREP_TAB];
    const double A = use_u ? -tk[glb::TAB_A] : tk[glb::TAB_A];
    const double C = use_u ? tk[glb::TAB_AB] : tk[glb::TAB_B];
    const double rr = std::fma(A, v, C);
    const double tt = std::fma(rr, p2, -p1);
    p1 = p2;
    p2 = tt;
  }
  return t[(size_t)k * glb::PREP_TAB + glb::TAB_ALPHA] * p2;
}

int main() {
  const int ms[] = {0, 2, 17};
  // rings next to the pole, in the cap, at the switch of the variable, in the belt, next to the equator
  const int rings[] = {0, 1, 2, 5, 40, 700, 3000, 3838, 3839, 3840, 3841, 4095, 4096, 6000, 8000, 8188, 8189, 8190, 8191};
  double worst[3] = {0.0, 0.0, 0.0};
  for (int m : ms) {
    const int K = (LMAX - m) / 2 + 1;
    std::vector<double> t((size_t)K * glb::PREP_TAB);
    glb::prep_tables_for_m(LMAX, m, t.data());
    const int k = (LMAX - m) / 2;  // the highest even-offset l
    const int l = m + 2 * k;
    // scale of the function over these rings, to quote errors relative to its maximum
    ld fmax = 0.0L;
    std::vector<ld> ref;
    for (int r : rings) {
      ld z, s;
      ring_ld(r, z, s);
      const ld v = ratio_ref(l, m, z) * powl(s, (ld)m);
      ref.push_back(v);
      if (fabsl(v) > fmax) fmax = fabsl(v);
    }
    for (size_t q = 0; q < sizeof(rings) / sizeof(rings[0]); ++q) {
      double z, s;
      ring_d(rings[q], z, s);
      const double sm = std::pow(s, (double)m);
      for (int mode = 0; mode < 3; ++mode) {
        const double got = ratio_kernel(t, k, z, s, mode) * sm;
        const double e = (double)(fabsl((ld)got - ref[q]) / fmax);
        if (e > worst[mode]) worst[mode] = e;
      }
    }
  }
  std::printf("worst error relative to the function's maximum, l ~ %d:\n", LMAX);
  std::printf("  per-ring variable (kernel) %.2e\n  y = z^2 everywhere        %.2e\n  u = sin^2 everywhere      %.2e\n", worst[0],
              worst[1], worst[2]);
  if (!(worst[0] < 2e-11)) return 1;                       // what the kernel does
  if (!(worst[1] > 10 * worst[0] && worst[2] > 10 * worst[0])) return 1;  // why
  std::printf("legendre recurrence ok\n");
  return 0;
}
