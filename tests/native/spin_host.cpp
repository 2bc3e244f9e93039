// Host replay of the recurrence arithmetic of spin_legendre_synth_kernel (csrc/sht_spin.cu) at
// nside 4096, lmax 8191, spin 2: coefficient table as spin_tables_kernel computes it (plain
// double), the per-ring choice of the variable (x = cos(theta) where x < 1/2, t = 1 - x =
// 2 sin^2(theta/2) otherwise) and the FMA chain  p_{l+1} = fma(fma(v, P, Q), p_l, -p_{l-1})
// against the same recurrence in 80-bit arithmetic on exact ring geometry.  Also replays "x
// everywhere" to document why the kernel switches.  Built and run by tests/test_cpu_host.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

typedef long double ld;
static const int N = 4096, LMAX = 8191, S = 2;

template <typename T>
static void tables(int m, std::vector<T>& A, std::vector<T>& B) {
  const int l0 = std::max(m, S);
  auto R = [&](int l) {
    const T dl = l, dm = m, ds = S;
    const T v = ((dl - dm) * (dl + dm)) * ((dl - ds) * (dl + ds));
    return (T)std::sqrt(std::max(v, (T)0));
  };
  T sig_lm1 = 1, sig_l = 1, R_l = R(l0);
  for (int l = l0; l <= LMAX; ++l) {
    const T dl = l, dm = m, ds = S;
    const T R_lp1 = R(l + 1);
    const T nr = std::sqrt((2 * dl + 3) / (2 * dl + 1));
    const T Al = (2 * dl + 1) * (dl + 1) * nr / R_lp1;
    const T Bl = (2 * dl + 1) * dm * ds * nr / (dl * R_lp1);
    T sig_lp1 = 1;
    if (l > l0) sig_lp1 = (dl + 1) * R_l * std::sqrt((2 * dl + 3) / (2 * dl - 1)) / (dl * R_lp1) * sig_lm1;
    const T ratio = sig_l / sig_lp1;
    A.push_back(Al * ratio);
    B.push_back(Bl * ratio);
    sig_lm1 = sig_l;
    sig_l = sig_lp1;
    R_l = R_lp1;
  }
}

int main() {
  const int ms[] = {0, 1, 3};
  // pole, cap, around x = 1/2 (ring pair ~5120: z = (2N - i) 2 / (3N) = 1/2), belt, equator
  const int rings[] = {0, 1, 2, 7, 100, 1500, 2896, 4095, 5118, 5119, 5120, 5121, 6000, 8191};
  double worst_k = 0.0, worst_x = 0.0;
  for (int m : ms) {
    std::vector<double> A, B;
    std::vector<ld> AL, BL;
    tables<double>(m, A, B);
    tables<ld>(m, AL, BL);
    const int nl = (int)A.size() - 1;
    std::vector<ld> ref;
    ld fmax = 0;
    std::vector<double> got_k, got_x;
    for (int r : rings) {
      const int i = r + 1;
      ld zx, tx;  // exact cos(theta) and 1 - cos(theta)
      double zd, td;
      if (i < N) {
        tx = (ld)i * i / (3.0L * N * N);
        zx = 1.0L - tx;
        const double sh = std::sqrt(0.5 * ((double)i * i / (3.0 * N * N)));
        td = 2.0 * sh * sh;  // 2 sh^2 with sh = sqrt(omz / 2), as the kernel
        zd = 1.0 - (double)i * i / (3.0 * N * N);
      } else {
        zx = (2.0L * N - i) * 2.0L / (3.0L * N);
        tx = 1.0L - zx;
        zd = (2.0 * N - i) * 2.0 / (3.0 * N);
        const double sh = std::sqrt(0.5 * (1.0 - zd));
        td = 2.0 * sh * sh;
      }
      // reference: lam+ chain (P = A, Q = +B) in long double with exact x
      ld p1 = 0, p2 = 1;
      for (int k = 0; k < nl; ++k) {
        const ld t = (zx * AL[k] + BL[k]) * p2 - p1;
        p1 = p2;
        p2 = t;
      }
      ref.push_back(p2);
      fmax = std::max(fmax, fabsl(p2));
      for (int mode = 0; mode < 2; ++mode) {
        const bool use_t = (mode == 0) && zd >= 0.5;
        double q1 = 0, q2 = 1;
        for (int k = 0; k < nl; ++k) {
          const double rp = use_t ? std::fma(td, -A[k], A[k] + B[k]) : std::fma(zd, A[k], B[k]);
          const double t = std::fma(rp, q2, -q1);
          q1 = q2;
          q2 = t;
        }
        (mode == 0 ? got_k : got_x).push_back(q2);
      }
    }
    for (size_t q = 0; q < ref.size(); ++q) {
      worst_k = std::max(worst_k, (double)(fabsl((ld)got_k[q] - ref[q]) / fmax));
      worst_x = std::max(worst_x, (double)(fabsl((ld)got_x[q] - ref[q]) / fmax));
    }
  }
  std::printf("spin-2 recurrence to l = %d, worst error relative to the function's maximum:\n", LMAX);
  std::printf("  per-ring variable (kernel) %.2e\n  x = cos(theta) everywhere  %.2e\n", worst_k, worst_x);
  if (!(worst_k < 5e-11)) return 1;
  if (!(worst_x > 5 * worst_k)) return 1;
  std::printf("spin recurrence ok\n");
  return 0;
}
