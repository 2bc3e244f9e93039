// sht_tables.cuh -- the l,m-only coefficient table of the scalar Legendre stages, one m at a time.
//
//   tab[k] = {a_k, b_k, a_k + b_k, alpha_k, s1_k = alpha_k / e_{l+1}, c_k = e_{l+2} / e_{l+3}},  l = m + 2k
//   e_l = sqrt((l^2 - m^2) / (4 l^2 - 1)),  alpha_{k+1} = alpha_{k-1} e_l e_{l-1} / (e_{l+1} e_{l+2}),
//   a_k = alpha_k / (e_{l+1} e_{l+2} alpha_{k+1}),  b_k = -(e_{l+1}^2 + e_l^2) a_k
// so that lambda_{m+2k} = alpha_k p_k with p_{k+1} = (a_k x^2 + b_k) p_k - p_{k-1}.
//
// Evaluated in double-double and rounded once.  In plain double the product a_k a_{k-1} --
// which must equal 1 / (e_l e_{l-1})^2 for the "- p_{k-1}" of the rescaled recurrence to be
// exact -- came out with a mean relative error of +1 ulp, a SYSTEMATIC perturbation that the
// double root of the recurrence at the poles amplifies by k^2 / 2: 5e-10 of the map at
// l = 8191 (found by tests/test_gpu_fullsize.py; with correctly rounded entries the errors are
// independent and the same sum stays at 5e-12).  Host-testable: tests/native/tables_host.cpp.
#pragma once
#include "dd.cuh"

namespace glb {

constexpr int PREP_TAB = 6;
constexpr int TAB_A = 0, TAB_B = 1, TAB_AB = 2, TAB_ALPHA = 3, TAB_S1 = 4, TAB_C = 5;

GLB_DD_HD dd tab_eps2(int l, int m) {  // e_l^2; numerator and denominator are exact in double
  if (l <= m) return dd{0.0, 0.0};
  const double dl = (double)l, dm = (double)m;
  return dd_div(dd_from((dl - dm) * (dl + dm)), dd_from(4.0 * dl * dl - 1.0));
}

GLB_DD_HD void prep_tables_for_m(int lmax, int m, double* t) {
  const int K = (lmax - m) / 2 + 1;
  const dd one = dd_from(1.0);
  dd alpha_km1 = dd{0.0, 0.0}, alpha_k = one;
  dd e2_l = dd{0.0, 0.0};  // e_l^2 (l = m: zero)
  dd e_lm1 = dd{0.0, 0.0}, e_l = dd{0.0, 0.0};
  dd e2_lp1 = tab_eps2(m + 1, m), e2_lp2 = tab_eps2(m + 2, m);
  dd e_lp1 = dd_sqrt(e2_lp1), e_lp2 = dd_sqrt(e2_lp2);
  for (int k = 0; k < K; ++k) {
    const int l = m + 2 * k;
    const dd e2_lp3 = tab_eps2(l + 3, m), e2_lp4 = tab_eps2(l + 4, m);
    const dd e_lp3 = dd_sqrt(e2_lp3), e_lp4 = dd_sqrt(e2_lp4);
    const dd ee = dd_mul(e_lp1, e_lp2);
    const dd alpha_kp1 = (k == 0) ? one : dd_mul(alpha_km1, dd_div(dd_mul(e_l, e_lm1), ee));
    const dd a = dd_div(alpha_k, dd_mul(ee, alpha_kp1));
    const dd d = dd_add(e2_lp1, e2_l);  // e_{l+1}^2 + e_l^2
    double* tk = t + (long long)k * PREP_TAB;
    tk[TAB_A] = dd_to_double(a);
    tk[TAB_B] = -dd_to_double(dd_mul(d, a));
    tk[TAB_AB] = dd_to_double(dd_mul(dd_sub(one, d), a));
    tk[TAB_ALPHA] = dd_to_double(alpha_k);
    tk[TAB_S1] = dd_to_double(dd_div(alpha_k, e_lp1));
    tk[TAB_C] = dd_to_double(dd_div(e_lp2, e_lp3));
    alpha_km1 = alpha_k;
    alpha_k = alpha_kp1;
    e_lm1 = e_lp1;
    e_l = e_lp2;
    e2_l = e2_lp2;
    e_lp1 = e_lp3;
    e_lp2 = e_lp4;
    e2_lp1 = e2_lp3;
    e2_lp2 = e2_lp4;
  }
}

}  // namespace glb
