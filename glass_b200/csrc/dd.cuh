// dd.cuh -- double-double arithmetic (unevaluated sum hi + lo, ~106 significant bits) for the
// once-per-plan coefficient tables of the Legendre stages.  Plain IEEE operations plus fma, so
// the same code runs on the device and, for the unit test (tests/native/tables_host.cpp),
// on the host.
#pragma once
#include <math.h>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define GLB_DD_HD __host__ __device__ __forceinline__
#else
#define GLB_DD_HD inline
#endif

namespace glb {

struct dd {
  double hi, lo;
};

GLB_DD_HD dd dd_from(double a) { return dd{a, 0.0}; }
GLB_DD_HD dd dd_fast_two_sum(double a, double b) {  // |a| >= |b|
  const double s = a + b;
  return dd{s, b - (s - a)};
}
GLB_DD_HD dd dd_two_sum(double a, double b) {
  const double s = a + b;
  const double bb = s - a;
  return dd{s, (a - (s - bb)) + (b - bb)};
}
GLB_DD_HD dd dd_two_prod(double a, double b) {
  const double p = a * b;
  return dd{p, fma(a, b, -p)};
}
GLB_DD_HD dd dd_add(dd a, dd b) {
  dd s = dd_two_sum(a.hi, b.hi);
  const dd t = dd_two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = dd_fast_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return dd_fast_two_sum(s.hi, s.lo);
}
GLB_DD_HD dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
GLB_DD_HD dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
GLB_DD_HD dd dd_mul(dd a, dd b) {
  dd p = dd_two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return dd_fast_two_sum(p.hi, p.lo);
}
GLB_DD_HD dd dd_mul_d(dd a, double b) {
  dd p = dd_two_prod(a.hi, b);
  p.lo += a.lo * b;
  return dd_fast_two_sum(p.hi, p.lo);
}
GLB_DD_HD dd dd_div(dd a, dd b) {
  const double q1 = a.hi / b.hi;
  dd r = dd_sub(a, dd_mul_d(b, q1));
  const double q2 = r.hi / b.hi;
  r = dd_sub(r, dd_mul_d(b, q2));
  const double q3 = r.hi / b.hi;
  dd q = dd_fast_two_sum(q1, q2);
  return dd_add(q, dd_from(q3));
}
GLB_DD_HD dd dd_sqrt(dd a) {  // a >= 0
  if (a.hi <= 0.0) return dd{0.0, 0.0};
  const double s = sqrt(a.hi);
  const dd r = dd_sub(a, dd_two_prod(s, s));
  return dd_fast_two_sum(s, r.hi / (2.0 * s));
}
GLB_DD_HD double dd_to_double(dd a) { return a.hi + a.lo; }

}  // namespace glb
