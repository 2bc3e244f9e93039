// Host build of glass_b200/csrc/iternorm_core.cuh (the per-multipole body of K1,
// iternorm_step_kernel) behind the signature of glb_iternorm_step without the stream, so that the
// CPU suite can run the product's recursion against the vectors made by executing the reference's
// source (glass/fields.py:101-188) and use it as the test double of the C-ABI call in the
// host-flow tests.  Built as a shared library by tests/helpers.py::native_iternorm().
#include <cmath>
#include <cstdint>

#include "../../glass_b200/csrc/iternorm_core.cuh"

extern "C" int iternorm_step_host(int n, int k, int first, const double* row, double* m, double* a, double* s,
                                  double* tmp, double* w, int* flag) {
  for (int l = 0; l < n; ++l)
    if (glb::iternorm_step_one(n, k, first != 0, row + (int64_t)l * (k + 1), m + l, a + l, s + l, tmp + l,
                               w + (int64_t)l * (k + 1)))
      *flag |= 1;
  return 0;
}
