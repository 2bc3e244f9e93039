// iternorm_core.cuh -- one step of the iterative-normal recursion for ONE multipole
// (glass/fields.py:101-188), host/device.  The kernel (alm_sample.cu, K1) calls it with one
// thread per l on state arrays whose fastest index is l; tests/native/iternorm_host.cpp runs the
// same function on the CPU against vectors produced by executing the reference's source.
//
// What it computes: shell i is  x_i = sum_{j=1..k} a_j z_{i-j} + s z_i  with independent unit
// normals z, such that Cov(x_i, x_{i-j}) = c_j (row[j], j = 0..k).  The state keeps the k x k
// window M of the INVERSE of the (banded, lower-triangular) factor that maps z to x, so that the
// new row is a = M c and s^2 = c_0 - |a|^2.  Advancing the window by one shell appends the row
// [-a^T M, 1] / s and drops the oldest row and column.  A shell with s = 0 (e.g. a monopole
// that is identically zero) contributes no new deviate: its row is appended as [-a^T M, 0]
// undivided, as the reference does (fields.py:166-170).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GLB_ITN_HD __host__ __device__ __forceinline__
#else
#define GLB_ITN_HD inline
#endif

namespace glb {

// N: element stride between consecutive entries of the state of this multipole
//    (m[(r*k + c)*N], a[r*N], tmp[c*N]; N = number of multipoles in the kernel's layout).
// row: c_0..c_k of this multipole (contiguous), w: [a_1..a_k, s] (contiguous).
// Returns true where the reference raises "covariance matrix is not positive definite".
GLB_ITN_HD bool iternorm_step_one(int64_t N, int k, bool first, const double* row, double* m, double* a, double* s,
                                  double* tmp, double* w) {
  if (!first && k > 0) {
    const double sv = *s;
    const bool live = sv > 0.0;
    const double inv_num = live ? 1.0 : 0.0, den = live ? sv : 1.0;
    for (int c = 0; c < k; ++c) {  // tmp = a^T M
      double acc = 0.0;
      for (int r = 0; r < k; ++r) acc += a[r * N] * m[((int64_t)r * k + c) * N];
      tmp[c * N] = acc;
    }
    for (int r = 0; r + 1 < k; ++r) {  // slide the window up-left by one
      for (int c = 0; c + 1 < k; ++c) m[((int64_t)r * k + c) * N] = m[((int64_t)(r + 1) * k + c + 1) * N];
      m[((int64_t)r * k + k - 1) * N] = 0.0;
    }
    for (int c = 0; c + 1 < k; ++c) m[((int64_t)(k - 1) * k + c) * N] = -tmp[(c + 1) * N] / den;
    m[((int64_t)(k - 1) * k + k - 1) * N] = inv_num / den;
  }
  double norm2 = 0.0;
  for (int r = 0; r < k; ++r) {  // a = M c, c_j = row[k - j] (oldest shell first)
    double acc = 0.0;
    for (int c = 0; c < k; ++c) acc += m[((int64_t)r * k + c) * N] * row[k - c];
    a[r * N] = acc;
    w[r] = acc;
    norm2 += acc * acc;
  }
  const double s2 = row[0] - norm2;
  const double sv = sqrt(s2);
  *s = sv;
  w[k] = sv;
  return s2 < 0.0;
}

}  // namespace glb
