// alm_sample.cu -- correlated a_lm sampling across shells (K2), m-major from the start.
//
// Replaces, per shell (glass/fields.py:404-425):
//     z   = rng.standard_normal((N_lm, 2)) @ [1, 1j]                      fields.py:407
//     alm = sum_i multalm(y[i], w[:, i + mis])                            fields.py:420, harmonics.py:46-47
//     alm = _glass_to_healpix_alm(alm)                                    fields.py:422, 943-962
//     alm[:n] = Re + Im  (m = 0 modes real)                               fields.py:425
//
// The deviates are a pure function of (seed, shell, GLASS index j = l(l+1)/2 + m) through
// Philox4x32-10 + Box-Muller, but are written straight into m-major (HEALPix) order, so
// the Python loop of fancy-index gathers in _glass_to_healpix_alm disappears.  The
// combine reproduces the reference's operation order ((0 + z0*w0) + z1*w1) + ... with
// separately rounded products and sums, so that with *supplied* z it is bit-exact.
#include "common.cuh"
#include "rng.cuh"
#include "iternorm_core.cuh"

namespace glb {

// grid: (ceil((lmax+1)/256), lmax+1): blockIdx.y = m, threads over l
__global__ void __launch_bounds__(256) alm_draw_kernel(int lmax, uint32_t k0, uint32_t k1, uint32_t stream,
                                                       double2* __restrict__ z) {
  const int m = blockIdx.y;
  const int l = m + blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  const uint64_t j = (uint64_t)l * (l + 1) / 2 + m;  // GLASS (l-major) index = Philox counter
  const Philox4 r = philox4x32_10((uint32_t)j, (uint32_t)(j >> 32), stream, RNG_TAG_ALM, k0, k1);
  const double u1 = u01_open_closed(r.v[0], r.v[1]);
  const double u2 = u01_closed_open(r.v[2], r.v[3]);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  const int64_t idx = (int64_t)m * (2 * lmax + 1 - m) / 2 + l;
  z[idx] = make_double2(rad * c, rad * s);
}

__global__ void __launch_bounds__(256) alm_glass_to_healpix_kernel(int lmax, const double2* __restrict__ in,
                                                                    double2* __restrict__ out) {
  const int m = blockIdx.y;
  const int l = m + blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  out[(int64_t)m * (2 * lmax + 1 - m) / 2 + l] = in[(int64_t)l * (l + 1) / 2 + m];
}

constexpr int MAX_TERMS = 64;
struct CombineArgs {
  const double2* z[MAX_TERMS];
};

// alm[l,m] = ((z0*w[l,0]) + z1*w[l,1]) + ... ; m = 0: alm = Re + Im
// More than MAX_TERMS terms: consecutive launches over chunks of the terms, the running sum kept
// in alm (resume != 0) and the m = 0 fold applied by the last one (fold != 0) -- the same
// left-to-right order of separately rounded operations as one pass.
__global__ void __launch_bounds__(256) alm_combine_kernel(int lmax, int nterms, const CombineArgs a,
                                                          const double* __restrict__ w, int w_stride,
                                                          double2* __restrict__ alm, int resume, int fold) {
  const int m = blockIdx.y;
  const int l = m + blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  const int64_t idx = (int64_t)m * (2 * lmax + 1 - m) / 2 + l;
  const double* wl = w + (int64_t)l * w_stride;
  double re = 0.0, im = 0.0;
  if (resume) {
    const double2 v = alm[idx];
    re = v.x;
    im = v.y;
  }
  for (int i = 0; i < nterms; ++i) {
    const double2 zi = a.z[i][idx];
    const double wi = wl[i];
    // no FMA contraction: products and sums rounded separately as NumPy does
    re = __dadd_rn(re, __dmul_rn(zi.x, wi));
    im = __dadd_rn(im, __dmul_rn(zi.y, wi));
  }
  if (m == 0 && fold) {
    re = __dadd_rn(re, im);
    im = 0.0;
  }
  alm[idx] = make_double2(re, im);
}

__global__ void __launch_bounds__(256) almxfl_kernel(int lmax, const double* __restrict__ fl, int nfl,
                                                     double2* __restrict__ alm) {
  const int m = blockIdx.y;
  const int l = m + blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  const int64_t idx = (int64_t)m * (2 * lmax + 1 - m) / 2 + l;
  const double f = (l < nfl) ? fl[l] : 0.0;
  double2 v = alm[idx];
  v.x *= f;
  v.y *= f;
  alm[idx] = v;
}

// -------------------------------------------------------------------------------------
// K1: one step of the iterative-normal recursion (glass/fields.py:101-188) for all l at once.
// One thread per l; state m [k][k][n], a [k][n], s [n] with l fastest, so every access of a
// warp is one coalesced line.  The host version costs O(n k^2) NumPy time per shell -- 19 ms at
// k = 19, 180 ms at k = 59 (60 fully correlated shells, lmax 8191) against 39 ms of GPU time for
// the shell itself; here a step moves 3 k^2 n doubles (0.7 GB at k = 59).
//   atm = a m;  m <- [[m, 0], [-atm / s', u / s']] cropped to its last k rows and columns
//   (s' = s, u = 1 where s > 0, else s' = 1, u = 0);  a = m c, c_j = row[k - j];
//   s = sqrt(row[0] - a.a)  (negative: flag);  w = [a, s]
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) iternorm_step_kernel(int n, int k, int first, const double* __restrict__ row,
                                                            double* __restrict__ m, double* __restrict__ a,
                                                            double* __restrict__ s, double* __restrict__ tmp,
                                                            double* __restrict__ w, int* __restrict__ flag) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  const bool bad = iternorm_step_one(n, k, first != 0, row + (int64_t)l * (k + 1), m + l, a + l, s + l, tmp + l,
                                     w + (int64_t)l * (k + 1));
  if (bad) atomicOr(flag, 1);
}

static inline dim3 lm_grid(int lmax) { return dim3((unsigned)((lmax + 1 + 255) / 256), (unsigned)(lmax + 1)); }

}  // namespace glb

using namespace glb;

extern "C" {

int glb_alm_draw(int lmax, uint64_t seed, uint32_t shell, double* d_z, void* stream) {
  GLB_REQUIRE(lmax >= 0 && lmax < 65535, "lmax out of range");
  GLB_REQUIRE(d_z != nullptr, "null pointer");
  alm_draw_kernel<<<lm_grid(lmax), 256, 0, (cudaStream_t)stream>>>(lmax, (uint32_t)seed, (uint32_t)(seed >> 32), shell,
                                                                   reinterpret_cast<double2*>(d_z));
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_alm_glass_to_healpix(int lmax, const double* d_in, double* d_out, void* stream) {
  GLB_REQUIRE(lmax >= 0 && lmax < 65535, "lmax out of range");
  GLB_REQUIRE(d_in && d_out && d_in != d_out, "null or aliased pointers");
  alm_glass_to_healpix_kernel<<<lm_grid(lmax), 256, 0, (cudaStream_t)stream>>>(
      lmax, reinterpret_cast<const double2*>(d_in), reinterpret_cast<double2*>(d_out));
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_alm_combine(int lmax, int nterms, const double* const* h_zptrs, const double* d_w, int w_stride,
                    double* d_alm, void* stream) {
  GLB_REQUIRE(lmax >= 0 && lmax < 65535, "lmax out of range");
  GLB_REQUIRE(nterms >= 1, "nterms must be positive");
  GLB_REQUIRE(h_zptrs && d_w && d_alm, "null pointer");
  GLB_REQUIRE(w_stride >= nterms, "w_stride smaller than nterms");
  for (int t0 = 0; t0 < nterms; t0 += MAX_TERMS) {
    const int nt = nterms - t0 < MAX_TERMS ? nterms - t0 : MAX_TERMS;
    CombineArgs a;
    for (int i = 0; i < MAX_TERMS; ++i) a.z[i] = (i < nt) ? reinterpret_cast<const double2*>(h_zptrs[t0 + i]) : nullptr;
    alm_combine_kernel<<<lm_grid(lmax), 256, 0, (cudaStream_t)stream>>>(lmax, nt, a, d_w + t0, w_stride,
                                                                        reinterpret_cast<double2*>(d_alm), t0 > 0,
                                                                        t0 + nt == nterms);
    GLB_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  return GLB_OK;
}

int glb_iternorm_step(int n, int k, int first, const double* d_row, double* d_m, double* d_a, double* d_s,
                      double* d_tmp, double* d_w, int* d_flag, void* stream) {
  GLB_REQUIRE(n >= 1 && k >= 0, "bad size");
  GLB_REQUIRE(d_row && d_s && d_w && d_flag && (k == 0 || (d_m && d_a && d_tmp)), "null pointer");
  iternorm_step_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, k, first, d_row, d_m, d_a, d_s,
                                                                                     d_tmp, d_w, d_flag);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_almxfl(int lmax, double* d_alm, const double* d_fl, int nfl, void* stream) {
  GLB_REQUIRE(lmax >= 0 && lmax < 65535, "lmax out of range");
  GLB_REQUIRE(d_alm && d_fl && nfl >= 0, "null pointer");
  almxfl_kernel<<<lm_grid(lmax), 256, 0, (cudaStream_t)stream>>>(lmax, d_fl, nfl, reinterpret_cast<double2*>(d_alm));
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

}  // extern "C"
